"""TEST INFRASTRUCTURE ONLY — CPU oracle for the MR_RL rolling-microrobot hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the checker or the reported CPU baseline.  The
product package ``mr_rl_b200`` never imports from here and raises if its CUDA
library is missing.

Contents
--------
``live_reference.py``  imports the UNMODIFIED reference from ``/root/reference``
                       (build container only; the path does not exist on the GPU
                       box) behind stub ``gym``/``matplotlib``/``turtle`` modules
                       and a patched ``numpy.random.normal`` that pops from a
                       shared pre-generated standard-normal stream.  Used by
                       ``gen_golden.py`` to write ``tests/golden/*.npz``.
``mr_oracle.py``       scalar Python/numpy restatement of the reference algorithm
                       (MR_simulator.py:21-91, MR_env.py:70-201, scipy RK45
                       control flow, sklearn GPR.predict, the DDPG actor forward).
``scipy_env.py``       the same env restated on top of the real third-party
                       ``scipy.integrate.RK45`` — reproduces the reference's CPU
                       cost model; used as the reported CPU baseline.
``mr_oracle.c``        plain-C restatement (batched, OpenMP) for large-N parity and
                       as a second, much faster CPU baseline.

Parity pin: the reference holds NO golden vector or test for this path
(SURVEY.md §4, §8c).  The pin is therefore the live reference run in the build
container (scipy 1.18.1 / scikit-learn 1.9.0 / numpy 2.3.5); its outputs are
committed under ``tests/golden/`` together with ``gen_golden.py``.
"""
