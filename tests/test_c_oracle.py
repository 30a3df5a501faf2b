"""CPU: the plain-C oracle (oracle/mr_oracle.c) against the golden vectors of the live reference."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import c_oracle
from test_oracle_golden import SINGLE_CASES


@pytest.mark.parametrize("name", SINGLE_CASES)
def test_c_oracle_single(golden_single, name):
    g = golden_single.case(name)
    sig, a0, mism, prior = g["params"]
    r = c_oracle.rollout(np.asarray(g["init"], dtype=np.float64)[None], g["actions"][:, None, :2], sig, a0, bool(mism),
                         bool(prior), z=g["z"][None], want_attempts=True)
    assert r["bad"] == 0
    assert np.array_equal(r["done"][:, 0], g["done"])
    assert np.array_equal(r["attempts"][:, 0], g["attempts"])
    assert int(r["cursor"][0]) == int(g["cursor"][-1])
    assert rel_err(r["pos"][:, 0], g["pos"]) < 1e-9
    assert rel_err(r["final"][0, 2:4], g["carry_f"][-1]) < 1e-9 and rel_err(r["final"][0, 4], g["carry_h"][-1]) < 1e-9


@pytest.mark.parametrize("tag", ["sigma0", "sigma1", "mismatch"])
def test_c_oracle_batch(golden_batch, tag):
    sig, a0, mism, _ = golden_batch[f"{tag}/params"]
    r = c_oracle.rollout(golden_batch["init"].astype(np.float64), golden_batch["actions"], sig, a0, bool(mism), bool(mism) and False,
                         z=golden_batch["z"])
    assert r["bad"] == 0
    assert np.array_equal(r["done"], golden_batch[f"{tag}/done"])
    assert np.array_equal(r["cursor"], golden_batch[f"{tag}/cursor"][-1])
    assert rel_err(r["pos"], golden_batch[f"{tag}/pos"]) < 1e-9


def test_c_oracle_internal_noise_statistics():
    n, T = 4096, 2
    init = np.tile([[110.0, 105.0]], (n, 1))
    acts = np.zeros((T, n, 2)); acts[..., 0] = 5.0
    r = c_oracle.rollout(init, acts, 1.0, 1.0, seed=1)
    v = (r["pos"][1] - r["pos"][0]) / 0.03           # mean velocity over a step = 5 + averaged noise
    assert abs(v[:, 0].mean() - 5.0) < 0.05 and abs(v[:, 1].mean()) < 0.05
