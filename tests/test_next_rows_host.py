"""CPU checks for the host-side pieces of the 'next' rows: the torch restatement of the DDPG learner (its Adam against
torch.optim.Adam, its target initialisation, the OU recurrence) and DeviceGPR's kernel parameterisation against
sklearn's.  No GPU, no CUDA library calls."""
import math

import numpy as np
import torch


def test_oracle_adam_is_tf1_adam_and_matches_torch_when_eps_vanishes():
    """TF1: theta -= lr sqrt(1-b2^t)/(1-b1^t) m / (sqrt(v) + eps); torch differs only in where eps sits, so with a
    vanishing eps both must trace the same path on the critic's MSE loss."""
    from oracle.ddpg_oracle import CRITIC_FROZEN, Adam, critic_forward, init_critic
    g = torch.Generator().manual_seed(0)
    s, a, y = torch.randn(64, 5, generator=g) * 30, torch.rand(64, 2, generator=g) * 6, torch.randn(64, generator=g) * 5
    base = {k: v.double() for k, v in init_critic(3).items()}
    ours = {k: v.clone() for k, v in base.items()}
    opt = Adam(ours, 1e-2, eps=1e-300)
    theirs = {k: v.clone().requires_grad_(k not in CRITIC_FROZEN and k != "t1b") for k, v in base.items()}
    topt = torch.optim.Adam([v for v in theirs.values() if v.requires_grad], lr=1e-2, eps=1e-300)
    for _ in range(5):
        p = {k: v.clone().requires_grad_(k not in CRITIC_FROZEN) for k, v in ours.items()}
        loss = torch.mean((y.double() - critic_forward(p, s.double(), a.double())) ** 2)
        names = [k for k in p if p[k].requires_grad]
        grads = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
        opt.step(ours, dict(zip(names, grads)))
        topt.zero_grad()
        torch.mean((y.double() - critic_forward(theirs, s.double(), a.double())) ** 2).backward()
        topt.step()
    for k in ours:
        assert torch.allclose(ours[k], theirs[k].detach(), rtol=1e-9, atol=1e-12), k
    assert torch.equal(ours["t1b"], base["t1b"]) and torch.equal(ours["vc"], base["vc"])     # no gradient / frozen


def test_oracle_update_moves_the_right_things():
    from mr_rl_b200.actor import init_actor
    from oracle.ddpg_oracle import DDPGOracle, init_critic
    a, c, at, ct = init_actor(0), init_critic(1), init_actor(2), init_critic(3)
    o = DDPGOracle(a, c, at, ct, tau=0.001)
    # targets start at tau * theta + (1 - tau) * theta'_init (RL/MR_ddpg.py:237-241)
    assert torch.allclose(o.actor_t["w2"], 0.001 * a["w2"] + 0.999 * at["w2"])
    g = torch.Generator().manual_seed(1)
    s = torch.randn(64, 5, generator=g) * 30
    act = torch.rand(64, 2, generator=g) * torch.tensor([20.0, 6.28])
    before_t = {k: v.clone() for k, v in o.critic_t.items()}
    loss, _ = o.update(s, act, torch.full((64,), 10.0), torch.zeros(64), s + 1)
    assert loss > 0 and not torch.equal(o.critic["wo"], c["wo"]) and not torch.equal(o.actor["w3"], a["w3"])
    assert torch.equal(o.actor["m1"], a["m1"]) and torch.equal(o.critic["vc"], c["vc"])
    # one soft update: theta' moved by tau * (theta - theta')
    assert torch.allclose(o.critic_t["wo"], 0.001 * o.critic["wo"] + 0.999 * before_t["wo"])
    # a terminal transition's target ignores the bootstrap: with done = 1 everywhere y = r, so loss ~ (10 - Q)^2
    o2 = DDPGOracle(a, c, at, ct)
    l2, q2 = o2.update(s, act, torch.full((64,), 10.0), torch.ones(64), s + 1)
    assert abs(l2 - (10.0 - q2) ** 2) < 0.05 * l2


def test_ou_noise_oracle_recurrence():
    from oracle.ddpg_oracle import OUNoise
    ou = OUNoise((3, 2), sigma=0.3, theta=0.15, dt=1e-2)
    z1, z2 = torch.ones(3, 2), -torch.ones(3, 2)
    x1 = ou(z1).clone()
    assert torch.allclose(x1, torch.full((3, 2), 0.3 * math.sqrt(1e-2), dtype=torch.float64))
    x2 = ou(z2)
    assert torch.allclose(x2, x1 + 0.15 * (0 - x1) * 1e-2 - 0.3 * math.sqrt(1e-2))


def test_device_gpr_kernel_parameterisation_matches_sklearn():
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel
    from mr_rl_b200.gpr import _RBFWhite
    sk = RBF(length_scale=1.0, length_scale_bounds=(1e-2, 10.0)) + WhiteKernel()      # Learning_module.py:30
    ours = _RBFWhite(1.0, 1.0, (1e-2, 10.0), (1e-5, 1e5))
    assert np.allclose(ours.theta, sk.theta) and np.allclose(ours.bounds, sk.bounds)
    sk2 = RBF(0.3, (1e-2, 10.0)) + WhiteKernel(0.02)
    assert np.allclose(_RBFWhite(0.3, 0.02, (1e-2, 10.0), (1e-5, 1e5)).theta, sk2.theta)
    assert "RBF(length_scale=0.3)" in repr(_RBFWhite(0.3, 0.02, (1e-2, 10.0), (1e-5, 1e5)))


def test_ou_oracle_matches_the_reference_class():
    """oracle.OUNoise against RL/MR_ddpg.py's own OUNoise run on a fixed normal stream (golden ddpg_host.npz)."""
    from conftest import Golden
    from oracle.ddpg_oracle import OUNoise
    g = Golden("ddpg_host.npz")
    theta, sigma, dt = g["ou_params"]
    ou = OUNoise((2,), sigma=sigma, theta=theta, dt=dt)
    z = g["ou_z"].reshape(-1, 2)
    xs = np.array([ou(z[k]).numpy().copy() for k in range(len(z))])
    assert np.allclose(xs, g["ou_x"], rtol=1e-14, atol=1e-16)
    assert (theta, sigma, dt) == (0.15, 0.3, 1e-2)                  # the defaults our device OUNoise mirrors
