"""DDPG actor forward (RL/MR_ddpg.py:124-149) on the device: 5 -> FC64 -> BN -> ReLU -> FC64 -> BN
-> ReLU -> FC2 tanh -> * action_bound, fp32: the acting path (the in-loop policy of the fused rollout).
Training, critic and replay live in ddpg.py."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L

ORDER = ("w1", "b1", "g1", "be1", "m1", "v1", "w2", "b2", "g2", "be2", "m2", "v2", "w3", "b3")
SHAPES = {"w1": (5, 64), "b1": (64,), "g1": (64,), "be1": (64,), "m1": (64,), "v1": (64,),
          "w2": (64, 64), "b2": (64,), "g2": (64,), "be2": (64,), "m2": (64,), "v2": (64,),
          "w3": (64, 2), "b3": (2,)}


def init_actor(seed=0):
    """Random init following tflearn defaults: FC weights truncated-normal(0.02),
    biases 0, BN gamma ~ N(1, 0.002), beta 0, moving mean 0 / var 1, last layer U[-3e-3, 3e-3]."""
    g = torch.Generator().manual_seed(seed)

    def tn(*shape):
        w = torch.empty(*shape)
        torch.nn.init.trunc_normal_(w, mean=0.0, std=0.02, a=-0.04, b=0.04, generator=g)
        return w

    p = {
        "w1": tn(5, 64), "b1": torch.zeros(64), "g1": 1 + 0.002 * torch.randn(64, generator=g), "be1": torch.zeros(64),
        "m1": torch.zeros(64), "v1": torch.ones(64),
        "w2": tn(64, 64), "b2": torch.zeros(64), "g2": 1 + 0.002 * torch.randn(64, generator=g), "be2": torch.zeros(64),
        "m2": torch.zeros(64), "v2": torch.ones(64),
        "w3": (torch.rand(64, 2, generator=g) * 2 - 1) * 0.003, "b3": torch.zeros(2),
    }
    return {k: v.float() for k, v in p.items()}


def pack_actor(params, device="cuda"):
    """Flatten a parameter dict (torch tensors or numpy arrays, W[in][out]) into the packed float32
    layout the kernels read (include/mr_rl_b200.h)."""
    flat = []
    for k in ORDER:
        v = torch.as_tensor(np.asarray(params[k]) if not torch.is_tensor(params[k]) else params[k]).float()
        if tuple(v.shape) != SHAPES[k]:
            raise ValueError(f"{k}: expected {SHAPES[k]}, got {tuple(v.shape)}")
        flat.append(v.reshape(-1))
    w = torch.cat(flat).to(device)
    assert w.numel() == L.load().mr_actor_param_count()
    return w


def torch_reference(params, obs, action_high=(20.0, 2 * np.pi), eps=1e-5):
    """Plain torch fp32 forward of the same network (used by tests as the numerics reference)."""
    p = {k: (torch.as_tensor(v).float()) for k, v in params.items()}
    x = torch.as_tensor(obs).float().cpu()
    h = x @ p["w1"] + p["b1"]
    h = torch.relu(p["g1"] * (h - p["m1"]) / torch.sqrt(p["v1"] + eps) + p["be1"])
    h = h @ p["w2"] + p["b2"]
    h = torch.relu(p["g2"] * (h - p["m2"]) / torch.sqrt(p["v2"] + eps) + p["be2"])
    return torch.tanh(h @ p["w3"] + p["b3"]) * torch.tensor(action_high, dtype=torch.float32)


def actor_forward(packed, obs_soa, n=None, action_high=(20.0, 2 * np.pi), env_obs=False):
    """obs_soa: [5, stride] device tensor (SoA rows, e.g. VecMREnv._obs).  Returns actions [n, 2].
    env_obs=True: the rows come from MR_Env (goal rows 2, 3 identically zero) -> tensor-core kernel."""
    lib = L.load()
    n = int(n if n is not None else obs_soa.shape[1])
    dt = {torch.float64: L.MR_F64, torch.float32: L.MR_F32}[obs_soa.dtype]
    out = torch.empty(n, 2, dtype=obs_soa.dtype, device=obs_soa.device)
    hi = (C.c_double * 2)(float(action_high[0]), float(action_high[1]))
    stream = C.c_void_p(torch.cuda.current_stream(obs_soa.device).cuda_stream)
    fn = lib.mr_actor_forward_env if env_obs else lib.mr_actor_forward
    rc = fn(packed.data_ptr(), obs_soa.data_ptr(), obs_soa.stride(0), n, dt, hi, out.data_ptr(), stream)
    L.check(rc, "mr_actor_forward_env" if env_obs else "mr_actor_forward")
    return out
