"""DDPG learner on the device — the other caller of the env (RL/MR_ddpg.py).

Mirrors the reference's pieces with the same names and arguments where they exist:
``ReplayBuffer`` (:16-56) as a ring in HBM, ``OUNoise`` (:58-78) with one process per env, ``DDPGLearner`` = the
ActorNetwork / CriticNetwork pair (:80-231) with their Adam optimisers and target networks, and ``train`` (:233-323)
as a vectorised acting / learning loop.  One learner update (:285-305) is ONE kernel launch (csrc/mr_ddpg.cu).
Everything here fails without the CUDA library: there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L
from .actor import ORDER as ACTOR_ORDER, SHAPES as ACTOR_SHAPES, actor_forward, init_actor, pack_actor

CRITIC_ORDER = ("wc1", "bc1", "gc", "bec", "mc", "vc", "t1", "t1b", "t2", "t2b", "wo", "bo")
CRITIC_SHAPES = {"wc1": (5, 64), "bc1": (64,), "gc": (64,), "bec": (64,), "mc": (64,), "vc": (64,),
                 "t1": (64, 32), "t1b": (32,), "t2": (2, 32), "t2b": (32,), "wo": (32,), "bo": (1,)}


def init_critic(seed=1):
    """tflearn defaults as in create_critic_network (:202-222): truncated-normal(0.02) FC weights, zero biases,
    BN gamma ~ N(1, 0.002), moving mean 0 / variance 1, output layer U[-3e-3, 3e-3]."""
    g = torch.Generator().manual_seed(seed)

    def tn(*shape):
        w = torch.empty(*shape)
        torch.nn.init.trunc_normal_(w, mean=0.0, std=0.02, a=-0.04, b=0.04, generator=g)
        return w

    return {"wc1": tn(5, 64), "bc1": torch.zeros(64), "gc": 1 + 0.002 * torch.randn(64, generator=g), "bec": torch.zeros(64),
            "mc": torch.zeros(64), "vc": torch.ones(64), "t1": tn(64, 32), "t1b": torch.zeros(32), "t2": tn(2, 32),
            "t2b": torch.zeros(32), "wo": (torch.rand(32, generator=g) * 2 - 1) * 0.003, "bo": torch.zeros(1)}


def pack_critic(params, device="cuda"):
    flat = []
    for k in CRITIC_ORDER:
        v = torch.as_tensor(np.asarray(params[k]) if not torch.is_tensor(params[k]) else params[k]).float()
        if tuple(v.shape) != CRITIC_SHAPES[k]:
            raise ValueError(f"{k}: expected {CRITIC_SHAPES[k]}, got {tuple(v.shape)}")
        flat.append(v.reshape(-1))
    w = torch.cat(flat).to(device)
    assert w.numel() == L.load().mr_critic_param_count()
    return w


def _unpack(flat, order, shapes):
    out, o = {}, 0
    flat = flat.detach().cpu()
    for k in order:
        n = int(np.prod(shapes[k]))
        out[k] = flat[o:o + n].reshape(shapes[k]).clone()
        o += n
    return out


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _cuda_device(device):
    dev = torch.device(device)
    if dev.type != "cuda":
        raise L.MRLibraryError("the DDPG learner needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if dev.index is None else dev


class ReplayBuffer:
    """ReplayBuffer (:16-56) as a ring of float32 rows in HBM.  ``add`` takes the whole vectorised step."""

    def __init__(self, buffer_size, random_seed=123, device="cuda"):
        self.lib = L.load()
        self.device = _cuda_device(device)
        self.buffer_size = int(buffer_size)
        self.random_seed = int(random_seed)
        self.count = 0
        self.head = 0
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
        self.s, self.a, self.r, self.d, self.s2 = z(buffer_size, 5), z(buffer_size, 2), z(buffer_size), z(buffer_size), z(buffer_size, 5)
        self._c = L.Replay(self.s.data_ptr(), self.a.data_ptr(), self.r.data_ptr(), self.d.data_ptr(), self.s2.data_ptr(),
                           self.buffer_size)

    def add(self, obs_soa, actions, rew, done, obs_next_soa, n=None):
        """obs_soa / obs_next_soa: [5, stride] SoA rows (VecMREnv buffers), actions [n, 2], rew [n], done [n] uint8."""
        n = int(actions.shape[0] if n is None else n)
        if n > self.buffer_size:
            raise ValueError("more transitions in one step than the buffer holds")
        dt = {torch.float64: L.MR_F64, torch.float32: L.MR_F32}[obs_soa.dtype]
        with torch.cuda.device(self.device):
            rc = self.lib.mr_replay_add(C.byref(self._c), self.head, obs_soa.data_ptr(), obs_soa.stride(0), actions.data_ptr(),
                                        rew.data_ptr(), done.data_ptr(), obs_next_soa.data_ptr(), obs_next_soa.stride(0), n, dt,
                                        _stream(self.device))
        L.check(rc, "mr_replay_add")
        self.head = (self.head + n) % self.buffer_size
        self.count = min(self.count + n, self.buffer_size)

    def size(self):
        return self.count

    def sample_indices(self, batch_size, update_index=1):
        """``batch_size`` distinct rows of the filled part of the ring (random.sample, :43-46) as a device int64 tensor."""
        idx = torch.empty(int(batch_size), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.mr_replay_sample(self.count, int(batch_size), self.random_seed, int(update_index), idx.data_ptr(),
                                           _stream(self.device))
        L.check(rc, "mr_replay_sample")
        return idx

    def clear(self):
        self.count = 0
        self.head = 0


class OUNoise:
    """OUNoise (:58-78), one independent process per env and action dimension, state in HBM (float64)."""

    def __init__(self, num_envs, mu=0.0, sigma=0.3, theta=0.15, dt=1e-2, seed=0, env_base=0, device="cuda"):
        self.lib = L.load()
        self.device = _cuda_device(device)
        self.mu, self.sigma, self.theta, self.dt = float(mu), float(sigma), float(theta), float(dt)
        self.seed, self.env_base, self.calls = int(seed), int(env_base), 0
        self.x_prev = torch.zeros(num_envs, 2, dtype=torch.float64, device=self.device)

    def add_to(self, actions, reset_mask=None):
        """actions [n, 2] (float64 / float32 device tensor) += the next OU sample, in place."""
        dt = {torch.float64: L.MR_F64, torch.float32: L.MR_F32}[actions.dtype]
        with torch.cuda.device(self.device):
            rc = self.lib.mr_ou_noise_add(self.x_prev.data_ptr(), actions.data_ptr(),
                                          reset_mask.data_ptr() if reset_mask is not None else None, actions.shape[0], dt,
                                          self.theta, self.mu, self.sigma, self.dt, self.seed, self.calls, self.env_base,
                                          _stream(self.device))
        L.check(rc, "mr_ou_noise_add")
        self.calls += 1
        return actions

    def reset(self):
        self.x_prev.zero_()


class DDPGLearner:
    """ActorNetwork + CriticNetwork (:80-231): online and target parameters, Adam moments, one-launch update."""

    def __init__(self, actor=None, critic=None, action_bound=(20.0, 2 * math.pi), actor_lr=1e-3, critic_lr=1e-2, tau=1e-3,
                 gamma=0.99, seed=0, device="cuda", actor_target_init=None, critic_target_init=None, target_init="independent"):
        self.lib = L.load()
        self.device = _cuda_device(device)
        self.actor = pack_actor(actor if actor is not None else init_actor(seed), self.device)
        self.critic = pack_critic(critic if critic is not None else init_critic(seed + 1), self.device)
        # The reference initialises the target graphs independently (tf.global_variables_initializer) and then calls
        # update_target_network() once (:237-241), i.e. theta' = tau * theta + (1 - tau) * theta'_init — NOT a copy.
        # target_init="independent" reproduces that (with *_target_init dicts, or fresh draws); "copy" starts the
        # targets at the online values.
        if target_init == "copy":
            self.actor_target = self.actor.clone()
            self.critic_target = self.critic.clone()
        elif target_init == "independent":
            at = pack_actor(actor_target_init if actor_target_init is not None else init_actor(seed + 2), self.device)
            ct = pack_critic(critic_target_init if critic_target_init is not None else init_critic(seed + 3), self.device)
            self.actor_target = tau * self.actor + (1 - tau) * at
            self.critic_target = tau * self.critic + (1 - tau) * ct
        else:
            raise ValueError("target_init must be 'independent' or 'copy'")
        self._moments = [torch.zeros_like(self.actor), torch.zeros_like(self.actor), torch.zeros_like(self.critic),
                         torch.zeros_like(self.critic)]
        self._grads = [torch.zeros_like(self.actor), torch.zeros_like(self.critic)]
        self._c = L.DDPGState(self.actor.data_ptr(), self.actor_target.data_ptr(), self.critic.data_ptr(),
                              self.critic_target.data_ptr(), *[m.data_ptr() for m in self._moments],
                              *[g.data_ptr() for g in self._grads])
        self.hyper = L.DDPGHyper(gamma, tau, actor_lr, critic_lr, (C.c_double * 2)(*map(float, action_bound)), 0.9, 0.999, 1e-8)
        self.action_bound = tuple(map(float, action_bound))
        self.seed = int(seed)
        self.updates = 0
        self._info = torch.zeros(2, dtype=torch.float32, device=self.device)
        self._ws = None
        self.kernel_launches = 0

    def predict(self, obs_soa, n=None):
        """ActorNetwork.predict (:146-149) for the env's SoA observation rows -> actions [n, 2] (obs dtype)."""
        self.kernel_launches += 1
        return actor_forward(self.actor, obs_soa, n, self.action_bound, env_obs=True)

    def update(self, replay, batch_size=64, indices=None):
        """The update block of train() (:285-305).  ``indices`` (int64 device tensor) pins the minibatch; by default the
        kernel samples it without replacement like random.sample.  Returns the device tensor [critic loss, mean Q]."""
        idx_ptr = None
        if indices is not None:
            indices = torch.as_tensor(indices, dtype=torch.int64, device=self.device).contiguous()
            batch_size = int(indices.numel())
            idx_ptr = indices.data_ptr()
        self.updates += 1
        ws_ptr, ws_bytes = None, 0
        if batch_size > 64:                        # data-parallel kernels (one CTA per SM) need per-CTA gradient slabs
            ws_bytes = int(self.lib.mr_ddpg_workspace_bytes(int(batch_size)))
            if self._ws is None or self._ws.numel() * 8 < ws_bytes:
                self._ws = torch.empty((ws_bytes + 7) // 8, dtype=torch.float64, device=self.device)
            ws_ptr = self._ws.data_ptr()
        with torch.cuda.device(self.device):
            rc = self.lib.mr_ddpg_update(C.byref(self._c), C.byref(replay._c), replay.count, int(batch_size), idx_ptr,
                                         self.seed ^ replay.random_seed, self.updates, C.byref(self.hyper), self._info.data_ptr(),
                                         ws_ptr, ws_bytes, _stream(self.device))
        if rc:
            self.updates -= 1
        L.check(rc, "mr_ddpg_update")
        self.kernel_launches += 1
        return self._info

    def update_distributed(self, replay, batch_size=64, indices=None, group=None):
        """The same update with the gradient exchange of a data-parallel learner (one process per GPU, each with its own
        replay shard and the same parameters): every rank computes the gradients of ITS minibatch, the gradient vectors
        are summed over the ranks (torch.distributed all_reduce, NCCL) and every rank applies the mean — the effective
        minibatch is world_size * batch_size.  Without an initialised process group this is the single-rank split path.
        Returns the device tensor [critic loss, mean Q] over the global minibatch."""
        import torch.distributed as dist
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        idx_ptr = None
        if indices is not None:
            indices = torch.as_tensor(indices, dtype=torch.int64, device=self.device).contiguous()
            batch_size = int(indices.numel())
            idx_ptr = indices.data_ptr()
        ws_bytes = int(self.lib.mr_ddpg_workspace_bytes(int(batch_size)))
        if self._ws is None or self._ws.numel() * 8 < ws_bytes:
            self._ws = torch.empty((ws_bytes + 7) // 8, dtype=torch.float64, device=self.device)
        if getattr(self, "_gbuf", None) is None:
            self._gbuf = torch.zeros(self.actor.numel() + 8, dtype=torch.float32, device=self.device)
        g = self._gbuf
        self.updates += 1
        n_c, n_a = self.critic.numel(), self.actor.numel()
        seed = self.seed ^ replay.random_seed
        with torch.cuda.device(self.device):
            for which, n in ((0, n_c + 2), (1, n_a)):
                rc = self.lib.mr_ddpg_gradients(C.byref(self._c), C.byref(replay._c), replay.count, int(batch_size), idx_ptr, seed,
                                                self.updates, C.byref(self.hyper), which, g.data_ptr(), self._ws.data_ptr(),
                                                ws_bytes, _stream(self.device))
                L.check(rc, "mr_ddpg_gradients")
                if world > 1:
                    dist.all_reduce(g[:n], group=group)
                if which == 0:
                    self._info.copy_(g[n_c:n_c + 2] / float(batch_size * world))
                rc = self.lib.mr_ddpg_apply(C.byref(self._c), which, g.data_ptr(), 1.0 / world, self.updates, C.byref(self.hyper),
                                            _stream(self.device))
                L.check(rc, "mr_ddpg_apply")
        self.kernel_launches += 8
        return self._info

    def actor_params(self, target=False):
        return _unpack(self.actor_target if target else self.actor, ACTOR_ORDER, ACTOR_SHAPES)

    def critic_params(self, target=False):
        return _unpack(self.critic_target if target else self.critic, CRITIC_ORDER, CRITIC_SHAPES)


def train(env, learner, actor_noise, buffer_size=10000, min_batch=64, steps=1000, updates_per_step=1, replay=None,
          noise_var=1, a0=1, log_every=0, stale_state_warmup=False):
    """train() (:233-323) for a vectorised env: every iteration all N envs act with mu(s) + OU noise, the N transitions
    enter the ring, and once it holds ``min_batch`` transitions the learner takes ``updates_per_step`` updates.  ``env``
    is a VecMREnv built with auto_reset=True (an env that ends starts its next episode inside the same launch, the
    reference's ``break`` + ``env.reset()``).  Returns per-iteration (mean reward, critic loss, mean Q) as a numpy array.
    Forced by vectorisation: N transitions per iteration instead of one.
    ``stale_state_warmup=True`` reproduces the scalar loop's warm-up quirk: while the buffer holds fewer than
    ``min_batch`` transitions the loop ``continue``s before ``state = next_state`` (:281-284), so the policy keeps seeing —
    and the buffer keeps storing as ``s`` — the observation of the last reset, while the env itself moves on.  With
    N >= min_batch envs the buffer is full after the first iteration and the flag changes nothing."""
    n = env.num_envs
    replay = replay if replay is not None else ReplayBuffer(buffer_size, 0, device=env.device)
    obs_rows = env._obs
    env.reset(noise_var=noise_var, a0=a0)
    state = obs_rows.clone()                       # the loop's `state`: policy input and the stored s
    log = torch.zeros(steps, 3, dtype=torch.float64, device=env.device)
    for it in range(steps):
        actions = learner.predict(state, n)
        actor_noise.add_to(actions)
        _, rew, done, _ = env.step(actions)
        replay.add(state, actions, rew, done, obs_rows, n)
        log[it, 0] = rew[:n].mean()
        if replay.size() >= min_batch:
            for _ in range(updates_per_step):
                info = learner.update(replay, min_batch)
            log[it, 1:] = info.double()
            state.copy_(obs_rows)                  # state = next_state (:307)
        elif stale_state_warmup:
            ended = done[:n].bool()                # `break` -> env.reset() -> a fresh state; everyone else keeps the stale one
            state[:, :n][:, ended] = obs_rows[:, :n][:, ended]
        else:
            state.copy_(obs_rows)
        if log_every and (it + 1) % log_every == 0:
            row = log[it].cpu().numpy()
            print(f"iter {it + 1}/{steps}: mean reward {row[0]:.3f} critic loss {row[1]:.4g} mean Q {row[2]:.4g}")
    return log.cpu().numpy()
