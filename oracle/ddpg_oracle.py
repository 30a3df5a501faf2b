"""TEST INFRASTRUCTURE — CPU restatement of the reference's DDPG learner (RL/MR_ddpg.py), plain torch + autograd.

PARITY UNPINNED for the networks and optimisers: the reference learner is TensorFlow-1 / tflearn code and neither
package exists in this image, so that part of the restatement cannot be checked against the reference running here;
it follows the published semantics of the ops the reference calls, each cited below.  PINNED: the two TensorFlow-free
classes — OUNoise and ReplayBuffer are imported from the reference tree behind stubs (oracle/live_reference.load_ddpg)
and their behaviour is recorded in tests/golden/ddpg_host.npz (oracle/gen_golden.py:gen_ddpg_host).
Only tests/ may import this module.

  networks           ActorNetwork.create_actor_network   RL/MR_ddpg.py:124-139
                     CriticNetwork.create_critic_network RL/MR_ddpg.py:202-222 (action enters the 2nd hidden layer;
                     t1's bias is created but never used)
  batch norm         tflearn batch_normalization with tflearn.is_training never switched on (the reference runs raw
                     session calls): gamma * (x - moving_mean) / sqrt(moving_var + 1e-5) + beta, moving stats frozen
                     at (0, 1), gamma / beta trainable
  critic update      loss = mean((y - Q(s, a))^2), tf.train.AdamOptimizer(critic_lr)        :194-197, :290-296
  actor update       tf.gradients(scaled_out, params, -dQ/da) / batch_size, Adam(actor_lr)  :107-116, :298-301
  target update      theta' <- tau * theta + (1 - tau) * theta' over the trainable variables :98-103, :183-188
  Adam (TF1)         lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t); theta -= lr_t * m / (sqrt(v) + eps), eps = 1e-8
  OU noise           OUNoise.__call__   :67-71
  replay             ReplayBuffer       :16-56 (deque ring, random.sample without replacement)
"""
from __future__ import annotations

import math

import torch

ACTOR_ORDER = ("w1", "b1", "g1", "be1", "m1", "v1", "w2", "b2", "g2", "be2", "m2", "v2", "w3", "b3")
ACTOR_FROZEN = ("m1", "v1", "m2", "v2")
CRITIC_ORDER = ("wc1", "bc1", "gc", "bec", "mc", "vc", "t1", "t1b", "t2", "t2b", "wo", "bo")
CRITIC_SHAPES = {"wc1": (5, 64), "bc1": (64,), "gc": (64,), "bec": (64,), "mc": (64,), "vc": (64,),
                 "t1": (64, 32), "t1b": (32,), "t2": (2, 32), "t2b": (32,), "wo": (32,), "bo": (1,)}
CRITIC_FROZEN = ("mc", "vc")           # t1b is trainable but receives no gradient (not in the graph)
BN_EPS = 1e-5


def init_critic(seed=1):
    """tflearn defaults: FC weights truncated normal (std 0.02), biases 0, BN gamma ~ N(1, 0.002), last layer
    U[-3e-3, 3e-3] (RL/MR_ddpg.py:219-221)."""
    g = torch.Generator().manual_seed(seed)

    def tn(*shape):
        w = torch.empty(*shape)
        torch.nn.init.trunc_normal_(w, mean=0.0, std=0.02, a=-0.04, b=0.04, generator=g)
        return w

    return {"wc1": tn(5, 64), "bc1": torch.zeros(64), "gc": 1 + 0.002 * torch.randn(64, generator=g), "bec": torch.zeros(64),
            "mc": torch.zeros(64), "vc": torch.ones(64), "t1": tn(64, 32), "t1b": torch.zeros(32), "t2": tn(2, 32),
            "t2b": torch.zeros(32), "wo": (torch.rand(32, generator=g) * 2 - 1) * 0.003, "bo": torch.zeros(1)}


def actor_forward(p, s, bound):
    h = s @ p["w1"] + p["b1"]
    h = torch.relu(p["g1"] * (h - p["m1"]) / torch.sqrt(p["v1"] + BN_EPS) + p["be1"])
    h = h @ p["w2"] + p["b2"]
    h = torch.relu(p["g2"] * (h - p["m2"]) / torch.sqrt(p["v2"] + BN_EPS) + p["be2"])
    return torch.tanh(h @ p["w3"] + p["b3"]) * bound


def critic_forward(c, s, a):
    h = s @ c["wc1"] + c["bc1"]
    h = torch.relu(c["gc"] * (h - c["mc"]) / torch.sqrt(c["vc"] + BN_EPS) + c["bec"])
    h = torch.relu(h @ c["t1"] + a @ c["t2"] + c["t2b"])
    return h @ c["wo"] + c["bo"]                                    # [B]


class Adam:
    def __init__(self, params, lr, b1=0.9, b2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps, self.t = lr, b1, b2, eps, 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def step(self, params, grads):
        self.t += 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        for k, g in grads.items():
            if g is None:
                continue
            self.m[k] = self.b1 * self.m[k] + (1 - self.b1) * g
            self.v[k] = self.b2 * self.v[k] + (1 - self.b2) * g * g
            params[k] = params[k] - lr_t * self.m[k] / (torch.sqrt(self.v[k]) + self.eps)


class DDPGOracle:
    def __init__(self, actor, critic, actor_target_init, critic_target_init, bound=(20.0, 2 * math.pi), gamma=0.99, tau=0.001,
                 actor_lr=1e-3, critic_lr=1e-2, dtype=torch.float32):
        self.dtype = dtype
        self.actor = {k: torch.as_tensor(v).to(dtype).clone() for k, v in actor.items()}
        self.critic = {k: torch.as_tensor(v).to(dtype).clone() for k, v in critic.items()}
        # the target graphs are initialised independently, then update_target_network() runs once (:237-241)
        self.actor_t = {k: tau * self.actor[k] + (1 - tau) * torch.as_tensor(v).to(dtype) for k, v in actor_target_init.items()}
        self.critic_t = {k: tau * self.critic[k] + (1 - tau) * torch.as_tensor(v).to(dtype) for k, v in critic_target_init.items()}
        self.bound = torch.tensor(bound, dtype=dtype)
        self.gamma, self.tau = gamma, tau
        self.opt_a = Adam(self.actor, actor_lr)
        self.opt_c = Adam(self.critic, critic_lr)

    def update(self, s, a, r, d, s2):
        """One pass of the block RL/MR_ddpg.py:285-305 on an already sampled batch.  Returns (critic loss, mean Q)."""
        s, a, r, d, s2 = (torch.as_tensor(x).to(self.dtype) for x in (s, a, r, d, s2))
        B = s.shape[0]
        with torch.no_grad():
            q2 = critic_forward(self.critic_t, s2, actor_forward(self.actor_t, s2, self.bound))
            y = r + self.gamma * q2 * (1 - d)
        cp = {k: v.clone().requires_grad_(k not in CRITIC_FROZEN) for k, v in self.critic.items()}
        q = critic_forward(cp, s, a)
        loss = torch.mean((y - q) ** 2)
        names = [k for k in CRITIC_ORDER if k not in CRITIC_FROZEN]
        grads = torch.autograd.grad(loss, [cp[k] for k in names], allow_unused=True)
        self.opt_c.step(self.critic, dict(zip(names, grads)))
        # actor: dQ/da at a = mu(s) under the UPDATED critic
        ap = {k: v.clone().requires_grad_(k not in ACTOR_FROZEN) for k, v in self.actor.items()}
        a_out = actor_forward(ap, s, self.bound)
        a_in = a_out.detach().clone().requires_grad_(True)
        dqda = torch.autograd.grad(critic_forward(self.critic, s, a_in).sum(), a_in)[0]
        names_a = [k for k in ACTOR_ORDER if k not in ACTOR_FROZEN]
        ga = torch.autograd.grad(a_out, [ap[k] for k in names_a], grad_outputs=-dqda)
        self.opt_a.step(self.actor, {k: g / B for k, g in zip(names_a, ga)})
        for net, tgt, frozen in ((self.actor, self.actor_t, ACTOR_FROZEN), (self.critic, self.critic_t, CRITIC_FROZEN)):
            for k in net:
                if k not in frozen:
                    tgt[k] = self.tau * net[k] + (1 - self.tau) * tgt[k]
        return float(loss.detach()), float(q.detach().mean())


class OUNoise:
    """RL/MR_ddpg.py:58-78 with the normal draws supplied by the caller."""

    def __init__(self, shape, mu=0.0, sigma=0.3, theta=0.15, dt=1e-2):
        self.mu, self.sigma, self.theta, self.dt = mu, sigma, theta, dt
        self.x = torch.zeros(shape, dtype=torch.float64)

    def __call__(self, z):
        self.x = self.x + self.theta * (self.mu - self.x) * self.dt + self.sigma * math.sqrt(self.dt) * torch.as_tensor(z, dtype=torch.float64)
        return self.x
