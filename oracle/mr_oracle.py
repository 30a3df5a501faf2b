"""TEST INFRASTRUCTURE — scalar CPU restatement of the MR_RL hot path.

Never imported by the product (``mr_rl_b200``).  Every function cites the reference
lines (relative to /root/reference) or the third-party routine whose published
algorithm it restates.  Pinned against the live reference by
``oracle/gen_golden.py`` -> ``tests/golden/*.npz`` (scipy 1.18.1, scikit-learn 1.9.0,
numpy 2.3.5); the reference itself holds no golden vector for this path.

The third-party arithmetic restated here (sources NOT under /root/reference):
  * scipy.integrate.RK45 (scipy 1.18.1, ``integrate/_ivp/rk.py``, ``common.py``,
    ``base.py``) — Dormand–Prince 5(4) with scipy's step-size controller;
  * sklearn.gaussian_process.GaussianProcessRegressor.predict (scikit-learn 1.9.0,
    ``gaussian_process/_gpr.py``) with ``RBF + WhiteKernel``;
  * tflearn fully_connected / batch_normalization inference forms (RL/MR_ddpg.py:124-138).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# ---- constants -------------------------------------------------------------------------
TIME_SPAN = 0.030            # MR_simulator.py:12
NUMBER_ITERATIONS = 100      # MR_simulator.py:13
RTOL = TIME_SPAN / NUMBER_ITERATIONS   # MR_simulator.py:91  (3e-4)
ATOL = 1e-4                  # MR_simulator.py:91
MAX_TIMESTEPS = 50           # MR_env.py:62
MIN_DIST2GOAL = 30           # MR_env.py:63
OBS_LOW = (-5000.0, -5000.0, -5000.0, -5000.0, 0.0)          # MR_env.py:37-39
OBS_HIGH = (5000.0, 5000.0, 5000.0, 5000.0, 80000.0)
ACTION_HIGH = (20.0, 2.0 * math.pi)                          # MR_env.py:34-36
INIT_LOW, INIT_HIGH = (100.0, 100.0), (120.0, 120.0)         # MR_env.py:40-42

# scipy rk.py: RK45 tableau (only B and E matter: the RHS ignores t and y)
RK_B = (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84)
RK_E = (-71 / 57600, 0.0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40)
SAFETY, MIN_FACTOR, MAX_FACTOR = 0.9, 0.2, 10.0
ERR_EXPONENT = -1.0 / 5.0
SQRT2 = 2 ** 0.5             # common.norm: x.size ** 0.5 with n = 2

STATUS_OK, STATUS_SOLVER_FAILED, STATUS_NONFINITE, STATUS_NOISE_OVERFLOW = 0, 1, 2, 4


class NoiseCursor:
    """One env's view of the shared pre-generated standard-normal stream."""

    def __init__(self, z):
        self.z = z
        self.cursor = 0

    def normal(self, mu, sigma):
        # numpy.random.normal(mu, sigma, 1)[0] == mu + sigma * z
        v = mu + sigma * float(self.z[self.cursor])
        self.cursor += 1
        return v


class ZeroNoise:
    cursor = 0

    def normal(self, mu, sigma):
        self.cursor += 1
        return mu


def _rms2(u, v):
    # scipy common.norm: np.linalg.norm(x) / x.size ** 0.5 ; linalg.norm = sqrt(x.x)
    return math.sqrt(u * u + v * v) / SQRT2


def rhs(f_t, alpha_t, a0, sigma, mism, noise):
    """Simulator.simulate + a0_linear, MR_simulator.py:55-88.  Draw order per
    evaluation: [a0-noise,] x-noise, y-noise."""
    if mism:
        a0m = a0 + (f_t / 4) * 0.8 + noise.normal(0, sigma / 4)              # :56,78
        dx1 = a0m * f_t * float(np.cos(alpha_t + 0.1)) + noise.normal(0, sigma) + 0.2    # :79
        dx2 = a0m * f_t * float(np.sin(alpha_t - 0.15)) + noise.normal(0, sigma) - 0.1   # :80
    else:
        dx1 = a0 * f_t * float(np.cos(alpha_t)) + noise.normal(0, sigma)     # :82
        dx2 = a0 * f_t * float(np.sin(alpha_t)) + noise.normal(0, sigma)     # :83
    return dx1, dx2


@dataclass
class SimState:
    """Everything the reference carries between env steps for one env."""
    x: float = 0.0
    y: float = 0.0
    fx: float = 0.0          # integrator.f (carried derivative, stage K0 of the next step)
    fy: float = 0.0
    h_abs: float = 0.0       # integrator.h_abs
    t: float = 0.0           # integrator.t
    t_bound: float = 0.0
    counter: int = 0         # MR_Env.counter
    a0: float = 0.0          # MR_simulator.py:16
    sigma: float = 0.0       # noise_var, MR_simulator.py:18
    mism: bool = False       # is_mismatched, MR_simulator.py:19
    spx: float = 0.0         # state_prime (last RHS evaluation), MR_simulator.py:87
    spy: float = 0.0
    status: int = 0
    attempts: int = 0
    noise: object = field(default_factory=ZeroNoise)


def ctor(s: SimState, t0, act):
    """scipy RK45.__init__ + common.select_initial_step as reached from
    MR_simulator.py:31-34,46-50,90-91."""
    f_t, al = act
    s.t = t0
    s.t_bound = t0 + TIME_SPAN
    il = abs(s.t_bound - t0)
    f0x, f0y = rhs(f_t, al, s.a0, s.sigma, s.mism, s.noise)
    if il == 0.0:
        s.fx, s.fy, s.h_abs, s.spx, s.spy = f0x, f0y, 0.0, f0x, f0y
        return
    scx = ATOL + abs(s.x) * RTOL
    scy = ATOL + abs(s.y) * RTOL
    d0 = _rms2(s.x / scx, s.y / scy)
    d1 = _rms2(f0x / scx, f0y / scy)
    if d0 < 1e-5 or d1 < 1e-5:
        h0 = 1e-6
    else:
        h0 = 0.01 * d0 / d1
    h0 = min(h0, il)
    f1x, f1y = rhs(f_t, al, s.a0, s.sigma, s.mism, s.noise)
    d2 = _rms2((f1x - f0x) / scx, (f1y - f0y) / scy) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** 0.2
    s.fx, s.fy = f0x, f0y
    s.h_abs = min(100 * h0, h1, il)
    s.spx, s.spy = f1x, f1y


def sim_reset(s: SimState, init):
    """Simulator.reset_start_pos, MR_simulator.py:21-34."""
    s.x, s.y = float(init[0]), float(init[1])
    s.status = STATUS_OK
    if not (math.isfinite(s.x) and math.isfinite(s.y)):
        s.status |= STATUS_NONFINITE      # scipy check_arguments would raise ValueError
    ctor(s, 0.0, (0.0, 0.0))


def sim_step(s: SimState, act):
    """Simulator.step, MR_simulator.py:36-52 (scipy OdeSolver.step / RK45._step_impl)."""
    f_t, al = float(act[0]), float(act[1])
    s.attempts = 0
    while not (s.t - s.t_bound >= 0):
        t, x, y = s.t, s.x, s.y
        min_step = 10 * abs(float(np.nextafter(t, np.inf)) - t)
        h_abs = max(s.h_abs, min_step)
        rejected = False
        while True:
            if h_abs < min_step:
                s.status |= STATUS_SOLVER_FAILED
                return
            t_new = t + h_abs
            if t_new - s.t_bound > 0:
                t_new = s.t_bound
            h = t_new - t
            h_abs = abs(h)
            kx = [s.fx, 0, 0, 0, 0, 0, 0]
            ky = [s.fy, 0, 0, 0, 0, 0, 0]
            for i in range(1, 6):
                kx[i], ky[i] = rhs(f_t, al, s.a0, s.sigma, s.mism, s.noise)
            sx = sy = 0.0
            for i in range(6):
                sx += kx[i] * RK_B[i]
                sy += ky[i] * RK_B[i]
            xn = x + h * sx
            yn = y + h * sy
            kx[6], ky[6] = rhs(f_t, al, s.a0, s.sigma, s.mism, s.noise)
            s.attempts += 1
            scx = ATOL + max(abs(x), abs(xn)) * RTOL
            scy = ATOL + max(abs(y), abs(yn)) * RTOL
            ex = ey = 0.0
            for i in range(7):
                ex += kx[i] * RK_E[i]
                ey += ky[i] * RK_E[i]
            en = _rms2(ex * h / scx, ey * h / scy)
            if en < 1:
                if en == 0:
                    fac = MAX_FACTOR
                else:
                    fac = min(MAX_FACTOR, SAFETY * en ** ERR_EXPONENT)
                if rejected:
                    fac = min(1, fac)
                h_abs *= fac
                break
            elif en >= 1:
                h_abs *= max(MIN_FACTOR, SAFETY * en ** ERR_EXPONENT)
                rejected = True
            else:  # NaN error norm: scipy would spin shrinking by NaN; flag and stop
                s.status |= STATUS_NONFINITE
                return
        s.t, s.x, s.y, s.fx, s.fy, s.h_abs = t_new, xn, yn, kx[6], ky[6], h_abs
    if not (math.isfinite(s.x) and math.isfinite(s.y)):
        s.status |= STATUS_NONFINITE
    ctor(s, s.t, (f_t, al))          # uses the action JUST applied (MR_simulator.py:46-50)


def convert_state(x, y):
    """MR_Env.convert_state with init_goal = (0,0), MR_env.py:57,100-116."""
    return (x, y, 0.0, 0.0, math.sqrt(x * x + y * y))


def is_done(obs, counter, max_steps=None):
    """MR_Env.end, MR_env.py:136-152 with a bounds-only Box.contains (NaN -> not contained).
    max_steps: MR_Env.max_timesteps when a caller changed the attribute (default 50, MR_env.py:62)."""
    inside = all(OBS_LOW[i] <= obs[i] <= OBS_HIGH[i] for i in range(5))
    limit = MAX_TIMESTEPS if max_steps is None else max_steps
    return (not inside or counter > limit) or (obs[4] < MIN_DIST2GOAL)


def shaped_reward(obs, counter):
    """MR_Env.calculate_reward, MR_env.py:118-134 (defined but unused by step)."""
    if obs[4] < MIN_DIST2GOAL:
        return 100.0
    inside = all(OBS_LOW[i] <= obs[i] <= OBS_HIGH[i] for i in range(5))
    if not inside or counter > MAX_TIMESTEPS:
        return -100.0
    return -0.1


def env_reset(s: SimState, init, noise_var=1, a0=1, is_mismatched=False):
    """MR_Env.reset, MR_env.py:164-201: the integrator is built (:181) BEFORE
    is_mismatched is assigned (:183)."""
    s.sigma, s.a0 = float(noise_var), float(a0)
    sim_reset(s, init)
    s.mism = bool(is_mismatched)
    s.counter = 0
    return convert_state(s.x, s.y)


def env_step(s: SimState, act, max_steps=None):
    """MR_Env.step, MR_env.py:70-98."""
    s.counter += 1
    sim_step(s, act)
    obs = convert_state(s.x, s.y)
    return obs, 10, is_done(obs, s.counter, max_steps), {}


def rollout(actions, init, noise_var, a0, is_mismatched, z, mism_before_reset=False):
    """Same record layout as oracle.live_reference.rollout."""
    T = len(actions)
    s = SimState(noise=NoiseCursor(z) if z is not None else ZeroNoise())
    s.mism = mism_before_reset
    out = {
        "pos": np.zeros((T, 2)), "obs": np.zeros((T, 5)), "done": np.zeros(T, np.uint8),
        "counter": np.zeros(T, np.int32), "state_prime": np.zeros((T, 2)),
        "cursor": np.zeros(T, np.int64), "attempts": np.zeros(T, np.int32),
        "carry_f": np.zeros((T, 2)), "carry_h": np.zeros(T), "t": np.zeros(T),
        "rew": np.zeros(T), "status": np.zeros(T, np.int32),
    }
    out["reset_obs"] = np.array(env_reset(s, init, noise_var, a0, is_mismatched))
    out["reset_cursor"] = np.int64(s.noise.cursor)
    out["reset_carry_f"] = np.array([s.fx, s.fy])
    out["reset_carry_h"] = np.float64(s.h_abs)
    out["reset_state_prime"] = np.array([s.spx, s.spy])
    for k in range(T):
        obs, rew, done, _ = env_step(s, actions[k])
        out["pos"][k] = (s.x, s.y)
        out["obs"][k] = obs
        out["done"][k] = done
        out["rew"][k] = rew
        out["counter"][k] = s.counter
        out["state_prime"][k] = (s.spx, s.spy)
        out["cursor"][k] = s.noise.cursor
        out["attempts"][k] = s.attempts
        out["carry_f"][k] = (s.fx, s.fy)
        out["carry_h"][k] = s.h_abs
        out["t"][k] = s.t
        out["status"][k] = s.status
    return out


def t_table(n):
    """t after k env steps: t_k = fl(t_{k-1} + 0.03) (MR_simulator.py:49-50)."""
    t = np.zeros(n + 1)
    for k in range(n):
        t[k + 1] = t[k] + TIME_SPAN
    return t


# ---- Gaussian-process inference (Learning_module.py:10-24,186-224 -> sklearn GPR.predict) ----

@dataclass
class GPModel:
    """The fitted quantities sklearn's GPR holds after ``fit`` (Learning_module.py:122-123):
    kernel_ = RBF(length_scale) + WhiteKernel(noise_level), X_train_, alpha_, L_."""
    X_train: np.ndarray      # [n, d]
    alpha: np.ndarray        # [n]
    L: np.ndarray            # [n, n] lower Cholesky of K + (noise_level + 1e-10) I
    length_scale: float
    noise_level: float

    @classmethod
    def from_sklearn(cls, gpr):
        return cls(np.asarray(gpr.X_train_, dtype=np.float64), np.asarray(gpr.alpha_, dtype=np.float64).ravel(),
                   np.asarray(gpr.L_, dtype=np.float64), float(gpr.kernel_.k1.length_scale),
                   float(gpr.kernel_.k2.noise_level))

    def kinv(self):
        """K^-1 from the Cholesky factor (what the device variance kernel contracts with)."""
        from scipy.linalg import solve_triangular
        Linv = solve_triangular(self.L, np.eye(self.L.shape[0]), lower=True)
        return Linv.T @ Linv


def rbf_row(model: GPModel, Xq):
    """sklearn kernels.RBF.__call__(X, Y): exp(-0.5 * cdist(X/l, Y/l, 'sqeuclidean'));
    WhiteKernel contributes 0 when Y is given."""
    A = np.atleast_2d(Xq) / model.length_scale
    Bm = model.X_train / model.length_scale
    d2 = ((A[:, None, :] - Bm[None, :, :]) ** 2).sum(-1)
    return np.exp(-0.5 * d2)


def gp_predict(model: GPModel, Xq, return_std=True):
    """GaussianProcessRegressor.predict posterior branch (normalize_y=False)."""
    from scipy.linalg import solve_triangular
    Kt = rbf_row(model, Xq)
    mean = Kt @ model.alpha
    if not return_std:
        return mean
    V = solve_triangular(model.L, Kt.T, lower=True)
    var = (1.0 + model.noise_level) - np.einsum("ij,ji->i", V.T, V)   # kernel_.diag = 1 + noise
    var[var < 0] = 0.0
    return mean, np.sqrt(var)


def fit_fixed_gp(X, y, length_scale, noise_level):
    """sklearn GPR.fit with optimizer=None: K = RBF + White + 1e-10 I,
    L = chol(K), alpha = K^-1 y."""
    from scipy.linalg import cho_solve, cholesky
    X = np.atleast_2d(X).reshape(len(y), -1)
    A = X / length_scale
    d2 = ((A[:, None, :] - A[None, :, :]) ** 2).sum(-1)
    K = np.exp(-0.5 * d2)
    K[np.diag_indices_from(K)] += noise_level + 1e-10
    L = cholesky(K, lower=True, check_finite=False)
    alpha = cho_solve((L, True), y, check_finite=False)
    return GPModel(X, alpha, L, float(length_scale), float(noise_level))


def lm_error(gpx: GPModel, gpy: GPModel, vd):
    """LearningModule.error, Learning_module.py:186-196."""
    a = np.array([[math.atan2(vd[1], vd[0])]])
    mx, sx = gp_predict(gpx, a)
    my, sy = gp_predict(gpy, a)
    return mx, my, sx, sy


def lm_objective(alpha, a0, freq, v_d, gpx, gpy, Dx, Dy):
    """objective, Learning_module.py:10-24."""
    X = np.array([[float(alpha)]])
    mux = gp_predict(gpx, X, return_std=False)
    muy = gp_predict(gpy, X, return_std=False)
    return ((a0 * freq) ** 2 + (mux + Dx - v_d[0]) ** 2 + 2 * a0 * freq * np.cos(alpha) * (mux + Dx - v_d[0])
            + (muy + Dy - v_d[1]) ** 2 + 2 * a0 * freq * np.sin(alpha) * (muy + Dy - v_d[1]))


# ---- DDPG actor forward (RL/MR_ddpg.py:124-149) -------------------------------------------------

def actor_init(seed=0, s_dim=5, a_dim=2, hidden=64):
    """Random init per tflearn defaults: FC W ~ truncated-normal(std 0.02),
    b = 0; BN gamma ~ N(1, 0.002), beta = 0, moving mean 0 / var 1; last FC W ~ U[-3e-3, 3e-3]."""
    rng = np.random.default_rng(seed)

    def tn(shape):
        w = rng.standard_normal(shape)
        bad = np.abs(w) > 2
        while bad.any():
            w[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(w) > 2
        return (0.02 * w).astype(np.float32)

    p = {
        "w1": tn((s_dim, hidden)), "b1": np.zeros(hidden, np.float32),
        "g1": (1 + 0.002 * rng.standard_normal(hidden)).astype(np.float32), "be1": np.zeros(hidden, np.float32),
        "m1": np.zeros(hidden, np.float32), "v1": np.ones(hidden, np.float32),
        "w2": tn((hidden, hidden)), "b2": np.zeros(hidden, np.float32),
        "g2": (1 + 0.002 * rng.standard_normal(hidden)).astype(np.float32), "be2": np.zeros(hidden, np.float32),
        "m2": np.zeros(hidden, np.float32), "v2": np.ones(hidden, np.float32),
        "w3": rng.uniform(-0.003, 0.003, (hidden, a_dim)).astype(np.float32), "b3": np.zeros(a_dim, np.float32),
    }
    return p


def actor_forward(p, obs, bn_eps=1e-5):
    """obs[.,5] -> FC64 -> BN -> ReLU -> FC64 -> BN -> ReLU -> FC2 tanh -> * action_bound
    (RL/MR_ddpg.py:124-138, action_bound = action_space.high, :345).  fp32 like TF."""
    x = np.asarray(obs, dtype=np.float32)
    h = x @ p["w1"] + p["b1"]
    h = p["g1"] * (h - p["m1"]) / np.sqrt(p["v1"] + np.float32(bn_eps)) + p["be1"]
    h = np.maximum(h, 0)
    h = h @ p["w2"] + p["b2"]
    h = p["g2"] * (h - p["m2"]) / np.sqrt(p["v2"] + np.float32(bn_eps)) + p["be2"]
    h = np.maximum(h, 0)
    o = np.tanh(h @ p["w3"] + p["b3"])
    return o * np.asarray(ACTION_HIGH, dtype=np.float32)
