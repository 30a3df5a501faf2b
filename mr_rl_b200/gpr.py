"""``DeviceGPR``: the slice of sklearn's ``GaussianProcessRegressor`` that Learning_module.py uses
(:28-31 constructor, :122-123 fit, :17-18/:193-194 predict, :126 score), with every O(n^2) / O(n^3) piece on the GPU.

The hyper-parameter search keeps sklearn's structure (_gpr.py ``fit``): L-BFGS-B on theta = (log length_scale,
log noise_level) from the kernel's initial value, then ``n_restarts_optimizer`` restarts drawn log-uniformly from
the bounds with the estimator's RandomState, best optimum wins.  Only the optimiser's bookkeeping (two scalars per
iterate) lives on the host; each objective evaluation is one ``mr_gp_fit`` call (kernel matrix, blocked Cholesky,
L^-1, alpha, log marginal likelihood and its gradient — csrc/mr_gpfit.cu).  Kernel family: RBF + WhiteKernel only.
"""
from __future__ import annotations

import numpy as np

from .gp import DeviceGP


class _RBFWhite:
    """What callers read off ``gpr.kernel_`` for Sum(RBF, WhiteKernel): .k1.length_scale, .k2.noise_level, .theta."""

    class _Part:
        pass

    def __init__(self, length_scale, noise_level, ls_bounds, noise_bounds):
        self.k1, self.k2 = self._Part(), self._Part()
        self.k1.length_scale, self.k1.length_scale_bounds = float(length_scale), tuple(ls_bounds)
        self.k2.noise_level, self.k2.noise_level_bounds = float(noise_level), tuple(noise_bounds)

    @property
    def theta(self):
        return np.log([self.k1.length_scale, self.k2.noise_level])

    @property
    def bounds(self):
        return np.log([self.k1.length_scale_bounds, self.k2.noise_level_bounds])

    def __repr__(self):
        return f"RBF(length_scale={self.k1.length_scale:.3g}) + WhiteKernel(noise_level={self.k2.noise_level:.3g})"


class DeviceGPR:
    def __init__(self, length_scale=1.0, length_scale_bounds=(1e-2, 10.0), noise_level=1.0, noise_level_bounds=(1e-5, 1e5),
                 alpha=1e-10, optimizer="fmin_l_bfgs_b", n_restarts_optimizer=0, random_state=None, device="cuda"):
        self.kernel = _RBFWhite(length_scale, noise_level, length_scale_bounds, noise_level_bounds)
        self.alpha = float(alpha)
        self.optimizer = optimizer
        self.n_restarts_optimizer = int(n_restarts_optimizer)
        self.random_state = random_state
        self.device = device
        self.n_objective_evals = 0

    # ---- objective ----------------------------------------------------------------------------------
    def _fit_at(self, theta, eval_gradient):
        ls, noise = np.exp(theta)
        self.n_objective_evals += 1
        return DeviceGP.fit(self._Xd, self._yd, float(ls), float(noise), jitter=self.alpha, device=self.device,
                            eval_gradient=eval_gradient)

    def log_marginal_likelihood(self, theta=None, eval_gradient=False):
        if theta is None:
            return self.log_marginal_likelihood_value_
        gp = self._fit_at(np.asarray(theta, float), eval_gradient)
        if eval_gradient:
            return gp.log_marginal_likelihood_value_, gp.log_marginal_likelihood_gradient_
        return gp.log_marginal_likelihood_value_

    def _optimise(self, theta0, bounds):
        import scipy.optimize

        def obj(theta):
            try:
                lml, grad = self.log_marginal_likelihood(theta, eval_gradient=True)
            except np.linalg.LinAlgError:          # sklearn: a failed factorisation scores -inf with zero gradient
                return np.inf, np.zeros_like(theta)
            return -lml, -grad

        res = scipy.optimize.minimize(obj, theta0, method="L-BFGS-B", jac=True, bounds=bounds)
        return res.x, res.fun

    def fit(self, X, y):
        import torch
        from sklearn.utils import check_random_state
        if torch.is_tensor(X):                                   # already in HBM (mr_learn_preprocess): no round trip
            self._Xd = X.to(device=self.device, dtype=torch.float64).reshape(X.shape[0], -1).contiguous()
            self._yd = torch.as_tensor(y).to(device=self.device, dtype=torch.float64).reshape(-1).contiguous()
            X, y = self._Xd.cpu().numpy(), self._yd.cpu().numpy()
        else:
            X = np.asarray(X, dtype=np.float64)
            X = X.reshape(len(X), -1)
            y = np.asarray(y, dtype=np.float64).ravel()
            self._Xd = torch.from_numpy(X).to(self.device)
            self._yd = torch.from_numpy(y).to(self.device)
        self.X_train_, self.y_train_ = X, y
        self._rng = check_random_state(self.random_state)
        theta0, bounds = self.kernel.theta, self.kernel.bounds
        if self.optimizer is not None:
            # (tried: the independent restarts from host threads on separate streams, to fill the GPU during the serial
            # diagonal-block steps — identical optima but 0.97 s instead of 0.37 s for the reference-sized fit: launches
            # from several threads serialise in the driver and every stream grows its own allocator pool)
            optima = [self._optimise(theta0, bounds)]
            for _ in range(self.n_restarts_optimizer):
                if not np.isfinite(bounds).all():
                    raise ValueError("Multiple optimizer restarts (n_restarts_optimizer>0) requires that all bounds are finite.")
                optima.append(self._optimise(self._rng.uniform(bounds[:, 0], bounds[:, 1]), bounds))
            theta = optima[int(np.argmin([o[1] for o in optima]))][0]
        else:
            theta = theta0
        ls, noise = np.exp(theta)
        self.kernel_ = _RBFWhite(ls, noise, self.kernel.k1.length_scale_bounds, self.kernel.k2.noise_level_bounds)
        self.model_ = self._fit_at(np.asarray(theta, float), eval_gradient=False)     # the factorisation kept in HBM
        self.log_marginal_likelihood_value_ = self.model_.log_marginal_likelihood_value_
        return self

    # ---- what Learning_module reads -------------------------------------------------------------------
    @property
    def alpha_(self):
        return self.model_._alpha[: self.model_.n_train].cpu().numpy()

    def device_model(self):
        return self.model_

    def predict(self, X, return_std=False):
        out = self.model_.predict(np.asarray(X, dtype=np.float64), return_std=return_std)
        if return_std:
            return out[0].cpu().numpy(), out[1].cpu().numpy()
        return out.cpu().numpy()

    def score(self, X, y):
        y = np.asarray(y, dtype=np.float64).ravel()
        u = ((y - self.predict(X)) ** 2).sum()
        v = ((y - y.mean()) ** 2).sum()
        return 1.0 - u / v
