"""Device Gaussian-process inference for Learning_module's disturbance model.

``DeviceGP`` holds what sklearn's fitted ``GaussianProcessRegressor`` holds (X_train_, alpha_, L_,
kernel_ = RBF(l) + WhiteKernel(noise); Learning_module.py:30-33,122-123) in HBM, padded for the
tiled kernels, and evaluates ``predict(q, return_std)`` with the CUDA kernels in csrc/mr_gp.cu.
``DeviceGP.fit`` factorises on the device at given hyper-parameters (csrc/mr_gpfit.cu); the hyper-parameter search
around it is gpr.DeviceGPR; ``from_sklearn`` takes a model fitted on the host instead.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L


class DeviceGP:
    def __init__(self, X_train, alpha, L_chol, length_scale, noise_level, device="cuda", jitter=1e-10):
        self.lib = L.load()
        self.jitter = float(jitter)
        self._proj = None
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.MRLibraryError("DeviceGP needs a CUDA device: there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        X = np.asarray(X_train, dtype=np.float64)
        X = X.reshape(len(X), -1)
        n, d = X.shape
        if d not in (1, 2):
            raise ValueError("GP input dimension must be 1 (Learning_module) or 2 (Learning_module_2d)")
        self.n_train, self.dim = n, d
        self.n_pad = (n + L.GP_PAD - 1) // L.GP_PAD * L.GP_PAD
        self.length_scale = float(length_scale)
        self.noise_level = float(noise_level)
        xs = np.zeros((self.n_pad, d))
        xs[:n] = X / self.length_scale                      # kernels.RBF scales both operands first
        al = np.zeros(self.n_pad)
        al[:n] = np.asarray(alpha, dtype=np.float64).ravel()
        self._xs = torch.from_numpy(xs).to(self.device)
        self._alpha = torch.from_numpy(al).to(self.device)
        self._linv = None
        if L_chol is not None:
            from scipy.linalg import solve_triangular
            Linv = solve_triangular(np.asarray(L_chol, dtype=np.float64), np.eye(n), lower=True, check_finite=False)
            W = np.zeros((self.n_pad, self.n_pad))
            W[:n, :n] = np.tril(Linv)
            self._linv = torch.from_numpy(W).to(self.device)
        self._c = L.GPModel(self._xs.data_ptr(), self._alpha.data_ptr(),
                            self._linv.data_ptr() if self._linv is not None else None,
                            n, self.n_pad, d, 0, self.length_scale, self.noise_level)
        self._ws = None
        self.kernel_launches = 0

    @classmethod
    def from_sklearn(cls, gpr, device="cuda", with_std=True):
        """From a fitted sklearn GaussianProcessRegressor with kernel_ = RBF + WhiteKernel."""
        k = gpr.kernel_
        if getattr(gpr, "_y_train_std", 1.0) != 1.0 or np.any(np.asarray(getattr(gpr, "_y_train_mean", 0.0)) != 0.0):
            raise ValueError("normalize_y=True models are not supported (the reference uses the default False)")
        return cls(gpr.X_train_, gpr.alpha_, gpr.L_ if with_std else None, k.k1.length_scale, k.k2.noise_level, device,
                   jitter=float(np.ravel(gpr.alpha)[0]) if np.ndim(gpr.alpha) == 0 or np.size(gpr.alpha) == 1 else 1e-10)

    @classmethod
    def fit(cls, X, y, length_scale, noise_level, jitter=1e-10, device="cuda", eval_gradient=False):
        """GaussianProcessRegressor.fit at FIXED hyper-parameters, on the device (csrc/mr_gpfit.cu): what sklearn
        computes once theta is known (K, cholesky, alpha_, L^-1 and the log marginal likelihood;
        Learning_module.py:122-123).  ``jitter`` is sklearn's ``alpha`` (default 1e-10).  The optimiser over theta
        can stay on the host and call this as its objective (``eval_gradient=True``): ``log_marginal_likelihood_value_``
        and ``log_marginal_likelihood_gradient_`` (w.r.t. log length_scale, log noise_level) — see gpr.DeviceGPR."""
        self = cls.__new__(cls)
        self.lib = L.load()
        self.jitter = float(jitter)
        self._proj = None
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.MRLibraryError("DeviceGP needs a CUDA device: there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        Xt = torch.as_tensor(X).to(device=self.device, dtype=torch.float64)
        Xt = Xt.reshape(Xt.shape[0], -1).contiguous()
        yt = torch.as_tensor(y).to(device=self.device, dtype=torch.float64).reshape(-1).contiguous()
        n, d = Xt.shape
        if d not in (1, 2):
            raise ValueError("GP input dimension must be 1 (Learning_module) or 2 (Learning_module_2d)")
        if yt.numel() != n:
            raise ValueError("X and y disagree on the number of samples")
        self.n_train, self.dim = n, d
        self.n_pad = (n + L.GP_PAD - 1) // L.GP_PAD * L.GP_PAD
        self.length_scale, self.noise_level = float(length_scale), float(noise_level)
        with torch.cuda.device(self.device):
            self._xs = torch.empty((self.n_pad, d), dtype=torch.float64, device=self.device)
            self._alpha = torch.empty(self.n_pad, dtype=torch.float64, device=self.device)
            self._linv = torch.empty((self.n_pad, self.n_pad), dtype=torch.float64, device=self.device)
            scal = torch.zeros(4, dtype=torch.float64, device=self.device)       # [lml, dlml/dlog l, dlml/dlog noise, info]
            ws_bytes = int(self.lib.mr_gp_fit_workspace_bytes(self.n_pad))
            ws = torch.empty((ws_bytes + 7) // 8, dtype=torch.float64, device=self.device)
            stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            rc = self.lib.mr_gp_fit(Xt.data_ptr(), yt.data_ptr(), n, self.n_pad, d, self.length_scale, self.noise_level,
                                    float(jitter), self._xs.data_ptr(), self._alpha.data_ptr(), self._linv.data_ptr(),
                                    scal.data_ptr(), scal.data_ptr() + 8 if eval_gradient else None, scal.data_ptr() + 24,
                                    ws.data_ptr(), ws_bytes, stream)
            L.check(rc, "mr_gp_fit")
            host = scal.cpu()                                                    # one 32-byte read back
            info = int(host[3:].view(torch.int32)[0])
            if info != 0:
                raise np.linalg.LinAlgError(f"{info}-th leading minor of the kernel matrix is not positive definite")
            self.log_marginal_likelihood_value_ = float(host[0])
            self.log_marginal_likelihood_gradient_ = host[1:3].numpy().copy() if eval_gradient else None
        self._c = L.GPModel(self._xs.data_ptr(), self._alpha.data_ptr(), self._linv.data_ptr(),
                            n, self.n_pad, d, 0, self.length_scale, self.noise_level)
        self._ws = None
        self.kernel_launches = 0
        return self

    def enable_spectral_variance(self, tol=1e-9, n_probe=512, max_fraction=0.5):
        """Replace the triangular variance contraction |L^-1 k|^2 (n^2 / 2 per query) by |P k|^2 with P the leading
        eigenpairs of K scaled by mu^-1/2 (rows u_i^T / sqrt(mu_i)).  An RBF Gram matrix is numerically low rank — its
        spectrum decays like exp(-i^2 c) down to the noise floor — and k(q, X) lives in the same leading subspace, so a
        few hundred rows give k^T K^-1 k to rounding.  The switch is made only if the two forms agree to ``tol``
        (relative, in the std) on ``n_probe`` queries spread over the training range; returns the number of rows used
        (0 = kept the triangular form).
        LIBRARY CODE, SET-UP ONLY: the eigendecomposition (torch.linalg.eigh = cuSOLVER syevd, ~0.2 s for n = 2000) and the
        Gram matrix behind it (torch.cdist / exp) run once per fitted model, outside every timed predict; the time is kept
        in ``self.spectral_build_ms``.  The per-query path (mr_gp_predict) is hand-written either way, and the reference's
        own algorithm — the triangular form — stays available (never call this method) and is what bench.py reports as
        ``reference_algorithm_triangular_ms``."""
        import time as _time
        _t0 = _time.perf_counter()
        if self._linv is None or self._c.proj_rows:
            return int(self._c.proj_rows)
        n, n_pad = self.n_train, self.n_pad
        with torch.cuda.device(self.device):
            xs = self._xs[:n]
            d2 = torch.cdist(xs, xs).square_()
            K = torch.exp(-0.5 * d2)
            K.diagonal().fill_(1.0 + self.noise_level + self.jitter)
            mu, U = torch.linalg.eigh(K)                                   # ascending
            floor = self.noise_level + self.jitter
            r = int((mu - floor > 1e-12 * float(mu[-1])).sum())
            r_pad = max(32, (r + 31) // 32 * 32)                            # the fused kernel takes 32 / 64 / 96 / 128 rows ...
            if r_pad > L.GP_PAD:
                r_pad = (r + L.GP_PAD - 1) // L.GP_PAD * L.GP_PAD           # ... or whole 128-row passes
            if r_pad > max_fraction * n_pad:
                return 0
            P = torch.zeros(r_pad, n_pad, dtype=torch.float64, device=self.device)
            top = slice(n - min(r_pad, n), n)
            P[: min(r_pad, n), :n] = (U[:, top] / torch.sqrt(mu[top])).T
            lo, hi = self._xs[:n].min(0).values, self._xs[:n].max(0).values
            g = torch.Generator(device=self.device).manual_seed(0)
            probe = (lo + (hi - lo) * torch.rand(n_probe, self.dim, generator=g, device=self.device, dtype=torch.float64))
            probe = probe * self.length_scale                             # predict() takes unscaled inputs
            _, s_tri = self.predict(probe, True)
            tri_ptr = self._c.linv
            self._proj = P
            self._c.linv, self._c.proj_rows = P.data_ptr(), r_pad
            _, s_spec = self.predict(probe, True)
            err = float(((s_spec - s_tri).abs() / s_tri.clamp_min(1e-300)).max())
            if not err <= tol:
                self._c.linv, self._c.proj_rows, self._proj = tri_ptr, 0, None
                return 0
            torch.cuda.synchronize(self.device)
        self.spectral_build_ms = (_time.perf_counter() - _t0) * 1e3
        return r_pad

    def predict(self, q, return_std=False):
        """q: [n_q] or [n_q, dim] float64 (device tensor or array).  Returns mean[, std] device tensors."""
        if self.device.index is not None and torch.cuda.current_device() != self.device.index:
            with torch.cuda.device(self.device):          # kernels launch on the runtime's current device
                return self.predict(q, return_std)
        qt = torch.as_tensor(q) if not torch.is_tensor(q) else q
        qt = qt.to(device=self.device, dtype=torch.float64).reshape(-1, self.dim).contiguous()
        n_q = qt.shape[0]
        mean = torch.empty(n_q, dtype=torch.float64, device=self.device)
        std = torch.empty(n_q, dtype=torch.float64, device=self.device) if return_std else None
        ws_ptr, ws_bytes = None, 0
        if return_std:
            if self._linv is None:
                raise ValueError("model was built without the Cholesky factor; std is unavailable")
            ws_bytes = int(self.lib.mr_gp_workspace_bytes(C.byref(self._c), n_q, 1))
            if self._ws is None or self._ws.numel() * 8 < ws_bytes:
                self._ws = torch.empty((ws_bytes + 7) // 8, dtype=torch.float64, device=self.device)
            ws_ptr = self._ws.data_ptr()
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        rc = self.lib.mr_gp_predict(C.byref(self._c), qt.data_ptr(), n_q, mean.data_ptr(),
                                    std.data_ptr() if std is not None else None, ws_ptr, ws_bytes, stream)
        L.check(rc, "mr_gp_predict")
        self.kernel_launches += 1
        return (mean, std) if return_std else mean
