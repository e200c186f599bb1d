"""Device fast path for the trajectory reductions of ``fl_scaling/est_scaling_params.py`` (a consumer of the simulators; the
analytics themselves stay in the reference).

* ``calc_nu_chunk`` / ``calc_var_chunk`` (EST.py:90-94, :131-138)  -> ``peeling_decoding.calc_nu_chunk_device``
* ``calc_theta_explicit_ss_bounds`` (EST.py:161-189) and ``calc_theta_explicit_ss_bounds_ppd`` (:211-243): the step x step
  correlation matrix ``DataFrame(r1s with zeros -> NaN).corr()`` comes from four exact int64 moment matrices accumulated on the
  device (``scldpc_pairwise_moments_accumulate``); the exponential fits (``curve_fit`` on <= 1000 points) stay on the host.

Same function names and arguments as the reference, except that ``r1s`` may also be a list of int32 device tensors
``[frames][steps]`` (what ``simulate_peeling_decoder_ldpc(..., device_r1=True)`` returns), so the trajectories never have
to leave the GPU, and that nothing is plotted.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib, engine


def _chunks(r1s):
    if isinstance(r1s, (list, tuple)):
        return [c.contiguous() for c in r1s]
    a = np.ascontiguousarray(np.asarray(r1s), dtype=np.int32)
    return [torch.as_tensor(a).to(engine._device())]


def pairwise_moments(r1s, start: int, stop: int, ivl: int = 1, acc: torch.Tensor | None = None) -> torch.Tensor:
    """int64 [4][K][K] = (N, Sx, Sxx, Sxy) over the sampled columns ``range(start, stop, ivl)`` (zeros are missing values),
    accumulated over all chunks (and into ``acc`` when given, e.g. across batches; all-reduce it across ranks)."""
    K = len(range(start, stop, ivl))
    chunks = _chunks(r1s)
    dev = chunks[0].device
    if acc is None:
        acc = torch.zeros((4, K, K), dtype=torch.int64, device=dev)
    for c in chunks:
        _lib.check(_lib.lib().scldpc_pairwise_moments_accumulate(ctypes.c_void_p(c.data_ptr()), int(c.shape[0]), int(c.shape[1]), int(start),
                                                                int(ivl), K, ctypes.c_void_p(acc.data_ptr()), engine._stream()))
    return acc


def corr_from_moments(acc) -> np.ndarray:
    """``DataFrame.corr()`` (Pearson, pairwise-complete observations, min_periods=1) from the four moment matrices."""
    n, sx, sxx, sxy = (np.asarray(a.cpu() if hasattr(a, "cpu") else a, dtype=np.float64) for a in acc)
    sy, syy = sx.T, sxx.T
    with np.errstate(invalid="ignore", divide="ignore"):
        cov = n * sxy - sx * sy
        c = cov / np.sqrt((n * sxx - sx * sx) * (n * syy - sy * sy))
    c[n < 2] = np.nan
    return c


def _fit_thetas(c, xdata, locs, fracs):
    from scipy.optimize import curve_fit
    thetas = []
    for l, loc in zip(fracs, locs):
        func = lambda x, theta: np.exp(-theta * np.abs(x - loc))   # noqa: E731
        ydata = c[:, int(c.shape[0] * l)]
        popt, _ = curve_fit(func, xdata, ydata, bounds=(0, 5))
        thetas.append(tuple(popt)[0])
    return float(np.mean(thetas))


def calc_theta_explicit_ss_bounds(r1s, start, stop, M):
    """EST.py:161-189: 1000 sampled steps between start and stop, fits at 50 reference steps; abscissa in units of M."""
    npoints = 1000
    ivl = int((stop - start) / npoints)
    c = corr_from_moments(pairwise_moments(r1s, start, stop, ivl))
    space = np.linspace(0.1, 0.9, 50)
    xdata = np.arange(start, stop, ivl) / M
    return _fit_thetas(c, xdata, [(start + (stop - start) * l) / M for l in space], space)


def calc_theta_explicit_ss_bounds_ppd(r1s, start, stop, M):
    """EST.py:211-243 (BP / parallel-peeling trajectories: every iteration between start and stop, 10 reference steps)."""
    c = corr_from_moments(pairwise_moments(r1s, start, stop, 1))
    space = np.linspace(0.3, 0.7, 10)
    xdata = np.arange(start, stop, 1)
    return _fit_thetas(c, xdata, [(start + (stop - start) * l) for l in space], space)
