"""Gaussian smoothing kernel for candidate selection (host side, numpy).

Restates ``GaussianKernel`` (alphadia/search/selection/kernel.py:47-218) including its quirks:
the covariance diagonal is (sigma_x, sigma_y) — sigma, not sigma squared (kernel.py:215) — and the
normalisation uses k = mu.shape[0] = 1 (kernel.py:37,42).  The kernel is computed once per
``CandidateSelection`` and handed to the device as a plain f32 matrix.
"""

from __future__ import annotations

import logging

import numpy as np

logger = logging.getLogger()


def multivariate_normal(x: np.ndarray, mu: np.ndarray, sigma: np.ndarray) -> np.ndarray:
    """kernel.py:14-44 (the N x N matrix of the reference only feeds its diagonal)."""
    k = mu.shape[0]
    dx = x - mu
    inv = np.linalg.inv(sigma)
    quad = np.einsum("ni,ij,nj->n", dx, inv, dx)
    a = np.exp(-1 / 2 * quad)
    b = (np.pi * 2) ** (-k / 2) * np.linalg.det(sigma) ** (-1 / 2)
    return a * b


FWHM_PER_SIGMA = 2.3548  # kernel.py:112,131: the constant the reference divides the FWHM by


def _sigma_in_grid_steps(fwhm: float, scale: float, step: float) -> float:
    """Peak width as a standard deviation in units of the grid spacing (kernel.py:98-139): ((fwhm / 2.3548) * scale) / step,
    evaluated in that order."""
    return fwhm / FWHM_PER_SIGMA * scale / step


def _density_on_grid(height: int, width: int, sigma_rows: float, sigma_cols: float) -> np.ndarray:
    """The density of ``multivariate_normal`` on the integer grid [-height/2, height/2) x [-width/2, width/2) around the
    origin, rows = mobility axis, columns = RT axis (kernel.py:185-218).  The grid points go in as float32 (column
    coordinate first), the covariance diagonal is (sigma_cols, sigma_rows) - see the module docstring."""
    cols = np.arange(-width // 2, width // 2)
    rows = np.arange(-height // 2, height // 2)
    points = np.empty((len(rows) * len(cols), 2), dtype=np.float32)
    points[:, 0] = np.tile(cols, len(rows))
    points[:, 1] = np.repeat(rows, len(cols))
    covariance = np.diag([float(sigma_cols), float(sigma_rows)])
    density = multivariate_normal(points, np.zeros((1, 2)), covariance)
    return density.reshape(len(rows), len(cols)).astype(np.float32)


class GaussianKernel:
    """Constructor arguments and ``get_dense_matrix`` as in the reference (kernel.py:47-96,141-183)."""

    def __init__(self, dia_data, fwhm_rt: float = 10.0, sigma_scale_rt: float = 1.0, fwhm_mobility: float = 0.03,
                 sigma_scale_mobility: float = 1.0, kernel_height: int = 30, kernel_width: int = 30):
        self.dia_data = dia_data
        self.fwhm_rt, self.sigma_scale_rt = fwhm_rt, sigma_scale_rt
        self.fwhm_mobility, self.sigma_scale_mobility = fwhm_mobility, sigma_scale_mobility
        # both extents are rounded up to even numbers (kernel.py:93-96)
        self.kernel_height, self.kernel_width = (int(np.ceil(v / 2) * 2) for v in (kernel_height, kernel_width))

    def get_dense_matrix(self, verbose: bool = True) -> np.ndarray:
        """f32 ``[kernel_height, kernel_width]``: the grid spacing is the mean cycle time / the mean scan-to-scan mobility step of
        the file; data without ion mobility gets a mobility sigma of 1 grid step (kernel.py:126-128)."""
        data = self.dia_data
        frames_per_cycle, scans = data.cycle.shape[1], data.cycle.shape[2]
        cycle_seconds = np.mean(np.diff(data.rt_values[::frames_per_cycle]))
        mobility_step = np.mean(np.diff(data.mobility_values[::-1]))
        sigma_rt = _sigma_in_grid_steps(self.fwhm_rt, self.sigma_scale_rt, cycle_seconds)
        sigma_mobility = (_sigma_in_grid_steps(self.fwhm_mobility, self.sigma_scale_mobility, mobility_step)
                          if data.has_mobility else 1.0)
        if verbose:
            logger.info(f"Duty cycle: {frames_per_cycle} frames in {cycle_seconds:.2f} s, {scans} scans {mobility_step:.5f} 1/K_0 apart")
            logger.info(f"FWHM {self.fwhm_rt:.2f} s -> sigma {sigma_rt:.2f} cycles; FWHM {self.fwhm_mobility:.3f} 1/K_0 -> "
                        f"sigma {sigma_mobility:.2f} scans")
        return _density_on_grid(self.kernel_height, self.kernel_width, sigma_mobility, sigma_rt)
