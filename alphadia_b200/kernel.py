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


class GaussianKernel:
    def __init__(
        self,
        dia_data,
        fwhm_rt: float = 10.0,
        sigma_scale_rt: float = 1.0,
        fwhm_mobility: float = 0.03,
        sigma_scale_mobility: float = 1.0,
        kernel_height: int = 30,
        kernel_width: int = 30,
    ):
        self.dia_data = dia_data
        self.fwhm_rt = fwhm_rt
        self.sigma_scale_rt = sigma_scale_rt
        self.fwhm_mobility = fwhm_mobility
        self.sigma_scale_mobility = sigma_scale_mobility
        self.kernel_height = int(np.ceil(kernel_height / 2) * 2)
        self.kernel_width = int(np.ceil(kernel_width / 2) * 2)

    def determine_rt_sigma(self, cycle_length_seconds: float):
        sigma = self.fwhm_rt / 2.3548
        return sigma * self.sigma_scale_rt / cycle_length_seconds

    def determine_mobility_sigma(self, mobility_resolution: float):
        if not self.dia_data.has_mobility:
            return 1.0
        sigma = self.fwhm_mobility / 2.3548
        return sigma * self.sigma_scale_mobility / mobility_resolution

    def get_dense_matrix(self, verbose: bool = True) -> np.ndarray:
        rt_datapoints = self.dia_data.cycle.shape[1]
        rt_resolution = np.mean(np.diff(self.dia_data.rt_values[::rt_datapoints]))
        mobility_datapoints = self.dia_data.cycle.shape[2]
        mobility_resolution = np.mean(np.diff(self.dia_data.mobility_values[::-1]))
        if verbose:
            logger.info(f"Duty cycle consists of {rt_datapoints} frames, {rt_resolution:.2f} seconds cycle time")
            logger.info(f"Duty cycle consists of {mobility_datapoints} scans, {mobility_resolution:.5f} 1/K_0 resolution")
        rt_sigma = self.determine_rt_sigma(rt_resolution)
        mobility_sigma = self.determine_mobility_sigma(mobility_resolution)
        if verbose:
            logger.info(f"FWHM in RT is {self.fwhm_rt:.2f} seconds, sigma is {rt_sigma:.2f}")
            logger.info(f"FWHM in mobility is {self.fwhm_mobility:.3f} 1/K_0, sigma is {mobility_sigma:.2f}")
        return self.gaussian_kernel_2d(self.kernel_width, self.kernel_height, rt_sigma, mobility_sigma).astype(
            np.float32
        )

    @staticmethod
    def gaussian_kernel_2d(size_x: int, size_y: int, sigma_x: float, sigma_y: float) -> np.ndarray:
        x, y = np.meshgrid(np.arange(-size_x // 2, size_x // 2), np.arange(-size_y // 2, size_y // 2))
        xy = np.column_stack((x.flatten(), y.flatten())).astype("float32")
        mu = np.array([[0.0, 0.0]])
        sigma_mat = np.array([[sigma_x, 0.0], [0.0, sigma_y]])
        weights = multivariate_normal(xy, mu, sigma_mat)
        return weights.reshape(size_y, size_x).astype(np.float32)
