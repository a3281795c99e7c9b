"""GPKernel — host mirror of `ravest.gp.GPKernel` (gp.py:13-156).

Only the record and the validation live here; the covariance is built inside the batched
Cholesky kernel (csrc/rvlp_gp.cuh) from the four hyperparameters.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

SUPPORTED_KERNELS = ["Quasiperiodic"]


class GPKernel:
    def __init__(self, kernel_type: str) -> None:
        self.kernel_type = kernel_type
        if self.kernel_type == "Quasiperiodic":
            self.expected_hyperparams = ["gp_amp", "gp_lambda_e", "gp_lambda_p", "gp_period"]
        else:
            raise ValueError(f"Unsupported kernel type: {kernel_type}. Supported kernels: {SUPPORTED_KERNELS}")

    def get_expected_hyperparams(self) -> List[str]:
        return self.expected_hyperparams.copy()

    def validate_hyperparams(self, hyperparams: Dict) -> None:
        """gp.py:52-79."""
        provided, expected = set(hyperparams.keys()), set(self.expected_hyperparams)
        missing = expected - provided
        if missing:
            raise ValueError(f"Missing required hyperparameters: {missing}")
        unexpected = provided - expected
        if unexpected:
            raise ValueError(f"Unexpected hyperparameters: {unexpected}")
        self._validate_hyperparams_values({k: p.value for k, p in hyperparams.items()})

    def _validate_hyperparams_values(self, hyperparams_values: Dict[str, float]) -> None:
        """gp.py:82-108 — finite and > 0 (the kernel applies the same test per sample)."""
        for key in self.expected_hyperparams:
            if not np.isfinite(hyperparams_values[key]):
                raise ValueError(f"Non-finite hyperparameter found in: {hyperparams_values}")
        for key in self.expected_hyperparams:
            if hyperparams_values[key] <= 0:
                raise ValueError(f"{key} must be positive, got {hyperparams_values[key]}")
