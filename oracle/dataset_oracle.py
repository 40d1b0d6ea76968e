"""CPU restatement of the reference's training-tensor prologue (``PDEDataset``) -- TEST INFRASTRUCTURE ONLY.

Follows src/models.py:132-207 of the reference call by call, on the CPU with torch (float32 cast, global
mean / unbiased std, z-score, ``F.interpolate(..., mode='bilinear', align_corners=True)`` to the fine grid,
channel concat in ``__getitem__``).  The reference hard-codes ``size=(40, 40)`` (src/models.py:184); the
restatement takes the fine size from the data so that other grid pairs can be checked too.
Pinned against the live reference class in tests/golden/make_golden_tensors.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def dataset_tensors(data: dict) -> dict:
    u_coarse = torch.from_numpy(np.ascontiguousarray(data["u_coarse"])).float()
    u_fine = torch.from_numpy(np.ascontiguousarray(data["u_fine"])).float()
    f_fine = torch.from_numpy(np.ascontiguousarray(data["f_fine"])).float()
    theta_fine = torch.from_numpy(np.ascontiguousarray(data["theta_fine"])).float()
    u_mean, u_std = u_fine.mean(), u_fine.std()
    f_mean, f_std = f_fine.mean(), f_fine.std()
    theta_is_constant = bool(theta_fine.std() < 1e-6)
    if theta_is_constant:
        theta_mean, theta_std = 0.0, 1.0
        theta_norm = theta_fine
    else:
        theta_mean, theta_std = theta_fine.mean(), theta_fine.std()
        theta_norm = (theta_fine - theta_mean) / theta_std
    u_fine_norm = (u_fine - u_mean) / u_std
    u_coarse_norm = (u_coarse - u_mean) / u_std
    f_norm = (f_fine - f_mean) / f_std
    nf = u_fine.shape[-1]
    up = F.interpolate(u_coarse_norm.unsqueeze(1), size=(nf, nf), mode="bilinear", align_corners=True)
    x = torch.cat([up, theta_norm.unsqueeze(1), f_norm.unsqueeze(1)], dim=1)
    y = u_fine_norm.unsqueeze(1)
    return {"x": x.numpy(), "y": y.numpy(),
            "stats": np.array([float(u_mean), float(u_std), float(f_mean), float(f_std), float(theta_mean), float(theta_std),
                               1.0 if theta_is_constant else 0.0, 0.0])}
