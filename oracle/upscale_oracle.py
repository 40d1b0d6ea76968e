"""CPU restatement of the reference's interpolation baselines and sub-domain tiling -- TEST INFRASTRUCTURE ONLY.

Follows src/resolution_comparison_enhanced.py:19-65 (``bilinear_multi_level_upscale`` /
``cubic_multi_level_upscale``: float32 cast, repeated x2 ``F.interpolate(..., align_corners=True)`` from 40 up
to the target) and src/resolution_comparison.py:123-158 (``split_into_subdomains`` / ``stitch_subdomains``)
call by call, with torch on the CPU.  Pinned against the live reference functions in
tests/golden/make_golden_upscale.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def multi_level_upscale(data: dict, target_resolution: int, mode: str) -> np.ndarray:
    """src/resolution_comparison_enhanced.py:19-41 (mode='bilinear'), :43-65 (mode='bicubic')"""
    current = torch.from_numpy(np.asarray(data["u"][40])).float().unsqueeze(0).unsqueeze(0)
    current_res = 40
    while current_res < target_resolution:
        next_res = current_res * 2
        current = F.interpolate(current, size=(next_res, next_res), mode=mode, align_corners=True)
        current_res = next_res
    return current.squeeze().numpy()


def upscale_batch(u: np.ndarray, levels: int, mode: str) -> np.ndarray:
    """The same cascade on a batch [B, n, n] from an arbitrary start size (what the C ABI exposes)."""
    cur = torch.from_numpy(np.ascontiguousarray(u)).float().unsqueeze(1)
    for _ in range(levels):
        m = cur.shape[-1] * 2
        cur = F.interpolate(cur, size=(m, m), mode=mode, align_corners=True)
    return cur[:, 0].numpy()


def split_into_subdomains(array: np.ndarray, subdomain_size: int) -> list:
    """src/resolution_comparison.py:123-139"""
    n = array.shape[0] // subdomain_size
    out = []
    for i in range(n):
        out.append([array[i * subdomain_size:(i + 1) * subdomain_size, j * subdomain_size:(j + 1) * subdomain_size]
                    for j in range(n)])
    return out


def stitch_subdomains(subdomains: list) -> np.ndarray:
    """src/resolution_comparison.py:141-158"""
    nr, nc = len(subdomains), len(subdomains[0])
    s = subdomains[0][0].shape[0]
    res = np.zeros((nr * s, nr * s))
    for i in range(nr):
        for j in range(nc):
            res[i * s:(i + 1) * s, j * s:(j + 1) * s] = subdomains[i][j]
    return res
