"""Drop-in for the reference's ``PDEDataset`` (src/models.py:132-207) -- the immediate consumer of the
generated samples ("next" row f1 of SURVEY.md section 8).

The float32 cast, the global mean / std, the z-scores, the bilinear 20->40 upsample of the coarse solution and
the channel concat run as two CUDA kernels over the whole dataset (``psn_dataset_tensors``); the tensors never
leave the GPU.  Attribute names, ``__len__``, ``__getitem__`` (returns ``(x[3,nf,nf], y[1,nf,nf])``) and
``denormalize`` follow the reference.  The U-Net and the training loops themselves are out of scope.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from . import _lib


class PDEDataset:
    def __init__(self, data_dict: dict, device: str = 'cuda'):
        import torch

        if not str(device).startswith('cuda'):
            raise _lib.PsnError("PDEDataset of this package builds its tensors with a CUDA kernel: device must be a CUDA device")
        self.device = device
        dev = torch.device(device)
        f64 = lambda k: torch.from_numpy(np.ascontiguousarray(data_dict[k], dtype=np.float64)).to(dev)
        uc, uf, ff, th = f64('u_coarse'), f64('u_fine'), f64('f_fine'), f64('theta_fine')
        n, nf = uf.shape[0], uf.shape[-1]
        self.x = torch.empty((n, 3, nf, nf), dtype=torch.float32, device=dev)
        self.y = torch.empty((n, 1, nf, nf), dtype=torch.float32, device=dev)
        _, _, stats = _lib.dataset_tensors(uc, uf, ff, th, x=self.x, y=self.y, stream=torch.cuda.current_stream(dev).cuda_stream)
        torch.cuda.synchronize(dev)
        # raw float32 copies, as the reference keeps them (src/models.py:144-147)
        self.u_coarse, self.u_fine, self.f_fine, self.theta_fine = uc.float(), uf.float(), ff.float(), th.float()
        self.has_subdomain_flag = 'is_subdomain' in data_dict
        if self.has_subdomain_flag:
            self.is_subdomain = torch.from_numpy(np.asarray(data_dict['is_subdomain'])).bool().to(dev)
        t32 = lambda v: torch.tensor(v, dtype=torch.float32, device=dev)
        self.u_mean, self.u_std = t32(stats[0]), t32(stats[1])
        self.f_mean, self.f_std = t32(stats[2]), t32(stats[3])
        self.theta_is_constant = bool(stats[6] != 0.0)
        if self.theta_is_constant:
            self.theta_mean, self.theta_std = 0, 1
        else:
            self.theta_mean, self.theta_std = t32(stats[4]), t32(stats[5])
        # normalised views (same names as the reference)
        self.u_coarse_upsampled = self.x[:, 0:1]
        self.theta_fine_norm = self.x[:, 1]
        self.f_fine_norm = self.x[:, 2]
        self.u_fine_norm = self.y[:, 0]
        self.u_coarse_norm = (self.u_coarse - self.u_mean) / self.u_std

    def __len__(self) -> int:
        return self.y.shape[0]

    def __getitem__(self, idx: int) -> Tuple["torch.Tensor", "torch.Tensor"]:
        return self.x[idx], self.y[idx]

    def denormalize(self, x):
        """src/models.py:205-207"""
        return x * self.u_std + self.u_mean
