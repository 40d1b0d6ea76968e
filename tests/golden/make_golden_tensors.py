"""Golden fixture for row f1 (consumer tensor prologue): runs the LIVE reference ``PDEDataset`` (src/models.py)
on the CPU on tests/golden/dataset_standard_seed0.npz and stores its tensors.  Build container only."""
import os, sys
import numpy as np, torch
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference/src")
import models as ref_models  # noqa: E402  (reference)
from oracle.dataset_oracle import dataset_tensors  # noqa: E402

ds = np.load(os.path.join(HERE, "dataset_standard_seed0.npz"))
data = {k: ds[k] for k in ds.files}
d = ref_models.PDEDataset(data, device="cpu")
x = torch.stack([d[i][0] for i in range(len(d))]).numpy()
y = torch.stack([d[i][1] for i in range(len(d))]).numpy()
o = dataset_tensors(data)
assert np.array_equal(x, o["x"]) and np.array_equal(y, o["y"]), "oracle != reference"
np.savez_compressed(os.path.join(HERE, "dataset_tensors_seed0.npz"), x=x, y=y,
                    stats=np.array([float(d.u_mean), float(d.u_std), float(d.f_mean), float(d.f_std)]))
print("written", x.shape, y.shape)
