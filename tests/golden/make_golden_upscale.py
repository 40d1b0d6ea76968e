"""Golden fixture for row f4 (interpolation baselines): runs the LIVE reference functions
(src/resolution_comparison_enhanced.py, src/resolution_comparison.py) on the CPU on the 40x40 solution of
tests/golden/multires_seed2.npz and stores their float32 outputs.  Build container only.
The plotting packages the reference scripts import at module level are stubbed (they are not on the path)."""
import os, sys, types
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference/src")
for name in ("matplotlib", "matplotlib.pyplot", "seaborn", "matplotlib.colors", "matplotlib.gridspec"):
    try:
        __import__(name)
    except Exception:
        sys.modules[name] = types.ModuleType(name)
import resolution_comparison_enhanced as ref_e  # noqa: E402  (reference)
import resolution_comparison as ref_r  # noqa: E402  (reference)
from oracle import upscale_oracle as uo  # noqa: E402

mr = np.load(os.path.join(HERE, "multires_seed2.npz"))
u40 = mr["u40"]
data = {"u": {40: u40}}
out = {}
for target in (80, 160, 320):
    b = ref_e.bilinear_multi_level_upscale(data, target)
    c = ref_e.cubic_multi_level_upscale(data, target)
    assert np.array_equal(b, uo.multi_level_upscale(data, target, "bilinear")), "oracle != reference (bilinear)"
    assert np.array_equal(c, uo.multi_level_upscale(data, target, "bicubic")), "oracle != reference (bicubic)"
    assert np.array_equal(b, uo.upscale_batch(u40[None], int(np.log2(target // 40)), "bilinear")[0])
    if target <= 160:  # 320 is checked above but not stored (fixture size)
        out[f"bilinear_{target}"] = b
        out[f"bicubic_{target}"] = c
parts = ref_r.split_into_subdomains(u40, 10)
mine = uo.split_into_subdomains(u40, 10)
assert all(np.array_equal(parts[i][j], mine[i][j]) for i in range(4) for j in range(4))
assert np.array_equal(ref_r.stitch_subdomains(parts), uo.stitch_subdomains(mine))
assert np.array_equal(ref_r.stitch_subdomains(parts), u40)
np.savez_compressed(os.path.join(HERE, "upscale_seed2.npz"), **out)
print("written", {k: v.shape for k, v in out.items()})
