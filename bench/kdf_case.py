"""One quickselect-faithful k-d build of 1 M points (for an ncu launch list of its rounds)."""
import numpy as np, torch
from grid_interpolation_b200 import interpolation as ip
rng = np.random.default_rng(0)
p = torch.as_tensor(rng.uniform(0, 4095.0, (2, 1_000_000)), device="cuda")
ip.kd_build(p, faithful=True); torch.cuda.synchronize()
print("done")
