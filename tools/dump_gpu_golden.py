"""Developer tool: evaluate the golden kinodynamic fixtures on the GPU and save the outputs (gpurun_out/) for
entry-by-entry analysis on the CPU box."""
import glob
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from hippopt_b200.evaluator import ALL, KinoEvaluator  # noqa: E402
from hippopt_b200.kino_layout import KinoSettings  # noqa: E402
from hippopt_b200.robot_model import synthetic_ergocub  # noqa: E402

model = synthetic_ergocub()
os.makedirs("gpurun_out", exist_ok=True)
for path in sorted(glob.glob("tests/golden/kino_*.npz")):
    d = np.load(path)
    smooth = bool(d["smooth"])
    ev = KinoEvaluator(model, KinoSettings(horizon=int(d["horizon"]), final_state_constraint=bool(d["final"]),
                                           periodicity_constraint=bool(d["periodicity"]),
                                           terrain="smooth_steps" if smooth else "planar",
                                           n_terrain_params=10 if smooth else 0))
    t = [torch.tensor(np.ascontiguousarray(d[k]), device="cuda:0") for k in ("x", "p", "lam", "sigma")]
    out = ev.eval(ALL, *t)
    torch.cuda.synchronize()
    np.savez_compressed(os.path.join("gpurun_out", "gpu_" + os.path.basename(path)), **{k: v.cpu().numpy() for k, v in out.items()})
    print("dumped", path)
