"""Convert the reference's numeric input data (RCI / control zonotopes, household time series)
into one .npz the package ships.  These are DATA inputs (reference: src/assets/*.pt, loaded at
envs/interfaces/rci_env.py:20-32 and envs/simulators/household.py:59-63), not source code.

Run in the build container only:  python tools/import_assets.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

SRC = Path("/root/reference/src/assets")
DST = Path(__file__).resolve().parent.parent / "safegbpo_b200" / "assets" / "assets.npz"


def main():
    out = {}
    for system in ("pendulum", "cartpole", "quadrotor"):
        for key in ("ctrl_center", "ctrl_generators", "rci_center", "rci_generators"):
            t = torch.load(SRC / f"{system}_{key}.pt", weights_only=True)
            out[f"{system}_{key}"] = t.to(torch.float64).numpy()
    for key in ("load_data", "pv_data", "outside_temperature_data"):
        t = torch.load(SRC / f"{key}.pt", weights_only=True)
        out[key] = t.to(torch.float64).numpy()
    np.savez_compressed(DST, **out)
    print("wrote", DST, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    sys.exit(main())
