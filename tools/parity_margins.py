"""Prints, for every committed golden case (outputs of the unmodified reference) and every precision mode, the measured
max |diff| / mean |diff| of the device mask next to the bound tests/test_gpu_parity.py asserts - the margins behind "parity green".

    python tools/parity_margins.py            (needs the B200; a few seconds)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import golden_cases, load_case  # noqa: E402
from voicesplit_b200.engine import MaskEngine  # noqa: E402

TOL = {"fp32": (1e-3, 1e-4), "fp16x3": (1e-3, 1e-4), "fp16_f8c": (3e-3, 2e-4), "bf16x3": (3e-3, 1e-4), "fp16": (1.5e-1, 1e-2),
       "bf16": (6e-1, 5e-2)}


def main():
    print(f"{'case':22s} {'mode':9s} {'max|diff|':>10s} {'bound':>8s} {'mae':>10s} {'bound':>8s}")
    for path in golden_cases():
        case = load_case(path)
        eng = MaskEngine(activation="mish" if case["model_name"] == "voicesplit" else "relu", **case["dims"])
        eng.load_state_dict_tensors({k: torch.from_numpy(np.asarray(v)).cuda() for k, v in case["state_dict"].items()
                                     if "num_batches" not in k})
        x, emb = torch.from_numpy(case["x"]).cuda(), torch.from_numpy(case["emb"]).cuda()
        for mode, (tmax, tmae) in TOL.items():
            mask = eng.forward(x, emb, precision=mode)
            torch.cuda.synchronize()
            d = np.abs(mask.cpu().numpy() - case["mask"])
            print(f"{case['name']:22s} {mode:9s} {d.max():10.3e} {tmax:8.1e} {d.mean():10.3e} {tmae:8.1e}", flush=True)
        eng.close()


if __name__ == "__main__":
    main()
