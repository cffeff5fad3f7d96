import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import golden_cases, load_case
from voicesplit_b200.engine import MaskEngine
for name in ("f601_mish_stress", "f257_mish_stress", "tiny_mish_stress"):
    case = load_case([p for p in golden_cases() if name in p][0])
    eng = MaskEngine(activation="mish", **case["dims"])
    eng.load_state_dict_tensors({k: torch.from_numpy(np.asarray(v)).cuda() for k, v in case["state_dict"].items() if "num_batches" not in k})
    conv = torch.from_numpy(case["conv_out"]).cuda(); emb = torch.from_numpy(case["emb"]).cuda(); x = torch.from_numpy(case["x"]).cuda()
    for prec in ("fp32", "fp16x3", "bf16x3", "fp16"):
        for rep in range(3):
            lo, mask = eng.debug_lstm_head(conv, emb, x, precision=prec)
            d = np.abs(lo.cpu().numpy() - case["lstm_out"])
            dm = np.abs(mask.cpu().numpy() - case["mask"])
            idx = np.unravel_index(d.argmax(), d.shape)
            print(f"{name} {prec} rep{rep}: lstm max {d.max():.3e} at {idx} mean {d.mean():.3e} | mask max {dm.max():.3e}")
