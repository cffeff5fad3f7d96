"""Minimal forward driver for ncu captures (never a bench number): python tools/prof_forward.py [B] [precision] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from voicesplit_b200 import synth
if os.environ.get("VOICESPLIT_ALT_SO"):      # A/B against a library built from another commit (same ABI), tools only
    from voicesplit_b200 import _cabi
    _cabi.LIB_PATH = os.environ["VOICESPLIT_ALT_SO"]
from voicesplit_b200.engine import MaskEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16_f8c"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dims = synth.make_dims(257, 256, 400, 600)
eng = MaskEngine(activation="mish", **dims)
eng.load_state_dict_tensors({k: torch.from_numpy(v).cuda() for k, v in synth.make_state_dict(dims, 0, "stress").items() if v.dtype == np.float32})
x, emb = synth.make_inputs(B, 601, dims, 1)
x, emb = torch.from_numpy(x).cuda(), torch.from_numpy(emb).cuda()
for _ in range(iters):
    eng.forward(x, emb, precision=prec, want_masked=True)
torch.cuda.synchronize()
print("done", B, prec)
