import ctypes, os, sys
os.environ["VOICESPLIT_LSTM_TIMING"] = "1"   # compile-time timers are off in the product kernel
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from voicesplit_b200 import synth
from voicesplit_b200.engine import MaskEngine
dims = synth.make_dims(257, 256, 400, 600)
eng = MaskEngine(activation="mish", **dims)
eng.load_state_dict_tensors({k: torch.from_numpy(v).cuda() for k, v in synth.make_state_dict(dims, 0, "stress").items() if v.dtype == np.float32})
names = ["prod_spin", "prod_issue", "mma_wait", "mma_issue", "cell_wait_acc", "cell_math", "cell_store", "cell_barrier"]
for B in (32, 256):
    T = 601
    x, emb = synth.make_inputs(B, T, dims, 1)
    x, emb = torch.from_numpy(x).cuda(), torch.from_numpy(emb).cuda()
    eng.forward(x, emb, precision="fp16x3"); eng.forward(x, emb, precision="fp16x3")
    torch.cuda.synchronize()
    out = (ctypes.c_int64 * 8)()
    eng.lib.vs_debug_lstm_timing(eng.handle, out)
    print(B, {n: round(out[i] / (T - 1)) for i, n in enumerate(names)}, "cycles/step")
