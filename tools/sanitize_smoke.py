"""Small end-to-end run for compute-sanitizer: eval forward in every precision, training step, audio round trip."""
import os, sys
os.environ.setdefault("VOICESPLIT_GEMM_CLUSTER", "3")     # small problems take the W-tile multicast pairs too
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from voicesplit_b200 import config, synth
from voicesplit_b200.engine import MaskEngine
from models.voicesplit.model import VoiceSplit

dims = synth.make_dims(33, 16, 24, 40)
sd = synth.make_state_dict(dims, 2, "stress")
eng = MaskEngine(activation="mish", **dims)
eng.load_state_dict_tensors({k: torch.from_numpy(v).cuda() for k, v in sd.items() if "num_batches" not in k})
x, emb = synth.make_inputs(5, 53, dims, 12)      # 265 rows: three M blocks, the last pair has an empty CTA
xt, et = torch.from_numpy(x).cuda(), torch.from_numpy(emb).cuda()
ref = eng.forward(xt, et, precision="fp32")
for p in ("fp16x3", "fp16_f8c", "bf16x3", "fp16", "bf16"):
    out = eng.forward(xt, et, precision=p, want_masked=True)[0]
    print(p, float((out - ref).abs().max()))
m = VoiceSplit(config.AttrDict(synth.make_config_dict(dims)))
m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
m = m.cuda().train()
for tc in (True, False):
    m.train_tensor_cores = tc
    m.zero_grad()
    xg = xt.clone().requires_grad_(True)
    m(xg, et).sum().backward()
    print("train tc", tc, float(m.fc2.weight.grad.abs().max()), "d/dx", float(xg.grad.abs().max()))
d601 = synth.make_dims(601, 8, 16, 24)
e2 = MaskEngine(activation="relu", **d601)
e2.load_state_dict_tensors({k: torch.from_numpy(v).cuda() for k, v in synth.make_state_dict(d601, 1, "default").items() if "num_batches" not in k})
e2.configure_audio()
w = torch.randn(2, 4000, device="cuda") * 0.05
s, ph = e2.wav2spec(w)
back = e2.spec2wav(s, ph)
print("audio", float((back - w[:, :back.shape[1]]).abs().max()))
torch.cuda.synchronize()
print("sanitize smoke done")
