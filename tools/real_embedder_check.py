"""One-off GPU check: the device GE2E encoder (k_lstm_uni_tc and the mel front end) on the reference's REAL
checkpoint notebooks/embedder.pt, against oracle/encoder_oracle.py with the same weights, on the real reference clips of
tests/golden/audio_demo_*.npz.  The 48 MB checkpoint cannot be committed: copy it to oracle/_ref/embedder.pt (git-ignored, travels
with gpurun) for the run; the log is committed as profiles/r02_real_embedder_check.txt."""
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from oracle import encoder_oracle as eo
from voicesplit_b200 import synth
from voicesplit_b200.engine import MaskEngine
from voicesplit_b200.speaker_encoder import SpeakerEncoder

ckpt = os.path.join(ROOT, "oracle", "_ref", "embedder.pt")
sd = torch.load(ckpt, map_location="cpu")
print("embedder.pt keys:", {k: tuple(v.shape) for k, v in sd.items()})
dims = synth.make_dims(601, 256, 400, 600)
eng = MaskEngine(activation="mish", **dims)
eng.load_state_dict_tensors({k: torch.from_numpy(v).cuda() for k, v in synth.make_state_dict(dims, 3, "stress").items() if "num_batches" not in k})
eng.configure_audio()
enc = SpeakerEncoder(engine=eng)
enc.load_state_dict(sd)
enc = enc.cuda()
esd = {k: v.numpy() for k, v in sd.items()}
worst = 0.0
for p in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "audio_demo_*.npz"))):
    z = np.load(p)
    ref = z["ref"].astype(np.float32) / 32768.0
    got = enc.embed_wav(torch.from_numpy(ref)[None].cuda()).cpu().numpy()[0]
    want = eo.speaker_encoder(esd, eo.get_mel(ref))
    d_or, d_gold = float(np.abs(got - want).max()), float(np.abs(got - z["dvec"]).max())
    cos = float(got @ want / (np.linalg.norm(got) * np.linalg.norm(want)))
    worst = max(worst, d_or)
    print(f"{os.path.basename(p)}: ref clip {len(ref)} samples  |dvec|={np.linalg.norm(got):.4f}  max|device - oracle|={d_or:.2e}  "
          f"max|device - golden|={d_gold:.2e}  cosine={cos:.7f}")
# batched (all three clips cropped to a common length) equals one by one
refs = [np.load(p)["ref"].astype(np.float32) / 32768.0 for p in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "audio_demo_*.npz")))]
n = min(len(r) for r in refs)
batch = torch.from_numpy(np.stack([r[:n] for r in refs])).cuda()
b = enc.embed_wav(batch).cpu().numpy()
one = np.stack([enc.embed_wav(batch[i:i + 1]).cpu().numpy()[0] for i in range(3)])
print("batched vs single max diff:", float(np.abs(b - one).max()))
assert worst < 5e-4, worst
print("real embedder check OK: trained GE2E weights through k_lstm_uni_tc within", worst, "of the float64 oracle")
