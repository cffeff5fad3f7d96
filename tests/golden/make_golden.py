"""Generate golden vectors by running the UNMODIFIED reference modules (imported from
/root/reference, CPU fp32) on seeded weights and inputs.  Run in the build container:

    python tests/golden/make_golden.py

Writes tests/golden/case_*.npz.  Weights are NOT stored: they are regenerated from the seed
by voicesplit_b200.synth.make_state_dict (numpy PCG64, stable); only the seeds, dims and the
reference outputs are stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from voicesplit_b200 import synth  # noqa: E402

# name, model, dims, B, T, flavour, weight seed, input seed
CASES = [
    ("tiny_mish_default", "voicesplit", synth.make_dims(33, 16, 24, 40), 2, 11, "default", 1, 11),
    ("tiny_mish_stress", "voicesplit", synth.make_dims(33, 16, 24, 40), 3, 37, "stress", 2, 12),
    ("tiny_relu_stress", "voicefilter", synth.make_dims(33, 16, 24, 40), 2, 37, "stress", 3, 13),
    ("t1_mish_stress", "voicesplit", synth.make_dims(17, 8, 16, 24), 1, 1, "stress", 4, 14),
    ("odd_mish_stress", "voicesplit", synth.make_dims(41, 20, 28, 36), 2, 70, "stress", 5, 15),
    ("f257_mish_stress", "voicesplit", synth.make_dims(257, 256, 400, 600), 1, 40, "stress", 6, 16),
    ("f601_mish_stress", "voicesplit", synth.make_dims(601, 256, 400, 600), 1, 24, "stress", 7, 17),
    ("f601_relu_default", "voicefilter", synth.make_dims(601, 256, 400, 600), 1, 9, "default", 8, 18),
]


def main():
    VoiceSplit, VoiceFilter, gu = ref_import.load()
    torch.set_num_threads(os.cpu_count())
    for name, model_name, dims, B, T, flavour, wseed, iseed in CASES:
        cfg = gu.AttrDict(synth.make_config_dict(dims, model_name))
        model = (VoiceSplit if model_name == "voicesplit" else VoiceFilter)(cfg).eval()
        sd = synth.make_state_dict(dims, wseed, flavour)
        model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
        x, emb = synth.make_inputs(B, T, dims, iseed)
        with torch.no_grad():
            xt, et = torch.from_numpy(x), torch.from_numpy(emb)
            mask = model(xt, et)
            conv = model.conv(xt.unsqueeze(1))                      # [B,8,T,F]
            conv_out = conv.transpose(1, 2).contiguous().view(B, T, -1)
            cat = torch.cat((conv_out, et.unsqueeze(1).repeat(1, T, 1)), dim=2)
            lstm_out, _ = model.lstm(cat)
            l3 = model.conv[:12](xt.unsqueeze(1))                   # after cnn3 (+BN+act)
        out = os.path.join(ROOT, "tests", "golden", f"case_{name}.npz")
        np.savez_compressed(
            out, model_name=model_name, flavour=flavour, wseed=wseed, iseed=iseed, B=B, T=T,
            dims=np.array([dims[k] for k in ("num_freq", "emb_dim", "lstm_dim", "fc1_dim", "fc2_dim")]),
            mask=mask.numpy(), masked=(xt * mask).numpy(),
            conv_out=conv_out.numpy().astype(np.float32),
            lstm_out=lstm_out.numpy().astype(np.float32),
            act3_sample=l3.numpy()[:, ::7, :, ::5].astype(np.float32),
            torch_version=torch.__version__)
        print(name, "mask range", float(mask.min()), float(mask.max()), os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
