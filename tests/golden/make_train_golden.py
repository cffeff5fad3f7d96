"""Golden vectors of the TRAINING-mode path: the UNMODIFIED reference modules (imported from /root/reference, CPU fp32) in
`.train()` mode - batch-statistics BatchNorm, running-buffer update, torch autograd through every layer - the way
train.py:84,94,108-110 drives them.  Run in the build container:

    python tests/golden/make_train_golden.py

Writes tests/golden/train_*.npz: the mask, d(loss)/d(parameter) of all 44 parameters (the five 5x5 conv weights as every 4th
element plus the sum and the sum of squares of the whole tensor, to keep the fixtures small), d(loss)/d(x), d(loss)/d(d-vector) and the
BatchNorm buffers after the step, for loss = sum(mask * gw) with a seeded gw.  Weights, inputs and gw are regenerated from the seeds
(voicesplit_b200.synth, numpy PCG64), not stored.  These pin oracle/torch_port.forward_train - the checker the GPU gradient tests
(tests/test_gpu_train.py, test_gpu_dp.py) compare the device against - to the reference itself (tests/test_train_oracle.py).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from voicesplit_b200 import synth  # noqa: E402

# name, model, dims, B, T, flavour, weight seed, input seed, gw seed   (the shapes of tests/test_gpu_train.py)
CASES = [
    ("tiny_mish_stress", "voicesplit", synth.make_dims(33, 16, 24, 40), 3, 21, "stress", 5, 6, 0),
    ("tiny_relu_stress", "voicefilter", synth.make_dims(17, 8, 16, 24), 2, 9, "stress", 5, 6, 0),
    ("odd_mish_stress", "voicesplit", synth.make_dims(41, 20, 28, 36), 2, 40, "stress", 5, 6, 0),
]


SAMPLE_ABOVE, SAMPLE_STEP = 50000, 4


def gw_for(seed, B, T, F):
    return np.random.default_rng(seed).standard_normal((B, T, F)).astype(np.float32)


def main():
    VoiceSplit, VoiceFilter, gu = ref_import.load()
    torch.set_num_threads(os.cpu_count())
    for name, model_name, dims, B, T, flavour, wseed, iseed, gseed in CASES:
        cfg = gu.AttrDict(synth.make_config_dict(dims, model_name))
        model = (VoiceSplit if model_name == "voicesplit" else VoiceFilter)(cfg)
        sd = synth.make_state_dict(dims, wseed, flavour)
        model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()}, strict=True)
        model.train()                                                   # train.py:84
        x, emb = synth.make_inputs(B, T, dims, iseed)
        xt, et = torch.from_numpy(x).requires_grad_(True), torch.from_numpy(emb).requires_grad_(True)
        mask = model(xt, et)                                            # train.py:94
        (mask * torch.from_numpy(gw_for(gseed, B, T, dims["num_freq"]))).sum().backward()      # stands in for train.py:108-110
        out = {}
        for k, p in model.named_parameters():
            g = p.grad.numpy()
            if g.size > SAMPLE_ABOVE:      # the five 5x5 conv weights (102 400 elements each): every 4th element + two float64 moments
                out["gradsample." + k] = g.reshape(-1)[::SAMPLE_STEP].copy()
                out["gradmoments." + k] = np.array([g.astype(np.float64).sum(), (g.astype(np.float64) ** 2).sum()])
            else:
                out["grad." + k] = g
        assert len([k for k in out if not k.startswith("gradmoments.")]) == 44
        out.update({"buf." + k: v.numpy() for k, v in model.state_dict().items() if "running" in k or "num_batches" in k})
        path = os.path.join(ROOT, "tests", "golden", f"train_{name}.npz")
        np.savez_compressed(
            path, model_name=model_name, flavour=flavour, wseed=wseed, iseed=iseed, gseed=gseed, B=B, T=T,
            dims=np.array([dims[k] for k in ("num_freq", "emb_dim", "lstm_dim", "fc1_dim", "fc2_dim")]),
            mask=mask.detach().numpy(), grad_x=xt.grad.numpy(), grad_emb=et.grad.numpy(), torch_version=torch.__version__, **out)
        print(name, "mask range", float(mask.min().detach()), float(mask.max().detach()), "max |grad|", max(float(np.abs(v).max()) for k, v in out.items() if k.startswith(("grad.", "gradsample."))),
              os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
