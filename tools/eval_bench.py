"""Evaluation-driver benchmark: N test items of 3 s through voicesplit_b200.evaluate.validation
(mask -> iSTFT with the mixture phase -> Si-SNR -> BSS-Eval SDR, all on the device, batched) and, beside it, the time the
CPU oracle needs for the SDR of one item (the reference computes it with mir_eval on the CPU, one item at a time).

    python tools/eval_bench.py --items 128"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voicesplit_b200 import config as vconfig, evaluate, synth  # noqa: E402
from voicesplit_b200.audio import DeviceAudioProcessor  # noqa: E402


def main():
    ap_ = argparse.ArgumentParser()
    ap_.add_argument("--items", type=int, default=128)
    ap_.add_argument("--batch", type=int, default=64)
    args = ap_.parse_args()
    from models.voicesplit.model import VoiceSplit
    dims = synth.make_dims(601, 256, 400, 600)
    model = VoiceSplit(vconfig.AttrDict(synth.make_config_dict(dims)))
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_state_dict(dims, 3, "default").items()})
    model = model.cuda().eval()
    eng = model.engine()
    ap = DeviceAudioProcessor(eng, dict(n_fft=1200, hop_length=160, win_length=400))
    L = 48000
    clean = torch.from_numpy(synth.make_reference_audio(args.items, L, 1)).cuda()
    mixed = clean + 0.7 * torch.from_numpy(synth.make_reference_audio(args.items, L, 2)).cuda()
    ms, mph = eng.wav2spec(mixed)
    cs, _ = eng.wav2spec(clean)
    angle = torch.atan2(mph[..., 1], mph[..., 0])
    emb = torch.randn(args.items, 256, device="cuda") * 0.05
    loader = [[(emb[i], cs[i], ms[i], clean[i], mixed[i], angle[i], torch.tensor([L]))] for i in range(args.items)]
    evaluate.validation(None, ap, model, loader, None, 0, loss_name="si_snr", test=True, batch_size=args.batch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mean_loss, mean_sdr = evaluate.validation(None, ap, model, loader, None, 0, loss_name="si_snr", test=True, batch_size=args.batch)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    est = clean[: args.batch] * 0.8 + 0.1 * mixed[: args.batch]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.sdr(clean[: args.batch], est)
    e0.record()
    for _ in range(5):
        eng.sdr(clean[: args.batch], est)
    e1.record(); torch.cuda.synchronize()
    from oracle import sdr_oracle
    c0 = time.perf_counter()
    sdr_oracle.sdr(clean[0].cpu().numpy(), est[0].cpu().numpy())
    cpu_s = time.perf_counter() - c0
    print(json.dumps({"metric": "evaluated test items/s (mask + iSTFT + Si-SNR + BSS-Eval SDR, wall clock incl. Python)", "value": args.items / dt,
                      "items": args.items, "batch": args.batch, "mean_loss": mean_loss, "mean_sdr": mean_sdr,
                      "sdr_kernel_ms_per_batch": e0.elapsed_time(e1) / 5, "sdr_cpu_oracle_s_per_item": cpu_s}))


if __name__ == "__main__":
    main()
