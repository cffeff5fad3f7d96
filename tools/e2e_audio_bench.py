"""BASELINE config 5 shape: LibriSpeech-length mixtures (3 s @ 16 kHz) end to end on the device,
waveform -> STFT -> CNN+BiLSTM+FC mask -> mask * spectrogram -> iSTFT with the mixture phase -> waveform.
Prints one JSON line: utterances/s with the audio resident on the device, and with pinned host buffers."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voicesplit_b200 import synth  # noqa: E402
from voicesplit_b200.engine import MaskEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--precision", default="fp16x3")
    args = ap.parse_args()
    dims = synth.make_dims(601, 256, 400, 600)
    eng = MaskEngine(activation="mish", **dims)
    sd = synth.make_state_dict(dims, 0, "stress")
    eng.load_state_dict_tensors({k: torch.from_numpy(v).cuda() for k, v in sd.items() if v.dtype == np.float32})
    eng.configure_audio()
    B, L = args.batch, int(args.seconds * 16000)
    rng = np.random.default_rng(0)
    wav_h = torch.from_numpy((0.05 * rng.standard_normal((B, L))).astype(np.float32)).pin_memory()
    emb_h = torch.from_numpy(rng.standard_normal((B, 256)).astype(np.float32)).pin_memory()
    wav, emb = wav_h.cuda(), emb_h.cuda()
    out_h = torch.empty(B, 160 * (L // 160), dtype=torch.float32).pin_memory()

    def timed(fn, n):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    ms_dev = timed(lambda: eng.separate(wav, emb, precision=args.precision), args.steps)

    def host_step():
        out = eng.separate(wav_h.cuda(non_blocking=True), emb_h.cuda(non_blocking=True), precision=args.precision)
        out_h.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    ms_host = timed(host_step, args.steps)
    eng.set_profiling(True)
    spec, ph = eng.wav2spec(wav)
    t_stft = sum(ms for _, ms in eng.profile_read())
    _, masked = eng.forward(spec, emb, precision=args.precision, want_masked=True)
    t_model = sum(ms for _, ms in eng.profile_read())
    eng.spec2wav(masked, ph)
    t_istft = sum(ms for _, ms in eng.profile_read())
    print(json.dumps({"metric": "utterances/s, 3 s @ 16 kHz waveform -> separated waveform (STFT + mask model + iSTFT)",
                      "value": B / (ms_dev / 1e3), "e2e_host_buffers": B / (ms_host / 1e3), "batch": B, "samples": L, "frames": 1 + L // 160,
                      "precision": args.precision, "ms_per_step": ms_dev, "ms": {"stft": t_stft, "mask_model": t_model, "istft": t_istft}}))


if __name__ == "__main__":
    main()
