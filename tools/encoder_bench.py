"""d-vector extraction benchmark: B reference clips of 3 s @ 16 kHz -> 256-d d-vectors,
the engine (vs_encoder_dvector) against the notebook's own recipe as stock torch ops on the same GPU
(torch.stft -> mel -> log10 -> unfold -> nn.LSTM (cuDNN) -> Linear -> normalise -> mean).

    python tools/encoder_bench.py --batch 128

Prints one JSON line.  Not the headline metric (bench.py is)."""
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
from voicesplit_b200.speaker_encoder import SpeakerEncoder  # noqa: E402


def mel_basis(sr=16000, n_fft=1200, n_mels=40):
    import torchaudio
    return torchaudio.functional.melscale_fbanks(n_freqs=n_fft // 2 + 1, f_min=0.0, f_max=sr / 2, n_mels=n_mels, sample_rate=sr, norm="slaney",
                                                 mel_scale="slaney").T


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--samples", type=int, default=48000)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    dims = synth.make_dims(601, 256, 400, 600)
    eng = MaskEngine(activation="mish", device=dev, **dims)
    eng.load_state_dict_tensors({k: torch.from_numpy(v).to(dev) for k, v in synth.make_state_dict(dims, 0, "default").items() if v.dtype == np.float32})
    eng.configure_audio()
    enc = SpeakerEncoder(engine=eng)
    sd = synth.make_encoder_state_dict(1, "stress")
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    enc = enc.to(dev)
    B, L = args.batch, args.samples
    wav = torch.from_numpy(synth.make_reference_audio(B, L, 5)).to(dev)

    def t_ms(fn, n):
        fn(); fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    ours = t_ms(lambda: enc.embed_wav(wav), args.steps)
    eng.set_profiling(True)
    enc.embed_wav(wav)
    prof = {}
    for name, ms in eng.profile_read():
        prof[name] = round(prof.get(name, 0.0) + ms, 3)
    eng.set_profiling(False)
    mel_ms = t_ms(lambda: enc.get_mel(wav), args.steps)
    # the notebook's recipe with stock torch ops on this GPU
    lstm = torch.nn.LSTM(40, 768, num_layers=3, batch_first=True).to(dev)
    lstm.load_state_dict({k[5:]: torch.from_numpy(v) for k, v in sd.items() if k.startswith("lstm.")})
    pw, pb = torch.from_numpy(sd["proj.linear_layer.weight"]).to(dev), torch.from_numpy(sd["proj.linear_layer.bias"]).to(dev)
    basis = mel_basis().to(dev)
    window = torch.hann_window(400, periodic=True, device=dev)

    @torch.no_grad()
    def eager():
        D = torch.stft(wav, 1200, 160, 400, window=window, center=True, pad_mode="reflect", return_complex=True)   # [B, 601, T]
        mel = torch.log10(basis @ (D.abs() ** 2) + 1e-6)                                                          # [B, 40, T]
        wins = mel.unfold(2, 80, 40).permute(0, 2, 3, 1).reshape(-1, 80, 40)                                      # [B*T', 80, 40]
        x = lstm(wins)[0][:, -1, :] @ pw.T + pb
        x = x / torch.norm(x, p=2, dim=1, keepdim=True)
        return x.view(B, -1, 256).mean(1)
    res = {}
    for name, tf32 in (("fp32", False), ("tf32", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        res[name] = t_ms(eager, args.steps)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    diff = float((enc.embed_wav(wav) - eager()).abs().max())
    print(json.dumps({"metric": "d-vectors/s (3 s reference clips -> 256-d GE2E embedding)", "value": B / (ours / 1e3), "unit": "utterances/s",
                      "batch": B, "samples": L, "ms_per_batch": round(ours, 3), "mel_front_end_ms": round(mel_ms, 3), "kernel_ms": prof,
                      "stock_torch_same_gpu": {"fp32_utt_per_s": B / (res["fp32"] / 1e3), "tf32_utt_per_s": B / (res["tf32"] / 1e3),
                                               "what": "torch.stft + matmul + nn.LSTM (cuDNN) + Linear, eager"},
                      "max_abs_diff_vs_torch_fp32": diff, "arithmetic": "fp16x3 tcgen05 (mel GEMM bf16x3), fp32 accumulate and cell state"}))


if __name__ == "__main__":
    main()
