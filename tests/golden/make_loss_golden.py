"""Golden vectors for the training-loss chain, produced by the UNMODIFIED
reference code: openVoiceFilterAudioProcessor.torch_spec2wav (utils/audio_processor.py:498-509, including
its Q1 quirk) followed by SiSNR_With_Pit (utils/generic_utils.py:417-474), with autograd for the gradient
w.r.t. the estimated spectrogram - the tensors train.py:95-109 pushes through them.

    python tests/golden/make_loss_golden.py          (build container only; needs /root/reference)

One substitution is unavoidable: torchaudio.functional.istft, which the reference calls, was removed from
torchaudio (2.11 here).  It was upstreamed verbatim as torch.istft, so the shim below forwards to that (old
layout [..., F, T, 2] real pairs -> complex).  Everything else is the reference's own code.  Inputs are NOT
stored: they are regenerated from the seed by voicesplit_b200.synth.loss_inputs (numpy PCG64), only outputs are stored.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from voicesplit_b200.synth import loss_inputs  # noqa: E402

# name, (n_fft, hop, win), B, T, seed, lengths (None = full)
CASES = [
    ("loss_small", (64, 16, 32), 3, 40, 21, (624, 400, 97)),
    ("loss_small_full", (128, 32, 64), 2, 17, 22, None),
    ("loss_native", (1200, 160, 400), 2, 12, 23, (1760, 1501)),
]
MIN_DB, REF_DB = -100.0, 20.0


def load_reference_audio_processor():
    _, _, gu = ref_import.load()
    import torchaudio
    if not hasattr(torchaudio.functional, "istft"):
        def istft(stft_matrix, n_fft, hop_length=None, win_length=None, window=None, center=True, pad_mode="reflect",
                  normalized=False, onesided=True, length=None):
            return torch.istft(torch.view_as_complex(stft_matrix.contiguous()), n_fft, hop_length=hop_length, win_length=win_length,
                               window=window, center=center, normalized=normalized, onesided=onesided, length=length)
        torchaudio.functional.istft = istft
    lib = sys.modules["librosa"]
    lib.filters = types.SimpleNamespace(mel=lambda sr, n_fft, n_mels: np.zeros((n_mels, n_fft // 2 + 1), np.float32))
    for name in ("soundfile",):
        sys.modules.setdefault(name, types.ModuleType(name))
    ua = types.ModuleType("utils.audio"); ua.WaveGlowSTFT = object
    pkg = types.ModuleType("utils"); pkg.generic_utils = gu; pkg.audio = ua
    saved = {k: sys.modules.get(k) for k in ("utils", "utils.generic_utils", "utils.audio")}
    sys.modules.update({"utils": pkg, "utils.generic_utils": gu, "utils.audio": ua})
    try:
        spec = importlib.util.spec_from_file_location("_ref_audio_processor", os.path.join(ref_import.REF_ROOT, "utils", "audio_processor.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod, gu


def main():
    apm, gu = load_reference_audio_processor()
    crit = gu.SiSNR_With_Pit()
    for name, (n_fft, hop, win), B, T, seed, lengths in CASES:
        ap = apm.openVoiceFilterAudioProcessor(sample_rate=16000, n_fft=n_fft, num_freq=n_fft // 2 + 1, hop_length=hop, win_length=win,
                                               preemphasis=0.97, power=1.5, min_level_db=MIN_DB, ref_level_db=REF_DB, num_mels=40,
                                               griffin_lim_iters=60)
        est, tgt, phase = loss_inputs(n_fft, B, T, seed)
        L = hop * (T - 1)
        lens = np.array(lengths if lengths is not None else [L] * B, dtype=np.int64)
        e = torch.from_numpy(est).requires_grad_(True)
        # train.py:99-108 verbatim
        output = ap.torch_inv_spectrogram(e, torch.from_numpy(phase))
        target = ap.torch_inv_spectrogram(torch.from_numpy(tgt), torch.from_numpy(phase))
        wav_est = output.detach().numpy().copy()
        shape = list(target.shape)
        target = torch.reshape(target, [shape[0], 1] + shape[1:])
        output = torch.reshape(output, [shape[0], 1] + shape[1:])
        loss = crit(output, target, torch.from_numpy(lens))
        loss.backward()
        out = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
        np.savez_compressed(out, n_fft=n_fft, hop=hop, win=win, B=B, T=T, seed=seed, lengths=lens, min_db=MIN_DB, ref_db=REF_DB,
                            wav_est=wav_est.astype(np.float32), wav_tgt=target.detach().numpy()[:, 0].astype(np.float32),
                            loss=np.float32(loss.item()), grad_est=e.grad.numpy().astype(np.float32), torch_version=torch.__version__)
        print(name, "loss", loss.item(), "wav", wav_est.shape, "|grad| max", float(e.grad.abs().max()), os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
