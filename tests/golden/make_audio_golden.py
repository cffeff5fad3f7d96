"""Real-audio fixtures for BASELINE config 5 (SURVEY.md 8(d) C5 fallback): three (mixture, reference, clean) triples of the
reference's own demo set, cropped to 3 s @ 16 kHz, with the outputs of the all-oracle chain on them.

Run in the build container (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_audio_golden.py
Inputs:  /root/reference/datasets/LibriSpeech/test_demo.csv rows 0..2 (mixture = noise_utterance, reference = emb_utterance,
         clean = clean_utterance) and the GE2E checkpoint /root/reference/notebooks/embedder.pt (48 MB, cannot be committed).
Stored per clip (tests/golden/audio_demo_<i>.npz, ~250 KB each):
    mix, clean int16 [48000], ref int16 [<= 48000]          the cropped audio itself (the WAVs are int16 or float32 in [-1, 1))
    dvec              float32 [256]                         oracle/encoder_oracle.py on `ref` with the REAL embedder.pt weights
    spec_rows         float32 [16, 601], row_idx            every 20th frame of oracle/audio_oracle.wav2spec(mix)
    est               float16-rounded float32 [48000]       separated waveform of the oracle chain: wav2spec -> torch_port mask
                                                            (stress weights, seed 3, d-vector above) -> mask * spec -> spec2wav
    sdr, sisnr_loss   float64                               oracle SDR(clean, est) and the validation criterion(clean, est) (Q2 order)
The mask weights are the seeded "stress" flavour (no trained VoiceSplit checkpoint exists)."""
import csv
import os
import sys

import numpy as np
import scipy.io.wavfile
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import audio_oracle as ao, encoder_oracle as eo, loss_oracle, sdr_oracle, torch_port  # noqa: E402
from voicesplit_b200 import synth  # noqa: E402

REF = "/root/reference"
L = 48000


def load16k(path, start):
    sr, w = scipy.io.wavfile.read(os.path.join(REF, path))
    assert sr == 16000 and w.ndim == 1
    if w.dtype != np.int16:
        w = np.clip(np.round(w.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16)
    if len(w) < start + L:                               # short file: take its tail (the reference clips are ~2.5 s)
        start = max(0, len(w) - L)
    return w[start:start + L].copy()


def main():
    rows = list(csv.DictReader(open(os.path.join(REF, "datasets/LibriSpeech/test_demo.csv"))))[:3]
    esd = {k: v.numpy() for k, v in torch.load(os.path.join(REF, "notebooks/embedder.pt"), map_location="cpu").items()}
    dims = synth.make_dims(601, 256, 400, 600)
    sd = synth.make_state_dict(dims, 3, "stress")
    for i, r in enumerate(rows):
        start = 8000                                     # skip the leading half second (mostly silence in the demo clips)
        mix, ref, clean = (load16k(r[k], start) for k in ("noise_utterance", "emb_utterance", "clean_utterance"))
        f = lambda w: (w.astype(np.float32) / 32768.0)
        dvec = eo.speaker_encoder(esd, eo.get_mel(f(ref))).astype(np.float32)
        S, ph = ao.wav2spec(f(mix))
        mask = torch_port.forward(sd, S[None].astype(np.float32), dvec[None], "mish").numpy()[0]
        est = ao.spec2wav(mask * S, ph).astype(np.float32)
        assert len(mix) == L and len(clean) == L
        n = min(len(est), L)
        sdr = sdr_oracle.sdr(f(clean)[:n], est[:n])
        loss = float(loss_oracle.si_snr_c1(torch.from_numpy(f(clean)[:n])[None].double(), torch.from_numpy(est[:n])[None].double(),
                                           torch.tensor([n]))[0])
        idx = np.arange(0, S.shape[0], 20)
        out = os.path.join(ROOT, "tests", "golden", f"audio_demo_{i}.npz")
        np.savez_compressed(out, mix=mix, ref=ref, clean=clean, dvec=dvec, spec_rows=S[idx].astype(np.float32), row_idx=idx,
                            est=est.astype(np.float16).astype(np.float32), sdr=np.float64(sdr), sisnr_loss=np.float64(loss),
                            source=np.array([r["noise_utterance"], r["emb_utterance"], r["clean_utterance"]]))
        print(out, os.path.getsize(out), "sdr", round(float(sdr), 3), "loss", round(loss, 3), "mask range", float(mask.min()), float(mask.max()),
              "silence frac (S == 0)", float((S == 0).mean()))


if __name__ == "__main__":
    main()
