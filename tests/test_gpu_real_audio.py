"""BASELINE config 5 on REAL audio: the reference's own demo triples (datasets/LibriSpeech/test_demo.csv, cropped to 3 s) through
the device chain  wav -> STFT -> mask (stress weights: the mask spans 0..1, so a wrong mask kernel cannot hide) -> mask * spec
-> iSTFT with the mixture phase -> SDR / Si-SNR,  against the all-oracle chain stored by tests/golden/make_audio_golden.py.
The d-vector in the fixture comes from the reference's REAL GE2E checkpoint (notebooks/embedder.pt) through the encoder oracle."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import audio_oracle as ao, encoder_oracle as eo
from voicesplit_b200 import synth
from voicesplit_b200.engine import MaskEngine

pytestmark = pytest.mark.gpu
CLIPS = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "audio_demo_*.npz")))


def _f(w):
    return w.astype(np.float32) / 32768.0


@pytest.fixture(scope="module")
def engine():
    dims = synth.make_dims(601, 256, 400, 600)
    eng = MaskEngine(activation="mish", **dims)
    eng.load_state_dict_tensors({k: torch.from_numpy(v).cuda() for k, v in synth.make_state_dict(dims, 3, "stress").items() if "num_batches" not in k})
    eng.configure_audio()
    return eng


def test_fixtures_present():
    assert len(CLIPS) == 3


@pytest.mark.parametrize("path", CLIPS, ids=lambda p: os.path.basename(p)[:-4])
@pytest.mark.parametrize("precision", ["fp16x3", "fp16_f8c"])
def test_real_clip_chain_matches_oracle_chain(engine, path, precision):
    z = np.load(path)
    mix, clean = _f(z["mix"]), _f(z["clean"])
    wav = torch.from_numpy(mix)[None].cuda()
    spec, phasor = engine.wav2spec(wav)
    rows = spec[0].cpu().numpy()[z["row_idx"]]
    d = np.abs(rows - z["spec_rows"])
    # int16 speech: silent stretches sit on the -100 dB floor, where 1e-2 of the [0,1] range is 1 dB of a value the clip removes
    assert d.max() < 2e-2 and d.mean() < 2e-4, (d.max(), d.mean())
    emb = torch.from_numpy(z["dvec"])[None].cuda()                     # d-vector of the REAL embedder.pt (oracle-computed)
    _, masked = engine.forward(spec, emb, precision=precision, want_masked=True)
    est = engine.spec2wav(masked, phasor)[0].cpu().numpy()
    ref = z["est"][:len(est)]
    err = est - ref
    snr = 10 * np.log10((ref.astype(np.float64) ** 2).sum() / max((err.astype(np.float64) ** 2).sum(), 1e-30))
    print(f"{os.path.basename(path)} {precision}: device vs oracle chain {snr:.1f} dB")
    assert snr > (45.0 if precision == "fp16x3" else 35.0)            # the golden waveform is stored fp16-rounded (~ 66 dB floor)
    n = len(est)
    sdr = engine.sdr(torch.from_numpy(clean[:n])[None].cuda(), torch.from_numpy(est)[None].cuda()).cpu().numpy()[0]
    loss, _ = engine.sisnr_wav(torch.from_numpy(clean[:n])[None].cuda(), torch.from_numpy(est)[None].cuda(), torch.tensor([n]))   # Q2 order
    assert abs(sdr - float(z["sdr"])) < (2e-2 if precision == "fp16x3" else 5e-2), (sdr, float(z["sdr"]))
    assert abs(float(loss) - float(z["sisnr_loss"])) < (2e-2 if precision == "fp16x3" else 5e-2)


@pytest.mark.parametrize("path", CLIPS[:2], ids=lambda p: os.path.basename(p)[:-4])
def test_device_encoder_on_real_reference_clip(engine, path):
    """The real reference clip (2.5 s of int16 speech) through the device mel front end and GE2E recurrence with seeded
    weights, against the encoder oracle on the same audio (the real checkpoint itself is checked by tools/real_embedder_check.py
    in a one-off GPU run: profiles/r02_real_embedder_check.txt)."""
    from voicesplit_b200.speaker_encoder import SpeakerEncoder
    z = np.load(path)
    ref = _f(z["ref"])
    esd = synth.make_encoder_state_dict(4, "stress")
    enc = SpeakerEncoder(engine=engine)
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in esd.items()})
    enc = enc.cuda()
    got = enc.embed_wav(torch.from_numpy(ref)[None].cuda()).cpu().numpy()[0]
    want = eo.speaker_encoder(esd, eo.get_mel(ref))
    assert np.abs(got - want).max() < 5e-4, np.abs(got - want).max()
