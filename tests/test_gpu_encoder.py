"""GE2E speaker encoder on the device (vs_encoder_mel / _forward / _dvector) against
 (1) golden d-vectors from the notebook's unmodified SpeakerEncoder (tests/golden/encoder_*.npz),
 (2) the float64 oracle (oracle/encoder_oracle.py) for the mel front end and the wav -> d-vector chain."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import encoder_oracle as eo
from voicesplit_b200 import synth
from voicesplit_b200.engine import MaskEngine
from voicesplit_b200.speaker_encoder import SpeakerEncoder

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "encoder_*.npz")))


@pytest.fixture(scope="module")
def engine():
    dims = synth.make_dims(601, 16, 24, 40)          # only num_freq (= n_fft / 2 + 1 of the STFT) matters to the encoder
    eng = MaskEngine(activation="mish", **dims)
    sd = synth.make_state_dict(dims, 3, "default")
    eng.load_state_dict_tensors({k: torch.from_numpy(v).cuda() for k, v in sd.items() if "num_batches" not in k})
    eng.configure_audio()
    return eng


def _encoder(engine, seed, flavour):
    enc = SpeakerEncoder(engine=engine).cuda()
    sd = synth.make_encoder_state_dict(seed, flavour)
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    return enc.cuda(), sd


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_dvector_matches_reference_golden(engine, path):
    g = np.load(path)
    enc, _ = _encoder(engine, int(g["wseed"]), str(g["flavour"]))
    mels = synth.encoder_mel_inputs(int(g["iseed"]), [int(t) for t in g["frames"]])
    for mel, want in zip(mels, g["dvec"]):
        got = enc(torch.from_numpy(mel).cuda()).cpu().numpy()
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 2e-4, np.abs(got - want).max()        # entries ~0.06, |dvec| <= 1


def test_batched_forward_equals_single(engine):
    enc, sd = _encoder(engine, 5, "stress")
    mels = synth.encoder_mel_inputs(6, [200] * 5)
    batch = enc(torch.from_numpy(np.stack(mels)).cuda()).cpu().numpy()
    for i, mel in enumerate(mels):
        ref = eo.speaker_encoder(sd, mel)
        assert np.abs(batch[i] - ref).max() <= 2e-4
    # more sequences than one 128-row group: 40 utterances x 4 windows
    many = synth.encoder_mel_inputs(7, [200] * 40)
    out = enc(torch.from_numpy(np.stack(many)).cuda()).cpu().numpy()
    for i in (0, 17, 39):
        assert np.abs(out[i] - eo.speaker_encoder(sd, many[i])).max() <= 2e-4


@pytest.mark.parametrize("B,L", [(2, 16000), (3, 48000), (1, 20001)])
def test_mel_front_end_matches_oracle(engine, B, L):
    enc, _ = _encoder(engine, 1, "default")
    wav = synth.make_reference_audio(B, L, 10 + B)
    mel = enc.get_mel(torch.from_numpy(wav).cuda()).cpu().numpy()
    for b in range(B):
        ref = eo.get_mel(wav[b])
        assert mel[b].shape == ref.shape
        d = np.abs(mel[b] - ref)
        assert d.max() <= 2e-3 and d.mean() <= 1e-4, (d.max(), d.mean())       # log10 units; the max sits at the 1e-6 floor


def test_wav_to_dvector_matches_oracle_chain(engine):
    enc, sd = _encoder(engine, 9, "default")
    wav = synth.make_reference_audio(3, 48000, 77)
    got = enc.embed_wav(torch.from_numpy(wav).cuda()).cpu().numpy()
    for b in range(3):
        ref = eo.speaker_encoder(sd, eo.get_mel(wav[b]))
        assert np.abs(got[b] - ref).max() <= 3e-4
        assert abs(np.linalg.norm(got[b]) - np.linalg.norm(ref)) <= 1e-3


def test_errors_are_loud(engine):
    enc, _ = _encoder(engine, 2, "default")
    with pytest.raises(ValueError):
        enc(torch.zeros(40, 79, device="cuda"))                 # shorter than one window
    with pytest.raises(RuntimeError):
        enc(torch.zeros(40, 100))                               # CPU tensor
    with pytest.raises(RuntimeError):
        SpeakerEncoder().cuda()(torch.zeros(40, 100, device="cuda"))   # no engine attached
