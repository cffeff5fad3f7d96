"""STFT front end / iSTFT back end on the device against oracle/audio_oracle.py (itself cross-validated against
torch.stft/istft and scipy in tests/test_audio_oracle.py)."""
import numpy as np
import pytest
import torch

from oracle import audio_oracle as ao
from voicesplit_b200 import synth
from voicesplit_b200.engine import MaskEngine

pytestmark = pytest.mark.gpu


def _signals(B, L, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(L) / 16000.0
    out = []
    for b in range(B):
        f1, f2 = 150 + 90 * b, 900 + 333 * b
        out.append(0.03 * np.sin(2 * np.pi * f1 * t) + 0.02 * np.sin(2 * np.pi * f2 * t + b) + 0.004 * rng.standard_normal(L))
    return np.stack(out).astype(np.float32)


@pytest.fixture(scope="module")
def engine():
    dims = synth.make_dims(601, 256, 400, 600)
    eng = MaskEngine(activation="mish", **dims)
    sd = synth.make_state_dict(dims, 3, "stress")
    eng.load_state_dict_tensors({k: torch.from_numpy(v).cuda() for k, v in sd.items() if "num_batches" not in k})
    eng.configure_audio()
    return eng


@pytest.mark.parametrize("B,L", [(1, 16000), (3, 48000), (2, 7777)])
def test_wav2spec_matches_oracle(engine, B, L):
    wav = _signals(B, L)
    spec, phasor = engine.wav2spec(torch.from_numpy(wav).cuda())
    spec, phasor = spec.cpu().numpy(), phasor.cpu().numpy()
    for b in range(B):
        S, ph = ao.wav2spec(wav[b])
        assert spec[b].shape == S.shape
        # normalised dB magnitude: 1e-2 of the [0,1] range = 1 dB, only bins near the -100 dB floor differ that much
        d = np.abs(spec[b] - S)
        assert d.max() < 2e-2 and d.mean() < 2e-4, (d.max(), d.mean())
        strong = S > 0.35                                     # phase is only meaningful where there is energy
        err = np.abs(phasor[b][..., 0] + 1j * phasor[b][..., 1] - np.exp(1j * ph))[strong]
        assert err.max() < 5e-3


@pytest.mark.parametrize("B,L", [(2, 16000), (1, 48000)])
def test_spec2wav_matches_oracle_and_round_trips(engine, B, L):
    wav = _signals(B, L, seed=4)
    wt = torch.from_numpy(wav).cuda()
    spec, phasor = engine.wav2spec(wt)
    back = engine.spec2wav(spec, phasor).cpu().numpy()
    for b in range(B):
        S, ph = ao.wav2spec(wav[b])
        ref = ao.spec2wav(S, ph)
        assert back[b].shape == ref.shape
        assert np.abs(back[b] - ref).max() < 2e-4
        assert np.abs(back[b] - wav[b][:len(ref)]).max() < 5e-4          # analysis -> synthesis is the identity up to the dB floor
    # masked half-amplitude spectrum gives half-amplitude audio (linearity of the back end in the amplitude domain)
    half = spec - (20 * np.log10(2.0)) / 100.0
    quiet = engine.spec2wav(half.clamp(0, 1), phasor).cpu().numpy()
    keep = (spec.cpu().numpy() > 0.2)
    assert np.abs(quiet - 0.5 * back).max() < 2e-3 * max(1.0, np.abs(back).max()) or keep.mean() < 0.5


def test_separate_end_to_end_shapes(engine):
    wav = torch.from_numpy(_signals(2, 48000, seed=7)).cuda()
    emb = torch.randn(2, 256, device="cuda")
    out = engine.separate(wav, emb)
    assert out.shape == (2, 160 * (1 + 48000 // 160 - 1)) and torch.isfinite(out).all()
    assert out.abs().max() <= wav.abs().max() * 1.5


def test_separation_matches_full_oracle_pipeline(engine):
    """BASELINE config 5 parity: waveform -> STFT -> mask -> mask*spec -> iSTFT (mixture phase) on the device against
    the same chain built only from oracles (audio_oracle + torch_port with the same weights).  Reported as the
    Si-SNR of our output w.r.t. the oracle's output (the reference's own quality measure)."""
    from oracle import torch_port
    dims = synth.make_dims(601, 256, 400, 600)
    sd = synth.make_state_dict(dims, 3, "stress")
    wav = _signals(1, 20000, seed=11)
    emb = np.random.default_rng(2).standard_normal((1, 256)).astype(np.float32)
    got = engine.separate(torch.from_numpy(wav).cuda(), torch.from_numpy(emb).cuda()).cpu().numpy()[0]
    S, ph = ao.wav2spec(wav[0])
    mask = torch_port.forward(sd, S[None].astype(np.float32), emb, "mish").numpy()[0]
    ref = ao.spec2wav(mask * S, ph)
    assert got.shape == ref.shape
    err = got - ref
    si_snr = 10 * np.log10((ref ** 2).sum() / max((err ** 2).sum(), 1e-30))
    print(f"device vs oracle pipeline: SNR {si_snr:.1f} dB, max |diff| {np.abs(err).max():.2e}")
    assert si_snr > 50.0


def test_separation_from_reference_audio_matches_full_oracle_pipeline(engine):
    """Config 5 from raw audio on both inputs: reference clip -> GE2E d-vector (device) and mixture -> STFT -> mask ->
    iSTFT, against the all-oracle chain (encoder_oracle + audio_oracle + torch_port) with the same weights."""
    from oracle import encoder_oracle as eo, torch_port
    from voicesplit_b200.speaker_encoder import SpeakerEncoder
    dims = synth.make_dims(601, 256, 400, 600)
    sd = synth.make_state_dict(dims, 3, "stress")
    esd = synth.make_encoder_state_dict(4, "default")
    enc = SpeakerEncoder(engine=engine)
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in esd.items()})
    enc = enc.cuda()
    mix = _signals(1, 20000, seed=12)
    ref_clip = synth.make_reference_audio(1, 24000, 13)
    dvec = enc.embed_wav(torch.from_numpy(ref_clip).cuda())
    got = engine.separate(torch.from_numpy(mix).cuda(), dvec).cpu().numpy()[0]
    d_ref = eo.speaker_encoder(esd, eo.get_mel(ref_clip[0])).astype(np.float32)[None]
    assert np.abs(dvec.cpu().numpy() - d_ref).max() < 3e-4
    S, ph = ao.wav2spec(mix[0])
    mask = torch_port.forward(sd, S[None].astype(np.float32), d_ref, "mish").numpy()[0]
    ref = ao.spec2wav(mask * S, ph)
    err = got - ref
    snr = 10 * np.log10((ref ** 2).sum() / max((err ** 2).sum(), 1e-30))
    print(f"device vs oracle pipeline (d-vector from audio): SNR {snr:.1f} dB")
    assert snr > 50.0
