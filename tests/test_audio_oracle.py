"""oracle/audio_oracle.py (numpy restatement of the librosa calls the reference makes) cross-validated against
torch.stft / torch.istft and scipy.signal, which implement the same definition independently."""
import numpy as np
import torch

from oracle import audio_oracle as ao


def _signal(n=16000 * 3 // 4, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    return (0.4 * np.sin(2 * np.pi * 220 * t) + 0.2 * np.sin(2 * np.pi * 1333 * t + 1.0) + 0.05 * rng.standard_normal(n)).astype(np.float32)


def test_stft_matches_torch_and_scipy():
    y = _signal()
    D = ao.stft(y)
    win = torch.from_numpy(ao.hann_periodic(400)).float()
    Dt = torch.stft(torch.from_numpy(y), n_fft=1200, hop_length=160, win_length=400, window=win, center=True, pad_mode="reflect",
                    return_complex=True).numpy()
    assert D.shape == Dt.shape == (601, 1 + len(y) // 160)
    assert np.abs(D - Dt).max() < 2e-4 * np.abs(D).max()
    from scipy import signal
    yp = np.pad(y.astype(np.float64), 600, mode="reflect")
    _, _, Ds = signal.stft(yp, window=ao.padded_window(1200, 400), nperseg=1200, noverlap=1200 - 160, boundary=None, padded=False)
    Ds = Ds * ao.padded_window(1200, 400).sum()          # scipy normalises by the window sum
    assert np.abs(D - Ds).max() < 1e-8 * np.abs(D).max()


def test_istft_matches_torch_and_round_trips():
    y = _signal(seed=3)
    D = ao.stft(y)
    back = ao.istft(D)
    n = 160 * (D.shape[1] - 1)
    assert len(back) == n
    assert np.abs(back - y[:n]).max() < 1e-9          # perfect reconstruction where the window sum is non-zero
    win = torch.from_numpy(ao.hann_periodic(400)).float()
    bt = torch.istft(torch.from_numpy(D.astype(np.complex64)), n_fft=1200, hop_length=160, win_length=400, window=win, center=True).numpy()
    assert np.abs(bt - back[:len(bt)]).max() < 1e-5


def test_wav2spec_spec2wav_pipeline():
    y = 0.05 * _signal(seed=5)      # keep the STFT peaks under the +20 dB reference level (normalize clips above it)
    S, ph = ao.wav2spec(y)
    assert S.shape == ph.shape and S.shape[1] == 601 and S.min() >= 0.0 and S.max() <= 1.0
    back = ao.spec2wav(S, ph)
    # normalisation clips below -100 dB relative to the reference level: reconstruction is exact up to that floor
    assert np.abs(back - y[:len(back)]).max() < 1e-4


def test_real_audio_fixtures_are_reproducible_from_the_oracles():
    """tests/golden/audio_demo_*.npz (real clips of the reference's demo set): the stored spectrogram rows are what the audio
    oracle computes from the stored int16 mixture - fixture and oracle stay in step."""
    import glob
    import os
    paths = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "audio_demo_*.npz")))
    assert len(paths) == 3
    for p in paths:
        z = np.load(p)
        assert z["mix"].dtype == np.int16 and len(z["mix"]) == 48000 and len(z["ref"]) >= 16000
        S, _ = ao.wav2spec(z["mix"].astype(np.float32) / 32768.0)
        assert S.shape == (301, 601)
        assert np.abs(S[z["row_idx"]] - z["spec_rows"]).max() < 1e-5
        assert float(np.abs(z["dvec"]).max()) <= 1.0 and abs(float(np.linalg.norm(z["dvec"])) - 1.0) < 0.2   # mean of unit vectors
