"""TEST INFRASTRUCTURE - numpy restatement of the reference's audio front/back end
(utils/audio_processor.py:469-496,537-547: wav2spec, spec2wav with the mixture phase, amp_to_db,
db_to_amp, normalize, denormalize).

The STFT/iSTFT arithmetic itself lives in a third-party dependency that is ABSENT here: librosa
(requirements.txt:3, unpinned).  `stft` / `istft` below restate librosa's documented algorithm for the
arguments the reference passes (n_fft, hop_length, win_length; defaults window='hann' (periodic),
center=True, pad_mode='reflect'; istft: same window, window-sum-square normalisation, centre trimmed).
PARITY UNPINNED against librosa itself (it cannot be imported); instead tests/test_audio_oracle.py
cross-validates this file against two independent implementations of the same definition,
torch.stft/torch.istft and scipy.signal, and round-trips it.
Only tests/, smoke() and bench tools may import this module."""
import numpy as np


def hann_periodic(n):
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)).astype(np.float64)


def padded_window(n_fft, win_length):
    w = np.zeros(n_fft)
    lp = (n_fft - win_length) // 2
    w[lp:lp + win_length] = hann_periodic(win_length)
    return w


def stft(y, n_fft=1200, hop_length=160, win_length=400):
    """Complex STFT [1 + n_fft/2, T], T = 1 + len(y) // hop (librosa.stft, center=True, reflect padding)."""
    y = np.asarray(y, np.float64)
    yp = np.pad(y, n_fft // 2, mode="reflect")
    T = 1 + (len(yp) - n_fft) // hop_length
    w = padded_window(n_fft, win_length)
    frames = np.stack([yp[t * hop_length:t * hop_length + n_fft] * w for t in range(T)], axis=1)
    return np.fft.rfft(frames, axis=0)


def istft(D, hop_length=160, win_length=400):
    """Inverse of `stft` (librosa.istft): overlap-add of windowed irfft frames divided by the window sum-square."""
    n_fft = 2 * (D.shape[0] - 1)
    T = D.shape[1]
    w = padded_window(n_fft, win_length)
    n = n_fft + hop_length * (T - 1)
    y = np.zeros(n)
    wss = np.zeros(n)
    frames = np.fft.irfft(D, n=n_fft, axis=0)
    for t in range(T):
        y[t * hop_length:t * hop_length + n_fft] += w * frames[:, t]
        wss[t * hop_length:t * hop_length + n_fft] += w * w
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2:n - n_fft // 2]


# ---- the reference's own scalar maps (utils/audio_processor.py:537-547) -----------------------------
def amp_to_db(x):
    return 20.0 * np.log10(np.maximum(1e-5, x))


def db_to_amp(x):
    return np.power(10.0, x * 0.05)


def normalize(S, min_level_db=-100.0):
    return np.clip(S / -min_level_db, -1.0, 0.0) + 1.0


def denormalize(S, min_level_db=-100.0):
    return (np.clip(S, 0.0, 1.0) - 1.0) * -min_level_db


def wav2spec(y, n_fft=1200, hop_length=160, win_length=400, min_level_db=-100.0, ref_level_db=20.0):
    """audio_processor.py:469-476 -> (S [T, F] in [0,1], phase [T, F])."""
    D = stft(y, n_fft, hop_length, win_length)
    S = normalize(amp_to_db(np.abs(D)) - ref_level_db, min_level_db)
    return S.T, np.angle(D).T


def spec2wav(spectrogram, phase, hop_length=160, win_length=400, min_level_db=-100.0, ref_level_db=20.0):
    """audio_processor.py:483-491 (the mixture-phase branch) + istft_phase :478-482."""
    S = db_to_amp(denormalize(spectrogram.T, min_level_db) + ref_level_db)
    return istft(S * np.exp(1j * phase.T), hop_length, win_length)
