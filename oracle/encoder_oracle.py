"""TEST INFRASTRUCTURE - CPU restatement (float64 numpy) of the d-vector producer:

  openVoiceFilterAudioProcessor.get_mel      /root/reference/utils/audio_processor.py:456-468
  SpeakerEncoder / LinearNorm (GE2E)         /root/reference/notebooks/GE2E-Seungwonpark-ExtractSpeakerEmbedding-adaptado-para-openvoicefilter.py:63-85
  as driven by the extraction loop           same file :141-143   (mel = ap.get_mel(wav); emb = embedder(mel))

get_mel's arithmetic lives in librosa (requirements.txt:3, unpinned, ABSENT here): librosa.core.stft (restated in
oracle/audio_oracle.py) and librosa.filters.mel, whose documented default algorithm (Slaney mel scale: linear below
1 kHz, logarithmic above; triangular filters; 'slaney' area normalisation 2 / (f_{m+2} - f_m)) is restated in
mel_filterbank().  PARITY UNPINNED against librosa itself; tests/test_encoder_oracle.py cross-validates the filterbank
against an independent implementation of the same definition (torchaudio.functional.melscale_fbanks).
The encoder is stock torch.nn (nn.LSTM, nn.Linear): speaker_encoder() restates it and IS pinned, by golden vectors
produced by the notebook's own unmodified classes (tests/golden/make_encoder_golden.py).
Only tests/, smoke() and bench.py's baseline legs may import this module."""
import numpy as np

from oracle import audio_oracle


def hz_to_mel(f):
    f = np.asarray(f, np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, min_log_hz) / min_log_hz) / logstep, mels)


def mel_to_hz(m):
    m = np.asarray(m, np.float64)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr=16000, n_fft=1200, n_mels=40):
    """librosa.filters.mel(sr, n_fft, n_mels) with its defaults (fmin 0, fmax sr/2, htk False, slaney norm) -> [n_mels, 1 + n_fft/2]."""
    fft_freqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fft_freqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    return w * (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]


def get_mel(y, sr=16000, n_fft=1200, hop_length=160, win_length=400, n_mels=40):
    """audio_processor.py:460-468: |STFT|^2 -> mel basis -> log10(. + 1e-6); returns [n_mels, T]."""
    D = audio_oracle.stft(y, n_fft, hop_length, win_length)
    return np.log10(mel_filterbank(sr, n_fft, n_mels) @ (np.abs(D) ** 2) + 1e-6)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def lstm_stack(sd, x, layers):
    """nn.LSTM(batch_first=True, unidirectional): x [N, T, D] -> top-layer outputs [N, T, H]; gate order i, f, g, o."""
    for l in range(layers):
        wih, whh = np.asarray(sd[f"lstm.weight_ih_l{l}"], np.float64), np.asarray(sd[f"lstm.weight_hh_l{l}"], np.float64)
        b = np.asarray(sd[f"lstm.bias_ih_l{l}"], np.float64) + np.asarray(sd[f"lstm.bias_hh_l{l}"], np.float64)
        N, T, _ = x.shape
        H = whh.shape[1]
        h, c = np.zeros((N, H)), np.zeros((N, H))
        out = np.zeros((N, T, H))
        gx = x @ wih.T + b
        for t in range(T):
            g = gx[:, t] + h @ whh.T
            i, f, gg, o = _sigmoid(g[:, :H]), _sigmoid(g[:, H:2 * H]), np.tanh(g[:, 2 * H:3 * H]), _sigmoid(g[:, 3 * H:])
            c = f * c + i * gg
            h = o * np.tanh(c)
            out[:, t] = h
        x = out
    return x


def speaker_encoder(sd, mel, window=80, stride=40, layers=3):
    """SpeakerEncoder.forward (notebook :75-85): mel [n_mels, T] -> d-vector [emb_dim]."""
    mel = np.asarray(mel, np.float64)
    n_win = (mel.shape[1] - window) // stride + 1
    if n_win < 1:
        raise ValueError("reference audio shorter than one window")       # the notebook's except branch (:144-147)
    wins = np.stack([mel[:, i * stride:i * stride + window].T for i in range(n_win)])   # unfold + permute: [T', window, n_mels]
    last = lstm_stack(sd, wins, layers)[:, -1, :]
    e = last @ np.asarray(sd["proj.linear_layer.weight"], np.float64).T + np.asarray(sd["proj.linear_layer.bias"], np.float64)
    e = e / np.linalg.norm(e, axis=1, keepdims=True)
    return e.sum(0) / e.shape[0]
