"""Deterministic synthetic weights and inputs for the VoiceSplit/VoiceFilter mask path.

There is no network (no checkpoints, no LibriSpeech), so every test, the smoke run and the
benchmark use weights generated here from a seed with numpy's PCG64 stream.  Two flavours,
following SURVEY.md section 8(d):

* ``default`` - the scale of PyTorch's default initialisation of the reference layers
  (/root/reference/models/voicesplit/model.py:15-64): U(-1/sqrt(fan_in), 1/sqrt(fan_in))
  for conv/linear/LSTM weights and biases, BatchNorm gamma=1 beta=0 mean=0 var=1.  With these
  the mask lives in about [0.47, 0.53], so a 1e-3 tolerance is almost vacuous.
* ``stress`` - randomised BatchNorm statistics and scaled-up weights so that the mask
  spans (0, 1) and operand rounding in the kernels is visible.

The state dict uses exactly the reference key names/shapes.
"""
from __future__ import annotations

import numpy as np

# (sequential index of the conv, C_in, C_out, kh, kw, dilation_t) - the eight conv layers of
# /root/reference/models/voicesplit/model.py:15-52 (indices are positions in nn.Sequential).
CONV_LAYERS = (
    (1, 1, 64, 1, 7, 1),
    (5, 64, 64, 7, 1, 1),
    (9, 64, 64, 5, 5, 1),
    (13, 64, 64, 5, 5, 2),
    (17, 64, 64, 5, 5, 4),
    (21, 64, 64, 5, 5, 8),
    (25, 64, 64, 5, 5, 16),
    (28, 64, 8, 1, 1, 1),
)
# BatchNorm index that follows each conv in the Sequential
BN_INDEX = {1: 2, 5: 6, 9: 10, 13: 14, 17: 18, 21: 22, 25: 26, 28: 29}


def make_dims(num_freq=601, emb_dim=256, lstm_dim=400, fc1_dim=600, fc2_dim=None):
    return dict(num_freq=int(num_freq), emb_dim=int(emb_dim), lstm_dim=int(lstm_dim),
                fc1_dim=int(fc1_dim), fc2_dim=int(num_freq if fc2_dim is None else fc2_dim))


def make_config_dict(dims, model_name="voicesplit"):
    """A config.json-shaped dict (/root/reference/config.json:1-98, only the keys the module reads)."""
    return {
        "model_name": model_name,
        "model": {"lstm_dim": dims["lstm_dim"], "fc1_dim": dims["fc1_dim"],
                  "fc2_dim": dims["fc2_dim"], "emb_dim": dims["emb_dim"]},
        "audio": {"backend": "voicefilter",
                  "voicefilter": {"num_freq": dims["num_freq"], "n_fft": 2 * (dims["num_freq"] - 1),
                                  "sample_rate": 16000, "hop_length": 160, "win_length": 400}},
    }


def _uniform(rng, shape, bound):
    return rng.uniform(-bound, bound, size=shape).astype(np.float32)


def make_state_dict(dims, seed=0, flavour="default"):
    """Return {key: np.ndarray} with the reference's state_dict keys, shapes and dtypes."""
    assert flavour in ("default", "stress")
    rng = np.random.Generator(np.random.PCG64(seed))
    stress = flavour == "stress"
    sd = {}
    for idx, cin, cout, kh, kw, _dil in CONV_LAYERS:
        bound = 1.0 / np.sqrt(cin * kh * kw)
        w = _uniform(rng, (cout, cin, kh, kw), bound)
        b = _uniform(rng, (cout,), bound)
        if stress:
            w *= 3.0
        sd[f"conv.{idx}.weight"] = w
        sd[f"conv.{idx}.bias"] = b
        bn = BN_INDEX[idx]
        if stress:
            sd[f"conv.{bn}.weight"] = rng.uniform(0.8, 1.6, size=(cout,)).astype(np.float32)
            sd[f"conv.{bn}.bias"] = (0.2 * rng.standard_normal(cout)).astype(np.float32)
            sd[f"conv.{bn}.running_mean"] = (0.1 * rng.standard_normal(cout)).astype(np.float32)
            sd[f"conv.{bn}.running_var"] = rng.uniform(0.5, 1.5, size=(cout,)).astype(np.float32)
        else:
            sd[f"conv.{bn}.weight"] = np.ones(cout, np.float32)
            sd[f"conv.{bn}.bias"] = np.zeros(cout, np.float32)
            sd[f"conv.{bn}.running_mean"] = np.zeros(cout, np.float32)
            sd[f"conv.{bn}.running_var"] = np.ones(cout, np.float32)
        sd[f"conv.{bn}.num_batches_tracked"] = np.array(0, np.int64)
    H = dims["lstm_dim"]
    I = 8 * dims["num_freq"] + dims["emb_dim"]
    bound = 1.0 / np.sqrt(H)
    for sfx in ("", "_reverse"):
        w_ih = _uniform(rng, (4 * H, I), bound)
        w_hh = _uniform(rng, (4 * H, H), bound)
        if stress:
            w_ih *= 4.0
            w_hh *= 4.0
        sd[f"lstm.weight_ih_l0{sfx}"] = w_ih
        sd[f"lstm.weight_hh_l0{sfx}"] = w_hh
        sd[f"lstm.bias_ih_l0{sfx}"] = _uniform(rng, (4 * H,), bound)
        sd[f"lstm.bias_hh_l0{sfx}"] = _uniform(rng, (4 * H,), bound)
    b1 = 1.0 / np.sqrt(2 * H)
    sd["fc1.weight"] = _uniform(rng, (dims["fc1_dim"], 2 * H), b1) * (4.0 if stress else 1.0)
    sd["fc1.bias"] = _uniform(rng, (dims["fc1_dim"],), b1)
    b2 = 1.0 / np.sqrt(dims["fc1_dim"])
    sd["fc2.weight"] = _uniform(rng, (dims["fc2_dim"], dims["fc1_dim"]), b2) * (12.0 if stress else 1.0)
    sd["fc2.bias"] = _uniform(rng, (dims["fc2_dim"],), b2)
    return sd


def make_inputs(B, T, dims, seed=1234, normalised_emb=False):
    """Spectrogram in [0,1] (the reference's clip-normalised dB range,
    /root/reference/utils/audio_processor.py:543-544) and a random d-vector."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = rng.uniform(0.0, 1.0, size=(B, T, dims["num_freq"])).astype(np.float32)
    emb = rng.standard_normal((B, dims["emb_dim"])).astype(np.float32)
    if normalised_emb:
        emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    return x, emb


def loss_inputs(n_fft, B, T, seed):
    """Seeded inputs of the training-loss chain (tests/golden/make_loss_golden.py): estimate slightly outside [0, 1]
    (exercises the clamp of torch_spec2wav), target in [0, 1], phase angles in (-pi, pi]; all [B, T, n_fft // 2 + 1]."""
    F = n_fft // 2 + 1
    rng = np.random.Generator(np.random.PCG64(seed))
    est = (rng.random((B, T, F)) * 1.2 - 0.1).astype(np.float32)
    tgt = rng.random((B, T, F)).astype(np.float32)
    phase = ((rng.random((B, T, F)) * 2 - 1) * np.pi).astype(np.float32)
    return est, tgt, phase


def make_encoder_state_dict(seed, flavour="default", num_mels=40, hidden=768, layers=3, emb_dim=256):
    """GE2E speaker-encoder parameters under the notebook's state_dict keys (lstm.*_l{0..2}, proj.linear_layer.*).
    "default": PyTorch's U(-1/sqrt(H), 1/sqrt(H)); "stress": 3x larger weights and biases (still non-chaotic), so that
    gate saturation and error accumulation over 80 steps x 3 layers are exercised."""
    rng = np.random.Generator(np.random.PCG64(seed))
    k = 1.0 / np.sqrt(hidden)
    s = 3.0 if flavour == "stress" else 1.0
    sd = {}
    for l in range(layers):
        d_in = num_mels if l == 0 else hidden
        sd[f"lstm.weight_ih_l{l}"] = (rng.uniform(-k, k, (4 * hidden, d_in)) * s).astype(np.float32)
        sd[f"lstm.weight_hh_l{l}"] = (rng.uniform(-k, k, (4 * hidden, hidden)) * s).astype(np.float32)
        sd[f"lstm.bias_ih_l{l}"] = (rng.uniform(-k, k, 4 * hidden) * s).astype(np.float32)
        sd[f"lstm.bias_hh_l{l}"] = (rng.uniform(-k, k, 4 * hidden) * s).astype(np.float32)
    sd["proj.linear_layer.weight"] = (rng.uniform(-k, k, (emb_dim, hidden)) * s).astype(np.float32)
    sd["proj.linear_layer.bias"] = (rng.uniform(-k, k, emb_dim) * s).astype(np.float32)
    return sd


def make_reference_audio(B, L, seed):
    """Speech-like test signals [B, L] at 16 kHz: a few harmonics with slow amplitude modulation plus noise."""
    rng = np.random.Generator(np.random.PCG64(seed))
    t = np.arange(L) / 16000.0
    out = np.zeros((B, L))
    for b in range(B):
        f0 = rng.uniform(90, 250)
        for h in range(1, 9):
            out[b] += rng.uniform(0.2, 1.0) / h * np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 6.28)) * (0.6 + 0.4 * np.sin(2 * np.pi * rng.uniform(1, 4) * t))
        out[b] = 0.05 * out[b] / np.abs(out[b]).max() + 0.002 * rng.standard_normal(L)
    return out.astype(np.float32)


def encoder_mel_inputs(seed, frames, num_mels=40):
    """log-mel-like encoder inputs, one [num_mels, T] array per entry of `frames` (values in get_mel's range, about -6 .. 1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return [(rng.standard_normal((num_mels, T)) * 1.2 - 2.5).astype(np.float32) for T in frames]
