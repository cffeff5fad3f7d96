"""Device-side counterpart of the reference's configured audio backend, `openVoiceFilterAudioProcessor`
(utils/audio_processor.py:440-567), for the two methods that sit either side of the mask model:
wav2spec (STFT -> dB -> clip-normalise, plus the mixture phase) and spec2wav / inv_spectrogram with
that phase.  Same method names and argument meaning, but batched torch CUDA tensors in and out, and
the phase is carried as the unit phasor D/|D| (cos, sin) instead of an angle.  All arithmetic runs in
the engine (vs_wav2spec / vs_spec2wav: tcgen05 GEMMs + overlap-add); see DESIGN.md section 8."""
from __future__ import annotations

import torch

from .engine import MaskEngine


class DeviceAudioProcessor:
    def __init__(self, engine: MaskEngine, audio_config):
        """audio_config: the `audio[backend]` section of config.json (reference config.json:84-96)."""
        self.engine = engine
        self.sample_rate = audio_config.get("sample_rate", 16000)
        self.n_fft, self.hop_length, self.win_length = audio_config["n_fft"], audio_config["hop_length"], audio_config["win_length"]
        self.min_level_db = audio_config.get("min_level_db", -100.0)
        self.ref_level_db = audio_config.get("ref_level_db", 20.0)
        engine.configure_audio(self.n_fft, self.hop_length, self.win_length, self.min_level_db, self.ref_level_db)

    def wav2spec(self, y: torch.Tensor):
        """y [B, L] (or [L]) -> (S [B, T, F] in [0, 1], phasor [B, T, F, 2])."""
        single = y.dim() == 1
        S, ph = self.engine.wav2spec(y[None] if single else y)
        return (S[0], ph[0]) if single else (S, ph)

    get_spec_from_audio = wav2spec

    def spec2wav(self, spectrogram, phase=None):
        """(masked) spectrogram [B, T, F] + phasor from wav2spec -> waveform [B, hop * (T - 1)].
        The reference's drivers call this per item with NUMPY arrays and the phase ANGLE (validation(),
        utils/generic_utils.py:499-504: `ap.inv_spectrogram(est_mag, phase=mixed_phase)`): a numpy spectrogram [T, F] with a
        numpy angle array is accepted too and returns a numpy waveform, like openVoiceFilterAudioProcessor.spec2wav."""
        if not isinstance(spectrogram, torch.Tensor):
            import numpy as np
            if phase is None:
                raise ValueError("the device back end reconstructs with the mixture phase (Griffin-Lim is not provided)")
            S = torch.as_tensor(np.asarray(spectrogram, dtype=np.float32), device=self.engine.device)
            ang = torch.as_tensor(np.asarray(phase, dtype=np.float32), device=self.engine.device)
            return self.spec2wav(S, torch.stack((torch.cos(ang), torch.sin(ang)), dim=-1)).cpu().numpy()
        single = spectrogram.dim() == 2
        w = self.engine.spec2wav(spectrogram[None] if single else spectrogram, phase[None] if single else phase)
        return w[0] if single else w

    inv_spectrogram = spec2wav

    def configure_training(self, phase_mode="q1"):
        """Prepare torch_spec2wav / torch_inv_spectrogram (the differentiable iSTFT of the Si-SNR training loss)."""
        self.engine.configure_loss(self.n_fft, self.hop_length, self.win_length, self.min_level_db, self.ref_level_db, phase_mode)

    def torch_spec2wav(self, spectrogram: torch.Tensor, phase: torch.Tensor):
        """Reference torch_spec2wav (utils/audio_processor.py:498-509): spectrogram [B, T, F] + phase ANGLE [B, T, F] ->
        waveform [B, hop (T - 1)], differentiable w.r.t. the spectrogram.  Call configure_training() first."""
        from .losses import spec2wav_autograd
        return spec2wav_autograd(self.engine, spectrogram, phase)

    torch_inv_spectrogram = torch_spec2wav
