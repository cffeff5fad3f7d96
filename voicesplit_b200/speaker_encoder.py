"""Drop-in for the GE2E speaker encoder the reference uses to produce its d-vectors
(notebooks/GE2E-Seungwonpark-ExtractSpeakerEmbedding-adaptado-para-openvoicefilter.py:52-85): same class names,
constructor arguments and state_dict keys (lstm.weight_ih_l{k}, lstm.weight_hh_l{k}, lstm.bias_*_l{k},
proj.linear_layer.weight / .bias), so `embedder.load_state_dict(torch.load("embedder.pt"))` works unchanged.
The arithmetic runs in the engine (vs_encoder_forward / vs_encoder_dvector: tcgen05 GEMMs + persistent recurrent
kernel); the nn.LSTM / nn.Linear members only hold the parameters.  Inference only (the reference never trains it)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .engine import MaskEngine


class LinearNorm(nn.Module):
    def __init__(self, lstm_hidden, emb_dim):
        super().__init__()
        self.linear_layer = nn.Linear(lstm_hidden, emb_dim)


class SpeakerEncoder(nn.Module):
    def __init__(self, num_mels=40, lstm_layers=3, lstm_hidden=768, window=80, stride=40, emb_dim=256, engine: MaskEngine | None = None,
                 sample_rate=16000):
        """engine: the MaskEngine to run on (e.g. `model.engine()`), with configure_audio() done; attach later with .attach(engine)."""
        super().__init__()
        self.lstm = nn.LSTM(num_mels, lstm_hidden, num_layers=lstm_layers, batch_first=True)
        self.proj = LinearNorm(lstm_hidden, emb_dim)
        self.num_mels, self.lstm_layers, self.lstm_hidden = num_mels, lstm_layers, lstm_hidden
        self.window, self.stride, self.emb_dim, self.sample_rate = window, stride, emb_dim, sample_rate
        self._engine = None
        self._packed_sig = None
        if engine is not None:
            self.attach(engine)

    def attach(self, engine: MaskEngine):
        if getattr(engine, "audio", None) is None:
            engine.configure_audio()
        engine.configure_encoder(self.num_mels, self.lstm_layers, self.lstm_hidden, self.emb_dim, self.window, self.stride, self.sample_rate)
        self._engine, self._packed_sig = engine, None
        return self

    def _sync(self):
        if self._engine is None:
            raise RuntimeError("SpeakerEncoder has no engine: pass engine= or call .attach(model.engine())")
        sig = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if sig != self._packed_sig:              # load_state_dict / .cuda() invalidate the packing
            self._engine.load_encoder_state_dict(self.state_dict())
            self._packed_sig = sig
        return self._engine

    @torch.no_grad()
    def forward(self, mel: torch.Tensor) -> torch.Tensor:
        """mel [num_mels, T] -> d-vector [emb_dim] (the notebook's call), or batched [B, num_mels, T] -> [B, emb_dim]."""
        if not mel.is_cuda:
            raise RuntimeError("voicesplit_b200 runs on sm_100a CUDA devices only: move the mel spectrogram to the GPU")
        single = mel.dim() == 2
        m = (mel[None] if single else mel).transpose(1, 2)          # frames as rows
        if m.shape[1] < self.window:
            raise ValueError("reference audio shorter than one encoder window")    # the notebook's except branch (:144-147)
        d = self._sync().encoder_forward(m)
        return d[0] if single else d

    @torch.no_grad()
    def embed_wav(self, wav: torch.Tensor) -> torch.Tensor:
        """wav [L] or [B, L] at 16 kHz -> d-vector(s): get_mel + forward without leaving the device (notebook :141-143)."""
        single = wav.dim() == 1
        d = self._sync().encoder_dvector(wav[None] if single else wav)
        return d[0] if single else d

    @torch.no_grad()
    def get_mel(self, wav: torch.Tensor) -> torch.Tensor:
        """ap.get_mel (utils/audio_processor.py:460-468): wav [L] -> [num_mels, T] (or batched [B, num_mels, T])."""
        single = wav.dim() == 1
        m = self._sync().encoder_mel(wav[None] if single else wav).transpose(1, 2)
        return m[0] if single else m
