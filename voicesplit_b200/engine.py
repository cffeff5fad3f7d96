"""Host-side wrapper of the C-ABI engine: owns the handle, keeps the packed parameters in sync
with a state dict of torch CUDA tensors, and provides the workspace from torch's caching
allocator.  PyTorch is used for device memory and streams only; all arithmetic of the mask path
runs inside libvoicesplit_sm100.so."""
from __future__ import annotations

import ctypes

import torch

from . import _cabi

CONV_IDX = (1, 5, 9, 13, 17, 21, 25, 28)   # nn.Sequential positions (reference model.py:15-52)
BN_IDX = (2, 6, 10, 14, 18, 22, 26, 29)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class _DeviceArray:
    """Zero-copy view of raw device memory handed to a callback (CUDA array interface v2)."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _wrap_device_array(ptr, count, typestr, device):
    return torch.as_tensor(_DeviceArray(ptr, count, typestr), device=device)


class MaskEngine:
    def __init__(self, num_freq, emb_dim, lstm_dim, fc1_dim, fc2_dim, activation="mish", device=None):
        self.lib = _cabi.load()
        if not torch.cuda.is_available():
            raise RuntimeError("voicesplit_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dims = dict(num_freq=num_freq, emb_dim=emb_dim, lstm_dim=lstm_dim, fc1_dim=fc1_dim, fc2_dim=fc2_dim)
        self.activation = activation
        d = _cabi.VsDims(num_freq, emb_dim, lstm_dim, fc1_dim, fc2_dim,
                         _cabi.ACT_MISH if activation == "mish" else _cabi.ACT_RELU)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.vs_engine_create(ctypes.byref(d), ctypes.byref(h)), "vs_engine_create")
        self.handle = h
        self._ws = None
        self._keep = None

    def close(self):
        if getattr(self, "handle", None):
            self.lib.vs_engine_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters ---------------------------------------------------------------------------
    def load_state_dict_tensors(self, sd):
        """sd: {reference state_dict key: fp32 CUDA tensor}.  Folds BN, repacks, uploads."""
        def g(k):
            t = sd[k]
            if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                t = t.detach().to(self.device, torch.float32).contiguous()
            return t
        keep = []
        p = _cabi.VsParams()
        for l in range(8):
            for field, key in (("conv_w", f"conv.{CONV_IDX[l]}.weight"), ("conv_b", f"conv.{CONV_IDX[l]}.bias"),
                               ("bn_gamma", f"conv.{BN_IDX[l]}.weight"), ("bn_beta", f"conv.{BN_IDX[l]}.bias"),
                               ("bn_mean", f"conv.{BN_IDX[l]}.running_mean"), ("bn_var", f"conv.{BN_IDX[l]}.running_var")):
                t = g(key); keep.append(t)
                getattr(p, field)[l] = t.data_ptr()
        for d, sfx in enumerate(("", "_reverse")):
            for field, key in (("w_ih", "weight_ih"), ("w_hh", "weight_hh"), ("b_ih", "bias_ih"), ("b_hh", "bias_hh")):
                t = g(f"lstm.{key}_l0{sfx}"); keep.append(t)
                getattr(p, field)[d] = t.data_ptr()
        for field, key in (("fc1_w", "fc1.weight"), ("fc1_b", "fc1.bias"), ("fc2_w", "fc2.weight"), ("fc2_b", "fc2.bias")):
            t = g(key); keep.append(t)
            setattr(p, field, t.data_ptr())
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.lib.vs_engine_load_params(self.handle, ctypes.byref(p), ctypes.c_void_p(st)),
                        "vs_engine_load_params")
        # the packing kernels read the source tensors asynchronously: keep them (and any staging copies) referenced
        # until the next load instead of synchronising the host on every optimizer step
        self._keep = keep

    # ---- workspace ----------------------------------------------------------------------------
    def _workspace(self, B, T, prec):
        need = int(self.lib.vs_workspace_bytes(self.handle, B, T, prec))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws, need

    @staticmethod
    def _prec(precision):
        return _cabi.PRECISIONS[precision] if isinstance(precision, str) else int(precision)

    def _check_inputs(self, x, emb):
        if x.dim() != 3 or x.shape[2] != self.dims["num_freq"]:
            raise ValueError(f"x must be [B, T, {self.dims['num_freq']}], got {tuple(x.shape)}")
        if emb.dim() != 2 or emb.shape[0] != x.shape[0] or emb.shape[1] != self.dims["emb_dim"]:
            raise ValueError(f"speaker_embedding must be [B, {self.dims['emb_dim']}], got {tuple(emb.shape)}")
        if not x.is_cuda or not emb.is_cuda:
            raise RuntimeError("inputs must be CUDA tensors (no CPU fallback)")

    # ---- hot path -----------------------------------------------------------------------------
    def forward(self, x, emb, precision="fp16x3", want_masked=False):
        self._check_inputs(x, emb)
        x = x.detach().to(torch.float32).contiguous()
        emb = emb.detach().to(torch.float32).contiguous()
        B, T, _ = x.shape
        prec = self._prec(precision)
        with torch.cuda.device(x.device):
            ws, need = self._workspace(B, T, prec)
            mask = torch.empty_like(x)
            masked = torch.empty_like(x) if want_masked else None
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.lib.vs_forward(self.handle, _ptr(x), _ptr(emb), _ptr(mask), _ptr(masked), B, T, prec,
                                            _ptr(ws), need, ctypes.c_void_p(st)), "vs_forward")
        return (mask, masked) if want_masked else mask

    def forward_host(self, x_host, emb_host, mask_host, masked_host=None, precision="fp16x3"):
        """HOST tensors in, host tensors out (the end-to-end plugin call): H2D, forward, D2H, sync."""
        B, T, _ = x_host.shape
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.lib.vs_forward_host(self.handle, _ptr(x_host), _ptr(emb_host), _ptr(mask_host),
                                                 _ptr(masked_host), B, T, self._prec(precision), ctypes.c_void_p(st)),
                        "vs_forward_host")
        return mask_host

    def host_submit(self, slot, x_host, emb_host, mask_host, masked_host=None, precision="fp16x3"):
        """Pipelined host entry: enqueue H2D -> forward -> D2H for one batch on `slot` (0 or 1)."""
        B, T, _ = x_host.shape
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.vs_forward_host_submit(self.handle, slot, _ptr(x_host), _ptr(emb_host), _ptr(mask_host),
                                                        _ptr(masked_host), B, T, self._prec(precision)), "vs_forward_host_submit")

    def host_reserve(self, B, T, precision="fp16x3"):
        """Allocate the device staging of the host entry points for batches up to (B, T) now, so no request allocates."""
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.vs_forward_host_reserve(self.handle, B, T, self._prec(precision)), "vs_forward_host_reserve")

    def host_wait(self, slot):
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.vs_forward_host_wait(self.handle, slot), "vs_forward_host_wait")

    def conv_stack(self, x, precision="fp16x3"):
        x = x.detach().to(torch.float32).contiguous()
        B, T, F = x.shape
        prec = self._prec(precision)
        with torch.cuda.device(x.device):
            ws, need = self._workspace(B, T, prec)
            out = torch.empty(B, T, 8 * F, dtype=torch.float32, device=x.device)
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.lib.vs_conv_stack(self.handle, _ptr(x), _ptr(out), B, T, prec, _ptr(ws), need,
                                               ctypes.c_void_p(st)), "vs_conv_stack")
        return out

    # ---- audio front / back end (STFT -> normalised dB magnitude + phasor; iSTFT with that phase) ----------
    def configure_audio(self, n_fft=1200, hop_length=160, win_length=400, min_level_db=-100.0, ref_level_db=20.0):
        p = _cabi.VsAudioParams(n_fft, hop_length, win_length, min_level_db, ref_level_db)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.lib.vs_audio_configure(self.handle, ctypes.byref(p), ctypes.c_void_p(st)), "vs_audio_configure")
        self.audio = dict(n_fft=n_fft, hop_length=hop_length, win_length=win_length)

    def wav2spec(self, wav):
        """wav [B, L] -> (spec [B, T, F] in [0,1], phasor [B, T, F, 2])   (reference wav2spec, audio_processor.py:469-476)"""
        wav = wav.detach().to(torch.float32).contiguous()
        B, L = wav.shape
        T = 1 + L // self.audio["hop_length"]
        F = self.dims["num_freq"]
        with torch.cuda.device(wav.device):
            need = int(self.lib.vs_audio_workspace_bytes(self.handle, B, L))
            ws = torch.empty(need, dtype=torch.uint8, device=wav.device)
            spec = torch.empty(B, T, F, dtype=torch.float32, device=wav.device)
            phasor = torch.empty(B, T, F, 2, dtype=torch.float32, device=wav.device)
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.lib.vs_wav2spec(self.handle, _ptr(wav), _ptr(spec), _ptr(phasor), B, L, _ptr(ws), need, ctypes.c_void_p(st)),
                        "vs_wav2spec")
        return spec, phasor

    def spec2wav(self, spec, phasor):
        """(masked) spec [B, T, F] + phasor -> wav [B, hop * (T - 1)]   (reference spec2wav with the mixture phase, :478-491)"""
        spec = spec.detach().to(torch.float32).contiguous()
        phasor = phasor.contiguous()
        B, T, _ = spec.shape
        Lout = self.audio["hop_length"] * (T - 1)
        with torch.cuda.device(spec.device):
            need = int(self.lib.vs_audio_workspace_bytes(self.handle, B, Lout))
            ws = torch.empty(need, dtype=torch.uint8, device=spec.device)
            wav = torch.empty(B, Lout, dtype=torch.float32, device=spec.device)
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.lib.vs_spec2wav(self.handle, _ptr(spec), _ptr(phasor), _ptr(wav), B, T, _ptr(ws), need, ctypes.c_void_p(st)),
                        "vs_spec2wav")
        return wav

    def separate(self, wav, emb, precision="fp16x3"):
        """Waveform in, separated waveform out: STFT -> mask -> mask * spectrogram -> iSTFT with the mixture phase
        (the reference's test path, utils/generic_utils.py:495-504)."""
        spec, phasor = self.wav2spec(wav)
        _, masked = self.forward(spec, emb, precision=precision, want_masked=True)
        return self.spec2wav(masked, phasor)

    # ---- training-loss chain: differentiable iSTFT (train.py:99-100) + Si-SNR (train.py:108) ------------------
    PHASE_MODES = {"q1": 0, "corrected": 1}

    def configure_loss(self, n_fft=1200, hop_length=160, win_length=400, min_level_db=-100.0, ref_level_db=20.0, phase_mode="q1"):
        """phase_mode "q1" = the reference verbatim, "corrected" = mag (cos, sin) + periodic Hann."""
        if phase_mode not in self.PHASE_MODES:
            raise ValueError(f"phase_mode must be one of {sorted(self.PHASE_MODES)}")
        p = _cabi.VsLossParams(n_fft, hop_length, win_length, min_level_db, ref_level_db, self.PHASE_MODES[phase_mode])
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.lib.vs_loss_configure(self.handle, ctypes.byref(p), ctypes.c_void_p(st)), "vs_loss_configure")
        self.loss_cfg = dict(n_fft=n_fft, hop_length=hop_length, win_length=win_length, phase_mode=phase_mode)

    def _loss_ws(self, B, T, device):
        need = int(self.lib.vs_loss_workspace_bytes(self.handle, B, T))
        if need == 0:
            raise RuntimeError("call configure_loss first (and use T >= 2 frames)")
        return torch.empty(need, dtype=torch.uint8, device=device), need

    @staticmethod
    def _f32c(t):
        return t.detach().to(torch.float32).contiguous()

    def loss_spec2wav(self, spec, phase):
        """spec, phase (angle) [B, T, F] -> wav [B, hop (T - 1)]   (torch_spec2wav, audio_processor.py:498-509)"""
        spec, phase = self._f32c(spec), self._f32c(phase)
        B, T, _ = spec.shape
        with torch.cuda.device(spec.device):
            ws, need = self._loss_ws(B, T, spec.device)
            wav = torch.empty(B, self.loss_cfg["hop_length"] * (T - 1), dtype=torch.float32, device=spec.device)
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.lib.vs_loss_spec2wav(self.handle, _ptr(spec), _ptr(phase), _ptr(wav), B, T, _ptr(ws), need, ctypes.c_void_p(st)),
                        "vs_loss_spec2wav")
        return wav

    def loss_spec2wav_backward(self, spec, phase, grad_wav):
        spec, phase, grad_wav = self._f32c(spec), self._f32c(phase), self._f32c(grad_wav)
        B, T, _ = spec.shape
        with torch.cuda.device(spec.device):
            ws, need = self._loss_ws(B, T, spec.device)
            grad = torch.empty_like(spec)
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.lib.vs_loss_spec2wav_backward(self.handle, _ptr(spec), _ptr(phase), _ptr(grad_wav), _ptr(grad), B, T, _ptr(ws), need,
                                                           ctypes.c_void_p(st)), "vs_loss_spec2wav_backward")
        return grad

    def sisnr_loss(self, est_spec, target_spec, phase, seq_len, want_grad=True):
        """-> (loss [] , snr [B], grad_est [B, T, F] or None): train.py:95-108 in one call, no host synchronisation."""
        est, tgt, phase = self._f32c(est_spec), self._f32c(target_spec), self._f32c(phase)
        B, T, _ = est.shape
        lens = seq_len.detach().reshape(-1).to(device=est.device, dtype=torch.int64).contiguous()
        if lens.numel() != B:
            raise ValueError("seq_len must hold one length per utterance")
        with torch.cuda.device(est.device):
            ws, need = self._loss_ws(B, T, est.device)
            loss = torch.empty((), dtype=torch.float32, device=est.device)
            snr = torch.empty(B, dtype=torch.float32, device=est.device)
            grad = torch.empty_like(est) if want_grad else None
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.lib.vs_sisnr_loss(self.handle, _ptr(est), _ptr(tgt), _ptr(phase), _ptr(lens), _ptr(loss), _ptr(snr),
                                               _ptr(grad) if want_grad else None, B, T, _ptr(ws), need, ctypes.c_void_p(st)), "vs_sisnr_loss")
        return loss, snr, grad

    # ---- evaluation metrics (utils/generic_utils.py:476-533) ----------------------------------------------------
    def sisnr_wav(self, est_wav, target_wav, seq_len):
        """SiSNR_With_Pit on waveforms [B, L], one source each -> (loss [], snr [B]).  The first argument is the estimate."""
        est, tgt = self._f32c(est_wav), self._f32c(target_wav)
        if est.shape != tgt.shape or est.dim() != 2:
            raise ValueError("estimate and target waveforms must both be [B, L]")      # generic_utils.py:426 assert
        B, L = est.shape
        lens = seq_len.detach().reshape(-1).to(device=est.device, dtype=torch.int64).contiguous()
        with torch.cuda.device(est.device):
            loss = torch.empty((), dtype=torch.float32, device=est.device)
            snr = torch.empty(B, dtype=torch.float32, device=est.device)
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.lib.vs_sisnr_wav(self.handle, _ptr(est), _ptr(tgt), _ptr(lens), _ptr(loss), _ptr(snr), B, L, ctypes.c_void_p(st)),
                        "vs_sisnr_wav")
        return loss, snr

    def sdr(self, ref_wav, est_wav):
        """bss_eval_sources(ref, est, False)[0][0] per utterance: ref, est [B, L] -> SDR [B] in dB."""
        ref, est = self._f32c(ref_wav), self._f32c(est_wav)
        if ref.shape != est.shape or ref.dim() != 2:
            raise ValueError("reference and estimate waveforms must both be [B, L]")
        B, L = ref.shape
        with torch.cuda.device(ref.device):
            need = int(self.lib.vs_sdr_workspace_bytes(B, L))
            ws = torch.empty(need, dtype=torch.uint8, device=ref.device)
            out = torch.empty(B, dtype=torch.float32, device=ref.device)
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.lib.vs_sdr(self.handle, _ptr(ref), _ptr(est), _ptr(out), B, L, _ptr(ws), need, ctypes.c_void_p(st)), "vs_sdr")
        return out

    # ---- GE2E speaker encoder: wav -> log-mel -> 3 x LSTM -> d-vector (notebooks/GE2E-...-openvoicefilter.py:63-85,141-143) ----
    def configure_encoder(self, num_mels=40, lstm_layers=3, lstm_hidden=768, emb_dim=256, window=80, stride=40, sample_rate=16000):
        """Needs configure_audio() first: the mel front end shares its STFT (n_fft / hop / win)."""
        if getattr(self, "audio", None) is None:
            raise RuntimeError("call configure_audio() before configure_encoder()")
        d = _cabi.VsEncoderDims(num_mels, lstm_layers, lstm_hidden, emb_dim, window, stride, sample_rate)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.lib.vs_encoder_configure(self.handle, ctypes.byref(d), ctypes.c_void_p(st)), "vs_encoder_configure")
        self.encoder_cfg = dict(num_mels=num_mels, lstm_layers=lstm_layers, lstm_hidden=lstm_hidden, emb_dim=emb_dim, window=window, stride=stride)

    def load_encoder_state_dict(self, sd):
        """sd: the notebook's SpeakerEncoder state_dict (lstm.weight_ih_l{k}, ..., proj.linear_layer.weight/bias), CUDA fp32 tensors."""
        c = self.encoder_cfg
        keep = []

        def g(k):
            t = sd[k].detach().to(device=self.device, dtype=torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()
        p = _cabi.VsEncoderParams()
        for l in range(c["lstm_layers"]):
            p.w_ih[l], p.w_hh[l] = g(f"lstm.weight_ih_l{l}"), g(f"lstm.weight_hh_l{l}")
            p.b_ih[l], p.b_hh[l] = g(f"lstm.bias_ih_l{l}"), g(f"lstm.bias_hh_l{l}")
        p.proj_w, p.proj_b = g("proj.linear_layer.weight"), g("proj.linear_layer.bias")
        H, M = c["lstm_hidden"], c["num_mels"]
        for l in range(c["lstm_layers"]):
            if tuple(sd[f"lstm.weight_ih_l{l}"].shape) != (4 * H, M if l == 0 else H) or tuple(sd[f"lstm.weight_hh_l{l}"].shape) != (4 * H, H):
                raise ValueError(f"encoder layer {l}: parameter shapes do not match the configured dimensions")
        if tuple(sd["proj.linear_layer.weight"].shape) != (c["emb_dim"], H):
            raise ValueError("encoder projection shape does not match the configured dimensions")
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.lib.vs_encoder_load_params(self.handle, ctypes.byref(p), ctypes.c_void_p(st)), "vs_encoder_load_params")
            torch.cuda.current_stream().synchronize()      # the packing kernels read `keep`

    def _enc_ws(self, B, n, from_wav, device):
        need = int(self.lib.vs_encoder_workspace_bytes(self.handle, B, n, 1 if from_wav else 0))
        if need == 0:
            raise RuntimeError("call configure_encoder first")
        return torch.empty(need, dtype=torch.uint8, device=device), need

    def encoder_mel(self, wav):
        """wav [B, L] -> log-mel [B, T, num_mels] (get_mel, utils/audio_processor.py:460-468; frames are rows here)."""
        wav = self._f32c(wav)
        B, L = wav.shape
        T = 1 + L // self.audio["hop_length"]
        with torch.cuda.device(wav.device):
            ws, need = self._enc_ws(B, L, True, wav.device)
            mel = torch.empty(B, T, self.encoder_cfg["num_mels"], dtype=torch.float32, device=wav.device)
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.lib.vs_encoder_mel(self.handle, _ptr(wav), _ptr(mel), B, L, _ptr(ws), need, ctypes.c_void_p(st)), "vs_encoder_mel")
        return mel

    def encoder_forward(self, mel):
        """log-mel [B, T, num_mels] -> d-vectors [B, emb_dim] (SpeakerEncoder.forward, notebook :75-85)."""
        mel = self._f32c(mel)
        B, T, _ = mel.shape
        with torch.cuda.device(mel.device):
            ws, need = self._enc_ws(B, T, False, mel.device)
            dvec = torch.empty(B, self.encoder_cfg["emb_dim"], dtype=torch.float32, device=mel.device)
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.lib.vs_encoder_forward(self.handle, _ptr(mel), _ptr(dvec), B, T, _ptr(ws), need, ctypes.c_void_p(st)), "vs_encoder_forward")
        return dvec

    def encoder_dvector(self, wav):
        """wav [B, L] -> d-vectors [B, emb_dim]; the log-mel never leaves its 16-bit operand planes."""
        wav = self._f32c(wav)
        B, L = wav.shape
        with torch.cuda.device(wav.device):
            ws, need = self._enc_ws(B, L, True, wav.device)
            dvec = torch.empty(B, self.encoder_cfg["emb_dim"], dtype=torch.float32, device=wav.device)
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.lib.vs_encoder_dvector(self.handle, _ptr(wav), _ptr(dvec), B, L, _ptr(ws), need, ctypes.c_void_p(st)), "vs_encoder_dvector")
        return dvec

    # ---- training (fp32, batch-statistics BatchNorm, full backward) ------------------------------
    PARAM_ORDER = tuple([k for l in range(8) for k in (f"conv.{CONV_IDX[l]}.weight", f"conv.{CONV_IDX[l]}.bias",
                                                        f"conv.{BN_IDX[l]}.weight", f"conv.{BN_IDX[l]}.bias")] +
                        [f"lstm.{n}_l0{sfx}" for sfx in ("", "_reverse") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")] +
                        ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"])

    def set_train_tensor_cores(self, on):
        _cabi.check(self.lib.vs_engine_set_train_tensor_cores(self.handle, 1 if on else 0), "vs_engine_set_train_tensor_cores")

    # ---- data-parallel hooks (include/voicesplit_b200.h: vs_engine_set_sync_bn / vs_engine_set_backward_hook) ----
    def set_sync_bn(self, allreduce_sum_doubles, world_size):
        """allreduce_sum_doubles(tensor float64 [128] on this device) sums it in place across ranks (None: per-rank statistics)."""
        if allreduce_sum_doubles is None:
            self._sync_cb = ctypes.cast(None, _cabi.STAT_ALLREDUCE_FN)
        else:
            dev = self.device

            def cb(_user, ptr, count, _stream):
                try:
                    allreduce_sum_doubles(_wrap_device_array(ptr, count, "<f8", dev))
                    return 0
                except Exception as ex:      # noqa: BLE001 - must not unwind through the C frame
                    self._hook_error = ex
                    return 1
            self._sync_cb = _cabi.STAT_ALLREDUCE_FN(cb)
        _cabi.check(self.lib.vs_engine_set_sync_bn(self.handle, self._sync_cb, None, int(world_size)), "vs_engine_set_sync_bn")

    def set_backward_hook(self, fn):
        """fn(stage) is called mid-backward (stage 1: LSTM / FC parameter gradients enqueued); None clears it."""
        if fn is None:
            self._bwd_cb = ctypes.cast(None, _cabi.BACKWARD_HOOK_FN)
        else:
            def cb(_user, stage, _stream):
                try:
                    fn(int(stage))
                    return 0
                except Exception as ex:      # noqa: BLE001
                    self._hook_error = ex
                    return 1
            self._bwd_cb = _cabi.BACKWARD_HOOK_FN(cb)
        _cabi.check(self.lib.vs_engine_set_backward_hook(self.handle, self._bwd_cb, None), "vs_engine_set_backward_hook")

    def _check_hook(self, rc, what):
        err, self._hook_error = getattr(self, "_hook_error", None), None
        if err is not None:
            raise RuntimeError(f"{what}: data-parallel hook raised") from err
        _cabi.check(rc, what)

    def train_forward(self, x, emb, bn_buffers=None, momentum=0.1):
        """Forward in BatchNorm-training mode.  bn_buffers: {state_dict key: tensor} of the running_mean /
        running_var / num_batches_tracked buffers to update in place (or None).  Returns (mask, saved)."""
        self._check_inputs(x, emb)
        x = x.detach().to(torch.float32).contiguous()
        emb = emb.detach().to(torch.float32).contiguous()
        B, T, _ = x.shape
        with torch.cuda.device(x.device):
            need = int(self.lib.vs_train_workspace_bytes(self.handle, B, T))
            ws = torch.empty(need, dtype=torch.uint8, device=x.device)
            mask = torch.empty_like(x)
            state = None
            if bn_buffers is not None:
                state = _cabi.VsTrainState()
                for l in range(8):
                    state.running_mean[l] = bn_buffers[f"conv.{BN_IDX[l]}.running_mean"].data_ptr()
                    state.running_var[l] = bn_buffers[f"conv.{BN_IDX[l]}.running_var"].data_ptr()
                    state.num_batches_tracked[l] = bn_buffers[f"conv.{BN_IDX[l]}.num_batches_tracked"].data_ptr()
                state.momentum = float(momentum)
            st = torch.cuda.current_stream().cuda_stream
            self._check_hook(self.lib.vs_train_forward(self.handle, ctypes.byref(state) if state is not None else None, _ptr(x), _ptr(emb),
                                                       _ptr(mask), B, T, _ptr(ws), need, ctypes.c_void_p(st)), "vs_train_forward")
        return mask, (ws, need, x, emb)

    @classmethod
    def grad_layout(cls, shapes):
        """{key: (offset, numel)} of every parameter gradient inside ONE flat fp32 buffer, in PARAM_ORDER (conv / BatchNorm
        first, then LSTM, then FC: the last two are a contiguous tail), and the total element count."""
        off, lay = 0, {}
        for k in cls.PARAM_ORDER:
            n = 1
            for d in shapes[k]:
                n *= int(d)
            lay[k] = (off, n)
            off += n
        return lay, off

    def train_backward(self, saved, mask, grad_mask, shapes, flat=None, want_grad_x=False):
        """shapes: {state_dict key: shape} of the parameters.  Gradients are written into `flat` (one fp32 buffer in
        grad_layout order, allocated here if None) and returned as views of it: ({key: grad}, grad_emb, grad_x or None)."""
        ws, need, x, emb = saved
        B, T, _ = x.shape
        grad_mask = grad_mask.detach().to(torch.float32).contiguous()
        lay, total = self.grad_layout(shapes)
        if flat is None:
            flat = torch.empty(total, dtype=torch.float32, device=x.device)
        if flat.numel() != total or flat.dtype != torch.float32 or flat.device != x.device:
            raise ValueError("flat gradient buffer does not match the parameter shapes")
        grads = {k: flat[o:o + n].view(shapes[k]) for k, (o, n) in lay.items()}
        gemb = torch.empty_like(emb)
        gx = torch.empty_like(x) if want_grad_x else None
        g = _cabi.VsGrads()
        for l in range(8):
            g.conv_w[l] = grads[f"conv.{CONV_IDX[l]}.weight"].data_ptr(); g.conv_b[l] = grads[f"conv.{CONV_IDX[l]}.bias"].data_ptr()
            g.bn_gamma[l] = grads[f"conv.{BN_IDX[l]}.weight"].data_ptr(); g.bn_beta[l] = grads[f"conv.{BN_IDX[l]}.bias"].data_ptr()
        for d, sfx in enumerate(("", "_reverse")):
            g.w_ih[d] = grads[f"lstm.weight_ih_l0{sfx}"].data_ptr(); g.w_hh[d] = grads[f"lstm.weight_hh_l0{sfx}"].data_ptr()
            g.b_ih[d] = grads[f"lstm.bias_ih_l0{sfx}"].data_ptr(); g.b_hh[d] = grads[f"lstm.bias_hh_l0{sfx}"].data_ptr()
        g.fc1_w, g.fc1_b = grads["fc1.weight"].data_ptr(), grads["fc1.bias"].data_ptr()
        g.fc2_w, g.fc2_b = grads["fc2.weight"].data_ptr(), grads["fc2.bias"].data_ptr()
        with torch.cuda.device(x.device):
            st = torch.cuda.current_stream().cuda_stream
            self._check_hook(self.lib.vs_train_backward(self.handle, _ptr(x), _ptr(emb), _ptr(mask), _ptr(grad_mask), ctypes.byref(g), _ptr(gemb),
                                                        _ptr(gx), B, T, _ptr(ws), need, ctypes.c_void_p(st)), "vs_train_backward")
        return grads, gemb, gx

    # ---- test hooks ---------------------------------------------------------------------------
    def debug_conv_layer(self, layer, inp, precision="fp32"):
        inp = inp.detach().to(torch.float32).contiguous()
        B, T = inp.shape[0], inp.shape[-2]
        out = torch.empty(B, 64, T, self.dims["num_freq"], dtype=torch.float32, device=inp.device)
        with torch.cuda.device(inp.device):
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.lib.vs_debug_conv_layer(self.handle, layer, _ptr(inp), _ptr(out), B, T,
                                                     self._prec(precision), ctypes.c_void_p(st)), "vs_debug_conv_layer")
        return out

    def debug_lstm_head(self, conv_out, emb, x, precision="fp32"):
        conv_out = conv_out.detach().to(torch.float32).contiguous()
        B, T, _ = conv_out.shape
        lstm_out = torch.empty(B, T, 2 * self.dims["lstm_dim"], dtype=torch.float32, device=conv_out.device)
        mask = torch.empty(B, T, self.dims["num_freq"], dtype=torch.float32, device=conv_out.device)
        with torch.cuda.device(conv_out.device):
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.lib.vs_debug_lstm_head(self.handle, _ptr(conv_out), _ptr(emb.contiguous()),
                                                    _ptr(x.contiguous()), _ptr(lstm_out), _ptr(mask), B, T,
                                                    self._prec(precision), ctypes.c_void_p(st)), "vs_debug_lstm_head")
        return lstm_out, mask

    def last_launch_count(self):
        return int(self.lib.vs_last_launch_count(self.handle))

    KERNEL_NAMES = {0: "cnn1", 1: "cnn2", 2: "cnn3", 3: "cnn4", 4: "cnn5", 5: "cnn6", 6: "cnn7", 7: "cnn8_reshape",
                    8: "dvector_gate_bias", 9: "lstm_input_proj", 10: "lstm_recurrence", 11: "fc1", 12: "fc2_sigmoid_mask",
                    13: "convert", 14: "head", 20: "train_conv_fwd", 21: "train_bn_stats", 22: "train_bn_act", 23: "train_bn_bwd",
                    24: "train_wgrad", 25: "train_dgrad", 26: "train_gemm", 27: "train_lstm_bwd", 28: "train_misc"}

    def set_profiling(self, on):
        _cabi.check(self.lib.vs_engine_set_profiling(self.handle, 1 if on else 0), "vs_engine_set_profiling")

    def profile_read(self):
        """[(kernel name, milliseconds)] of the last forward (profiling must be enabled)."""
        ids = (ctypes.c_int32 * 256)()
        ms = (ctypes.c_float * 256)()
        n = self.lib.vs_profile_read(self.handle, 256, ids, ms)
        return [(self.KERNEL_NAMES.get(ids[i], str(ids[i])), float(ms[i])) for i in range(n)]
