"""The training criterion of the reference, two ways.

si_snr_with_pit   - SiSNR_With_Pit (utils/generic_utils.py:417-474) restated with plain torch ops for any number of
                    sources C (general PIT); elementwise/reduction glue on [B, C, L] tensors.
SpecSiSNRLoss     - what train.py:95-109 actually runs every step with loss_name "si_snr": both spectrograms through the
                    differentiable iSTFT ap.torch_inv_spectrogram (utils/audio_processor.py:498-509) and
                    then the C = 1 criterion, as ONE engine call (vs_sisnr_loss: tcgen05 iSTFT GEMMs, fused reductions,
                    analytic gradient, iSTFT backward) wrapped in an autograd Function.
spec2wav_autograd - the differentiable iSTFT alone (vs_loss_spec2wav / _backward), for callers that keep the two
                    steps separate like the reference does."""
from __future__ import annotations

from itertools import permutations

import torch

EPS = 1e-16   # generic_utils.py:420


def si_snr_with_pit(estimate, source, lengths):
    """estimate, source: [B, C, L]; lengths: [B] valid samples.  Returns the scalar loss 20 - mean(max Si-SNR)."""
    B, C, L = source.shape
    idx = torch.arange(L, device=source.device)[None, None, :]
    m = (idx < lengths.view(-1, 1, 1)).to(source.dtype)                  # get_mask, generic_utils.py:403-415 (no Python loop)
    n = lengths.view(-1, 1, 1).to(source.dtype)
    est = estimate * m
    zt = (source - source.sum(2, keepdim=True) / n) * m                  # zero-mean, re-masked (:432-441)
    ze = (est - est.sum(2, keepdim=True) / n) * m
    s_t, s_e = zt.unsqueeze(1), ze.unsqueeze(2)                          # [B,1,C,L], [B,C,1,L]
    dot = (s_e * s_t).sum(3, keepdim=True)
    energy = (s_t ** 2).sum(3, keepdim=True) + EPS
    proj = dot * s_t / energy
    noise = s_e - proj
    snr = 10 * torch.log10((proj ** 2).sum(3) / ((noise ** 2).sum(3) + EPS) + EPS)   # [B,C,C]
    perms = torch.tensor(list(permutations(range(C))), device=source.device)
    onehot = torch.zeros(perms.shape[0], C, C, device=source.device, dtype=source.dtype).scatter_(2, perms.unsqueeze(2), 1)
    snr_set = torch.einsum("bij,pij->bp", snr, onehot)
    return 20 - (snr_set.max(dim=1).values / C).mean()


class _SpecSiSNRFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, est_spec, target_spec, phase, seq_len, engine):
        loss, snr, grad = engine.sisnr_loss(est_spec, target_spec, phase, seq_len, want_grad=ctx.needs_input_grad[0])
        ctx.has_grad = grad is not None
        if ctx.has_grad:
            ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(snr)
        return loss, snr

    @staticmethod
    def backward(ctx, g_loss, _g_snr):
        if not ctx.has_grad:
            return None, None, None, None, None
        (grad,) = ctx.saved_tensors
        return grad * g_loss, None, None, None, None


class SpecSiSNRLoss(torch.nn.Module):
    """loss = SpecSiSNRLoss(engine, audio_config)(mixed * mask, target_spec, spec_phase, seq_len)

    replaces   output = ap.torch_inv_spectrogram(output, spec_phase); target = ap.torch_inv_spectrogram(target, spec_phase);
               loss = SiSNR_With_Pit()(output[:, None], target[:, None], seq_len)            (train.py:99-108)
    phase_mode "q1" reproduces the reference verbatim, "corrected" uses mag (cos, sin) and the analysis window.
    The gradient reaches est_spec only (the target and the phase are data), as in the reference's training step."""

    def __init__(self, engine, audio_config, phase_mode="q1"):
        super().__init__()
        self.engine = engine
        engine.configure_loss(audio_config["n_fft"], audio_config["hop_length"], audio_config["win_length"],
                              audio_config.get("min_level_db", -100.0), audio_config.get("ref_level_db", 20.0), phase_mode)
        self.last_snr = None

    def forward(self, est_spec, target_spec, spec_phase, seq_len):
        loss, self.last_snr = _SpecSiSNRFn.apply(est_spec, target_spec, spec_phase, seq_len, self.engine)
        return loss


class _Spec2WavFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec, phase, engine):
        ctx.engine = engine
        ctx.save_for_backward(spec, phase)
        return engine.loss_spec2wav(spec, phase)

    @staticmethod
    def backward(ctx, g_wav):
        spec, phase = ctx.saved_tensors
        return ctx.engine.loss_spec2wav_backward(spec, phase, g_wav), None, None


def spec2wav_autograd(engine, spec, phase):
    """Differentiable torch_spec2wav on the engine (configure_loss first): spec, phase [B, T, F] -> wav [B, hop (T - 1)]."""
    return _Spec2WavFn.apply(spec, phase, engine)
