"""TEST INFRASTRUCTURE - CPU restatement (float64, torch autograd) of the reference's training-loss chain:

  openVoiceFilterAudioProcessor.torch_spec2wav   /root/reference/utils/audio_processor.py:498-509
  SiSNR_With_Pit (with get_mask)                 /root/reference/utils/generic_utils.py:403-474
  as driven by                                   /root/reference/train.py:95-109

torch_spec2wav calls torchaudio.functional.istft (third-party, removed from current torchaudio; upstreamed as
torch.istft).  Its published algorithm is restated here from first principles - one-sided inverse real DFT of every
frame, multiplication by the centre-padded window, overlap-add, division by the overlap-added squared window,
trimming n_fft/2 samples each side - WITHOUT calling torch.istft, so the restatement is independent of it.
Pinned by tests/test_loss_oracle.py against golden vectors produced by the unmodified reference code
(tests/golden/make_loss_golden.py; loss value, waveforms and autograd gradient).

mode "q1" is the reference verbatim: real = mag * e^{cos phi}, imag = mag * e^{sin phi}, symmetric
Hann synthesis window.  mode "corrected" is what the code evidently meant: mag * (cos phi, sin phi) with the periodic
Hann of the analysis side (utils/audio_processor.py:511-514, librosa default).
Only tests/, smoke() and bench.py's baseline legs may import this module.
"""
import math

import torch

EPS = 1e-16   # generic_utils.py:420


def hann(win, periodic, dtype=torch.float64):
    n = torch.arange(win, dtype=dtype)
    return 0.5 - 0.5 * torch.cos(2 * math.pi * n / (win if periodic else win - 1))


def spec2wav(spec, phase, n_fft, hop, win, min_db=-100.0, ref_db=20.0, mode="q1"):
    """spec, phase: [B, T, F] (normalised dB magnitude, angle in radians) -> waveform [B, hop * (T - 1)].  Differentiable."""
    spec = spec.to(torch.float64)
    phase = phase.to(torch.float64)
    B, T, F = spec.shape
    S = (torch.clamp(spec, 0.0, 1.0) - 1.0) * -min_db + ref_db               # :502-503
    mag = torch.pow(10.0, S * 0.05)                                          # :505
    c, s = torch.cos(phase), torch.sin(phase)
    if mode == "q1":
        re, im = mag * torch.exp(c), mag * torch.exp(s)                      # :507-509, exp() of the (cos, sin) pair
    elif mode == "corrected":
        re, im = mag * c, mag * s
    else:
        raise ValueError(mode)
    # one-sided inverse DFT: x[n] = (1/N) sum_k c_k (re_k cos(2 pi k n / N) - im_k sin(2 pi k n / N)); c_0 = c_{N/2} = 1, else 2
    n = torch.arange(n_fft, dtype=torch.float64)
    k = torch.arange(F, dtype=torch.float64)
    ang = 2 * math.pi * torch.outer(k, n) / n_fft
    ck = torch.full((F, 1), 2.0, dtype=torch.float64)
    ck[0] = 1.0
    if n_fft % 2 == 0:
        ck[-1] = 1.0
    frames = (re @ (ck * torch.cos(ang)) - im @ (ck * torch.sin(ang))) / n_fft    # [B, T, n_fft]
    w = torch.zeros(n_fft, dtype=torch.float64)
    lp = (n_fft - win) // 2
    w[lp:lp + win] = hann(win, periodic=(mode == "corrected"))
    frames = frames * w
    total = n_fft + hop * (T - 1)
    idx = (torch.arange(T)[:, None] * hop + torch.arange(n_fft)[None, :]).reshape(-1)
    y = torch.zeros(B, total, dtype=torch.float64).index_add(1, idx, frames.reshape(B, -1))
    env = torch.zeros(total, dtype=torch.float64).index_add(0, idx, (w * w).repeat(T))
    half = n_fft // 2
    y, env = y[:, half:total - half], env[half:total - half]
    return y / torch.where(env > 1e-11, env, torch.ones_like(env))


def si_snr_c1(est, tgt, lengths):
    """est, tgt: [B, L] (one source per utterance, the only case train.py produces: C = 1); lengths [B] -> (loss, snr [B])."""
    B, L = tgt.shape
    n = lengths.to(est.dtype).view(B, 1)
    m = (torch.arange(L)[None, :] < lengths.view(B, 1)).to(est.dtype)     # get_mask :403-415
    est = est * m                                                          # :434
    zt = (tgt - tgt.sum(1, keepdim=True) / n) * m                          # :437-444 (the target mean runs over ALL samples)
    ze = (est - est.sum(1, keepdim=True) / n) * m
    dot = (ze * zt).sum(1, keepdim=True)
    energy = (zt * zt).sum(1, keepdim=True) + EPS
    proj = dot * zt / energy
    noise = ze - proj
    snr = 10 * torch.log10((proj * proj).sum(1) / ((noise * noise).sum(1) + EPS) + EPS)
    return 20 - snr.mean(), snr                                            # :470-473 with C = 1


def loss_and_grad(est_spec, tgt_spec, phase, lengths, n_fft, hop, win, min_db=-100.0, ref_db=20.0, mode="q1"):
    """numpy/torch in -> dict(loss, snr [B], wav_est, wav_tgt, grad_est) as float64 numpy arrays."""
    as_t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(a)
    e = as_t(est_spec).detach().to(torch.float64).requires_grad_(True)
    we = spec2wav(e, as_t(phase), n_fft, hop, win, min_db, ref_db, mode)
    wt = spec2wav(as_t(tgt_spec), as_t(phase), n_fft, hop, win, min_db, ref_db, mode)
    loss, snr = si_snr_c1(we, wt, as_t(lengths).to(torch.int64))
    loss.backward()
    return {"loss": float(loss.detach()), "snr": snr.detach().numpy(), "wav_est": we.detach().numpy(), "wav_tgt": wt.detach().numpy(),
            "grad_est": e.grad.numpy()}
