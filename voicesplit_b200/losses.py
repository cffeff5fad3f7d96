"""Si-SNR with PIT - the training criterion of the reference (utils/generic_utils.py:417-474), restated
with plain torch ops for the training benchmark and tests.  It sits AFTER the hot path (train.py:108)
and is elementwise/reduction work on [B, C, L] tensors, so it is host-side glue here, not a kernel.
The reference applies it to waveforms obtained by its (buggy) differentiable iSTFT,
which is a SURVEY section 8(f) "next" row; the benchmark applies it to the flattened masked
spectrograms, which exercises exactly the same backward through the mask."""
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
