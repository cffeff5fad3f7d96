"""Training-loss chain on the device (vs_loss_spec2wav / _backward, vs_sisnr_loss) against
 (1) the golden vectors produced by the unmodified reference (tests/golden/loss_*.npz) and
 (2) the float64 oracle (oracle/loss_oracle.py) on other shapes, lengths and both phase modes."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import loss_oracle
from voicesplit_b200 import synth
from voicesplit_b200.engine import MaskEngine
from voicesplit_b200.losses import SpecSiSNRLoss, spec2wav_autograd

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "loss_*.npz")))
_engines = {}


def _engine(n_fft):
    """The loss kernels only need an engine whose num_freq matches n_fft / 2 + 1 (the other dims are irrelevant)."""
    if n_fft not in _engines:
        dims = synth.make_dims(n_fft // 2 + 1, 16, 24, 40)
        eng = MaskEngine(activation="mish", **dims)
        sd = synth.make_state_dict(dims, 3, "default")
        eng.load_state_dict_tensors({k: torch.from_numpy(v).cuda() for k, v in sd.items() if "num_batches" not in k})
        _engines[n_fft] = eng
    return _engines[n_fft]


def _cuda(*arrays):
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrays]


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_matches_reference_golden(path):
    g = np.load(path)
    n_fft, hop, win, B, T = (int(g[k]) for k in ("n_fft", "hop", "win", "B", "T"))
    eng = _engine(n_fft)
    eng.configure_loss(n_fft, hop, win, float(g["min_db"]), float(g["ref_db"]), "q1")
    est, tgt, phase = synth.loss_inputs(n_fft, B, T, int(g["seed"]))
    e, t, p, lens = _cuda(est, tgt, phase, g["lengths"])
    wav = eng.loss_spec2wav(e, p).cpu().numpy()
    assert wav.shape == g["wav_est"].shape
    assert np.abs(wav - g["wav_est"]).max() <= 1e-4 * np.abs(g["wav_est"]).max()
    loss, snr, grad = eng.sisnr_loss(e, t, p, lens)
    assert abs(float(loss) - float(g["loss"])) <= 1e-3
    gs = np.abs(g["grad_est"]).max()
    d = np.abs(grad.cpu().numpy() - g["grad_est"])
    assert d.max() <= 1e-3 * gs and d.mean() <= 1e-4 * gs, (d.max() / gs, d.mean() / gs)


@pytest.mark.parametrize("mode", ["q1", "corrected"])
@pytest.mark.parametrize("n_fft,hop,win,B,T,ragged", [(1200, 160, 400, 3, 61, True), (128, 32, 64, 5, 50, True), (1200, 160, 400, 2, 301, False)])
def test_loss_and_gradient_match_oracle(mode, n_fft, hop, win, B, T, ragged):
    eng = _engine(n_fft)
    eng.configure_loss(n_fft, hop, win, -100.0, 20.0, mode)
    est, tgt, phase = synth.loss_inputs(n_fft, B, T, 100 + T)
    est = (0.6 * tgt + 0.4 * est).astype(np.float32)             # correlated estimate: a less degenerate Si-SNR than pure noise
    L = hop * (T - 1)
    lens = np.array([L - (b * L) // (2 * B) if ragged else L for b in range(B)], dtype=np.int64)
    ref = loss_oracle.loss_and_grad(est, tgt, phase, lens, n_fft, hop, win, mode=mode)
    e, t, p, ln = _cuda(est, tgt, phase, lens)
    loss, snr, grad = eng.sisnr_loss(e, t, p, ln)
    assert abs(float(loss) - ref["loss"]) <= 1e-3
    assert np.abs(snr.cpu().numpy() - ref["snr"]).max() <= 2e-3
    gs = np.abs(ref["grad_est"]).max()
    d = np.abs(grad.cpu().numpy() - ref["grad_est"])
    assert d.max() <= 1e-3 * gs and d.mean() <= 1e-4 * gs, (d.max() / gs, d.mean() / gs)
    # forward-only call gives the same loss and touches no gradient buffer
    loss2, _, none = eng.sisnr_loss(e, t, p, ln, want_grad=False)
    assert none is None and float(loss2) == float(loss)


def test_spec2wav_backward_matches_autograd_of_oracle():
    n_fft, hop, win, B, T = 128, 32, 64, 3, 33
    eng = _engine(n_fft)
    for mode in ("q1", "corrected"):
        eng.configure_loss(n_fft, hop, win, -100.0, 20.0, mode)
        est, _, phase = synth.loss_inputs(n_fft, B, T, 9)
        rng = np.random.Generator(np.random.PCG64(1))
        gw = rng.standard_normal((B, hop * (T - 1))).astype(np.float32) * 1e-3
        es = torch.from_numpy(est).to(torch.float64).requires_grad_(True)
        w = loss_oracle.spec2wav(es, torch.from_numpy(phase), n_fft, hop, win, mode=mode)
        (w * torch.from_numpy(gw)).sum().backward()
        e, p, g = _cuda(est, phase, gw)
        ed = e.clone().requires_grad_(True)
        wd = spec2wav_autograd(eng, ed, p)
        assert np.abs(wd.detach().cpu().numpy() - w.detach().numpy()).max() <= 1e-4 * float(w.detach().abs().max())
        wd.backward(g)
        ref = es.grad.numpy()
        assert np.abs(ed.grad.cpu().numpy() - ref).max() <= 1e-3 * np.abs(ref).max()


def test_module_gradient_flows_to_the_mask():
    n_fft, hop, win, B, T = 128, 32, 64, 2, 21
    eng = _engine(n_fft)
    crit = SpecSiSNRLoss(eng, dict(n_fft=n_fft, hop_length=hop, win_length=win), "q1")
    mixed, tgt, phase = synth.loss_inputs(n_fft, B, T, 31)
    mixed = np.clip(mixed, 0, 1)
    mask = np.random.Generator(np.random.PCG64(2)).random(mixed.shape).astype(np.float32)
    lens = np.array([hop * (T - 1)] * B, dtype=np.int64)
    m64 = torch.from_numpy(mask).to(torch.float64).requires_grad_(True)
    est64 = torch.from_numpy(mixed).to(torch.float64) * m64
    we = loss_oracle.spec2wav(est64, torch.from_numpy(phase), n_fft, hop, win, mode="q1")
    wt = loss_oracle.spec2wav(torch.from_numpy(tgt), torch.from_numpy(phase), n_fft, hop, win, mode="q1")
    l64, _ = loss_oracle.si_snr_c1(we, wt, torch.from_numpy(lens))
    (3.0 * l64).backward()
    x, t, p, ln, mk = _cuda(mixed, tgt, phase, lens.reshape(B, 1), mask)          # seq_len arrives as [B, 1] (utils/dataset.py:37)
    mk.requires_grad_(True)
    loss = crit(x * mk, t, p, ln)
    (3.0 * loss).backward()
    assert abs(float(loss) - float(l64)) <= 1e-3
    ref = m64.grad.numpy()
    assert np.abs(mk.grad.cpu().numpy() - ref).max() <= 1e-3 * np.abs(ref).max()
    assert crit.last_snr.shape == (B,)


def test_errors_are_loud():
    dims = synth.make_dims(65, 16, 24, 40)
    eng = MaskEngine(activation="mish", **dims)
    with pytest.raises(RuntimeError):
        eng.configure_loss(128, 32, 60)                     # win not a multiple of 8
    with pytest.raises(RuntimeError):
        eng.configure_loss(256, 32, 64)                     # n_fft / 2 + 1 != num_freq
    with pytest.raises(ValueError):
        eng.configure_loss(128, 32, 64, phase_mode="fixed")
    x = torch.zeros(1, 4, 65, device="cuda")
    with pytest.raises(RuntimeError):
        eng.loss_spec2wav(x, x)                             # not configured
