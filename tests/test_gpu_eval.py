"""Evaluation metrics and driver on the device (vs_sdr, vs_sisnr_wav, voicesplit_b200.evaluate) against the oracles
(oracle/sdr_oracle.py, oracle/loss_oracle.py) and an item-by-item all-oracle evaluation."""
import numpy as np
import pytest
import torch

from oracle import audio_oracle as ao, loss_oracle, sdr_oracle, torch_port
from voicesplit_b200 import config as vconfig, evaluate, synth
from voicesplit_b200.audio import DeviceAudioProcessor

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    from models.voicesplit.model import VoiceSplit
    dims = synth.make_dims(601, 256, 400, 600)
    m = VoiceSplit(vconfig.AttrDict(synth.make_config_dict(dims)))
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_state_dict(dims, 3, "stress").items()})
    return m.cuda().eval()


def _pairs(B, L, seed, snr_db=(25, 12, 3, -5)):
    rng = np.random.Generator(np.random.PCG64(seed))
    ref = synth.make_reference_audio(B, L, seed).astype(np.float64)
    est = np.zeros_like(ref)
    for b in range(B):
        other = synth.make_reference_audio(1, L, 1000 + seed + b)[0]
        g = 10 ** (-snr_db[b % len(snr_db)] / 20) * np.sqrt((ref[b] ** 2).sum() / (other ** 2).sum())
        est[b] = 0.8 * np.convolve(ref[b], [0.9, 0.3, -0.1])[:L] + g * other + 1e-4 * rng.standard_normal(L)
    return ref.astype(np.float32), est.astype(np.float32)


@pytest.mark.parametrize("B,L", [(4, 16000), (3, 48000), (2, 5000), (1, 700)])
def test_sdr_matches_oracle(model, B, L):
    ref, est = _pairs(B, L, 7 + B)
    got = model.engine().sdr(torch.from_numpy(ref).cuda(), torch.from_numpy(est).cuda()).cpu().numpy()
    for b in range(B):
        want = sdr_oracle.sdr(ref[b], est[b])
        assert abs(got[b] - want) <= 2e-3, (b, got[b], want)          # dB


def test_sdr_closed_form_and_scale_invariance(model):
    ref, est = _pairs(2, 8000, 3)
    eng = model.engine()
    r, e = torch.from_numpy(ref).cuda(), torch.from_numpy(est).cuda()
    a, b = eng.sdr(r, e).cpu().numpy(), eng.sdr(r, 4.0 * e).cpu().numpy()
    assert np.abs(a - b).max() <= 1e-3
    delayed = torch.zeros_like(r)
    delayed[:, 5:-8] = 0.4 * r[:, :-13]
    r2 = r.clone(); r2[:, -13:] = 0
    assert float(eng.sdr(r2, delayed).min()) > 60.0                   # a filtered copy of the reference lies in the span


def test_sisnr_wav_matches_oracle_in_both_argument_orders(model):
    ref, est = _pairs(3, 12000, 21)
    lens = np.array([12000, 9000, 11999], dtype=np.int64)
    eng = model.engine()
    r, e, ln = torch.from_numpy(ref).cuda(), torch.from_numpy(est).cuda(), torch.from_numpy(lens).cuda()
    loss, snr = eng.sisnr_wav(e, r, ln)
    want_loss, want = loss_oracle.si_snr_c1(torch.from_numpy(est).double(), torch.from_numpy(ref).double(), torch.from_numpy(lens))
    assert np.abs(snr.cpu().numpy() - want.numpy()).max() <= 1e-3 and abs(float(loss) - float(want_loss)) <= 1e-3
    _, swapped = eng.sisnr_wav(r, e, ln)                              # Q2: validation() passes (clean, est)
    want_sw = loss_oracle.si_snr_c1(torch.from_numpy(ref).double(), torch.from_numpy(est).double(), torch.from_numpy(lens))[1]
    assert np.abs(swapped.cpu().numpy() - want_sw.numpy()).max() <= 1e-3
    # for full-length zero-mean signals Si-SNR = 10 log10(cos^2 / (1 - cos^2)) is symmetric; only the length mask (item 1) breaks it
    d = np.abs(swapped.cpu().numpy() - snr.cpu().numpy())
    assert d[0] <= 1e-3 and d[2] <= 1e-3 and d[1] > 1e-5


def _loader(n_items, L=16000):
    """Items in the layout of utils/dataset.py:42-57, spectrograms made by the oracle front end."""
    items = []
    for i in range(n_items):
        clean = synth.make_reference_audio(1, L, 50 + i)[0]
        other = synth.make_reference_audio(1, L, 90 + i)[0]
        mixed = (clean + 0.7 * other).astype(np.float32)
        ms, mp = ao.wav2spec(mixed)
        cs, _ = ao.wav2spec(clean)
        emb = np.random.Generator(np.random.PCG64(i)).standard_normal(256).astype(np.float32) * 0.05
        items.append([(torch.from_numpy(emb), torch.from_numpy(cs.astype(np.float32)), torch.from_numpy(ms.astype(np.float32)), torch.from_numpy(clean),
                       torch.from_numpy(mixed), torch.from_numpy(mp.astype(np.float32)), torch.from_numpy(np.array([L])))])
    return items


def test_validation_driver_matches_item_by_item_oracle_evaluation(model):
    loader = _loader(5)
    ap = DeviceAudioProcessor(model.engine(), dict(n_fft=1200, hop_length=160, win_length=400))
    stats = {}
    mean_loss, mean_sdr = evaluate.validation(None, ap, model, loader, None, 0, cuda=True, loss_name="si_snr", test=True, batch_size=4, stats=stats)
    assert stats["items"] == 5 and stats["batches"] == 2 and stats["length_mismatch"] == 0
    dims = synth.make_dims(601, 256, 400, 600)
    sd = synth.make_state_dict(dims, 3, "stress")
    ref_losses, ref_sdrs = [], []
    for (emb, cs, ms, clean, mixed, mp, seq_len), in loader:
        mask = torch_port.forward(sd, ms[None].numpy(), emb[None].numpy(), "mish").numpy()[0]
        est = ao.spec2wav(mask * ms.numpy(), mp.numpy())
        # validation(): criterion(clean, est) - swapped (Q2)
        ref_losses.append(float(loss_oracle.si_snr_c1(clean[None].double(), torch.from_numpy(est)[None].double(), seq_len)[0]))
        ref_sdrs.append(sdr_oracle.sdr(clean.numpy(), est))
    assert abs(mean_loss - np.mean(ref_losses)) <= 5e-3 and abs(mean_sdr - np.mean(ref_sdrs)) <= 5e-3, (mean_loss, np.mean(ref_losses), mean_sdr, np.mean(ref_sdrs))
    # test=False: first item only, nothing returned (generic_utils.py:512-519)
    assert evaluate.validation(None, ap, model, loader, None, 0, loss_name="si_snr", test=False) is None
    with pytest.raises(ValueError):
        evaluate.validation(None, ap, model, loader, None, 0, loss_name="nope", test=True)      # no blanket except


def test_fast_si_snr_sweep_matches_fused_loss(model):
    loader = _loader(3)
    ap = DeviceAudioProcessor(model.engine(), dict(n_fft=1200, hop_length=160, win_length=400))
    batched = [tuple(torch.stack([it[0][j] for it in loader]) for j in range(7))]           # one batch of 3, as a DataLoader would collate
    got = evaluate.test_fast_with_si_srn(None, ap, model, batched)
    dims = synth.make_dims(601, 256, 400, 600)
    sd = synth.make_state_dict(dims, 3, "stress")
    emb, cs, ms, _, _, mp, seq_len = batched[0]
    mask = torch_port.forward(sd, ms.numpy(), emb.numpy(), "mish").numpy()
    want = loss_oracle.loss_and_grad(mask * ms.numpy(), cs.numpy(), mp.numpy(), seq_len.reshape(-1).numpy(), 1200, 160, 400, mode="q1")["loss"]
    assert abs(got - want) <= 2e-3
