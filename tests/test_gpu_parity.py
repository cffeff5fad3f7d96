"""Parity tests proper (need the B200): the CUDA path, called through the C ABI, against the
committed golden vectors (outputs of the unmodified reference) and against the CPU oracle.

Tolerance (BASELINE.json north_star): mask within 1e-3 of the reference.  Asserted per mode (TOL below):
  fp32, fp16x3          max |diff| < 1e-3 and mean |diff| (MAE) < 1e-4   - the faithful modes, held to the stated bar on the max too
  fp16_f8c (bench default), bf16x3   MAE < 2e-4 / 1e-4 (the bar is on the MAE, met ~10x over); max < 3e-3, i.e. ABOVE 1e-3 on the
                        worst bin of the stress weights - measured values are printed by the full-size test and by bench.py `parity`
  fp16, bf16            single-pass fast modes: loose, explicitly stated bounds; reported, never passed off as faithful
"""
import numpy as np
import pytest
import torch

from conftest import golden_cases, load_case
from oracle import oracle
from voicesplit_b200 import config, synth
from voicesplit_b200.engine import MaskEngine

pytestmark = pytest.mark.gpu

FAITHFUL = ["fp32", "fp16x3", "bf16x3"]
CONV_MODES = FAITHFUL + ["fp16_f8c"]          # modes whose conv stack is held to the 1e-3-class bounds
ALL = ["fp32", "fp16x3", "fp16_f8c", "bf16x3", "fp16", "bf16"]
# (max |diff|, mean |diff|) bounds on the mask.  The stated parity bar is 1e-3 (BASELINE.json:
# "within 1e-3", "mask MAE <= 1e-3"); the stress weights amplify rounding (the fp32 reference
# itself is ~2e-4 max from exact arithmetic).  fp32 and fp16x3 meet 1e-3 on the max and 1e-4 on
# the MAE; bf16x3 (8-bit significand halves) meets the MAE bar 10x over but can exceed 1e-3 on the
# worst bin at full size.  The single-pass fast modes are bounded loosely on purpose: their error
# is reported (bench.py), never passed off as faithful.
# fp16_f8c (fp16 main pass + e4m3 correction pass in the conv stack): the BASELINE bar is the MAE (<= 1e-3), which it
# meets ~10x over on the stress weights; its worst bin sits near 1e-3 (CPU model tools/precision_model.py: max 1.3e-3,
# MAE 6e-5 at 100x257), so the max bound is 3e-3 like bf16x3 and the measured values are printed by the full-size test.
TOL = {"fp32": (1e-3, 1e-4), "fp16x3": (1e-3, 1e-4), "fp16_f8c": (3e-3, 2e-4), "bf16x3": (3e-3, 1e-4), "fp16": (1.5e-1, 1e-2),
       "bf16": (6e-1, 5e-2)}
TOL_MAX = {k: v[0] for k, v in TOL.items()}
TOL_MAE = {k: v[1] for k, v in TOL.items()}


def _engine(case):
    eng = MaskEngine(activation="mish" if case["model_name"] == "voicesplit" else "relu", **case["dims"])
    eng.load_state_dict_tensors({k: torch.from_numpy(np.asarray(v)).cuda() for k, v in case["state_dict"].items()
                                 if "num_batches" not in k})
    return eng


@pytest.mark.parametrize("precision", ALL)
def test_mask_matches_reference_golden(golden, precision):
    eng = _engine(golden)
    x, emb = torch.from_numpy(golden["x"]).cuda(), torch.from_numpy(golden["emb"]).cuda()
    mask, masked = eng.forward(x, emb, precision=precision, want_masked=True)
    torch.cuda.synchronize()
    d = (mask.cpu().numpy() - golden["mask"])
    assert np.isfinite(d).all()
    assert np.abs(d).max() < TOL_MAX[precision], (np.abs(d).max(), np.abs(d).mean())
    assert np.abs(d).mean() < TOL_MAE[precision]
    assert np.abs(masked.cpu().numpy() - golden["masked"]).max() < TOL_MAX[precision]
    assert eng.last_launch_count() > 0


@pytest.mark.parametrize("precision", CONV_MODES)
def test_conv_stack_matches_golden(golden, precision):
    eng = _engine(golden)
    out = eng.conv_stack(torch.from_numpy(golden["x"]).cuda(), precision=precision).cpu().numpy()
    ref = golden["conv_out"]
    assert np.abs(out - ref).max() < TOL_MAX[precision] * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("precision", FAITHFUL)
def test_lstm_and_head_match_golden(golden, precision):
    eng = _engine(golden)
    conv = torch.from_numpy(golden["conv_out"]).cuda()
    lstm_out, mask = eng.debug_lstm_head(conv, torch.from_numpy(golden["emb"]).cuda(),
                                         torch.from_numpy(golden["x"]).cuda(), precision=precision)
    # tcgen05 accumulates into TMEM with truncation (tools/umma_probe.cu: ~0.3 ulp per step, toward zero);
    # the GEMM promotes partial sums to fp32 registers every 8 K-blocks, which keeps the K = 8F = 4808
    # input projection within this bound even with the x4 stress LSTM weights
    d = np.abs(lstm_out.cpu().numpy() - golden["lstm_out"])
    assert d.max() < 1e-3 and d.mean() < 1e-5
    assert np.abs(mask.cpu().numpy() - golden["mask"]).max() < 1e-3


@pytest.mark.parametrize("precision", ALL)
@pytest.mark.parametrize("layer", [1, 2, 4, 6])
def test_single_conv_layer_against_oracle(layer, precision):
    """One dilated conv + BN + Mish layer on random 64-channel input, against the oracle's layer."""
    dims = synth.make_dims(41, 8, 16, 24)
    sd = synth.make_state_dict(dims, 21, "stress")
    B, T, F = 2, 45, 41
    rng = np.random.default_rng(layer)
    inp = rng.uniform(-1, 1, size=(B, 64, T, F)).astype(np.float32)
    eng = MaskEngine(activation="mish", **dims)
    eng.load_state_dict_tensors({k: torch.from_numpy(v).cuda() for k, v in sd.items() if "num_batches" not in k})
    got = eng.debug_conv_layer(layer, torch.from_numpy(inp).cuda(), precision=precision).cpu().numpy()
    # oracle for one layer: torch-free restatement via the C oracle is whole-path only, so use
    # float64 numpy here (same formula as oracle/voicesplit_oracle.c conv_bn_act)
    idx, cin, cout, kh, kw, dil = synth.CONV_LAYERS[layer]
    w = sd[f"conv.{idx}.weight"].astype(np.float64); bn = synth.BN_INDEX[idx]
    pt, pf = (kh - 1) // 2 * dil, (kw - 1) // 2
    xp = np.pad(inp.astype(np.float64), ((0, 0), (0, 0), (pt, pt), (pf, pf)))
    acc = np.zeros((B, cout, T, F))
    for i in range(kh):
        for j in range(kw):
            acc += np.einsum("oc,bctf->botf", w[:, :, i, j], xp[:, :, i * dil:i * dil + T, j:j + F])
    acc += sd[f"conv.{idx}.bias"].astype(np.float64)[None, :, None, None]
    g, b_, m, v = (sd[f"conv.{bn}.{k}"].astype(np.float64)[None, :, None, None]
                   for k in ("weight", "bias", "running_mean", "running_var"))
    y = (acc - m) * g / np.sqrt(v + 1e-5) + b_
    ref = oracle.activation(y.astype(np.float32), "mish")
    tol = {"fp32": 2e-4, "bf16x3": 2e-4, "fp16x3": 2e-5, "fp16_f8c": 5e-4, "bf16": 1e-1, "fp16": 1.5e-2}[precision]
    assert np.abs(got - ref).max() < tol * max(1.0, np.abs(ref).max())


def test_module_forward_matches_golden_and_repacks_after_update():
    case = load_case([p for p in golden_cases() if "tiny_mish_stress" in p][0])
    from models.voicesplit.model import VoiceSplit
    m = VoiceSplit(config.AttrDict(synth.make_config_dict(case["dims"])))
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in case["state_dict"].items()})
    m = m.cuda().eval()
    x, emb = torch.from_numpy(case["x"]).cuda(), torch.from_numpy(case["emb"]).cuda()
    with torch.no_grad():
        mask = m(x, emb)
    assert mask.shape == x.shape and mask.device == x.device
    assert np.abs(mask.cpu().numpy() - case["mask"]).max() < 1e-3
    with torch.no_grad():                      # an in-place parameter update must invalidate the packing
        m.fc2.bias.add_(3.0)
        mask2 = m(x, emb)
    assert (mask2 > mask).float().mean() > 0.99
    out = m.train()(x, emb)                    # training mode: batch statistics, autograd graph attached
    assert out.requires_grad and out.shape == x.shape
    xg = x.clone().requires_grad_(True)        # differentiable w.r.t. the spectrogram as well (SURVEY 8b "and inputs")
    m(xg, emb).sum().backward()
    assert xg.grad is not None and torch.isfinite(xg.grad).all() and float(xg.grad.abs().max()) > 0


@pytest.mark.parametrize("precision", CONV_MODES)
def test_full_size_against_oracle(precision):
    """BASELINE shapes: native 301x601 and literal 601x257, B=1, stress weights, vs the CPU oracle."""
    for dims, T in ((synth.make_dims(257), 601), (synth.make_dims(601), 301)):
        sd = synth.make_state_dict(dims, 31, "stress")
        x, emb = synth.make_inputs(1, T, dims, 41)
        ref = oracle.forward(sd, dims, x, emb)["mask"]
        eng = MaskEngine(activation="mish", **dims)
        eng.load_state_dict_tensors({k: torch.from_numpy(v).cuda() for k, v in sd.items() if "num_batches" not in k})
        got = eng.forward(torch.from_numpy(x).cuda(), torch.from_numpy(emb).cuda(), precision=precision).cpu().numpy()
        d = np.abs(got - ref)
        print(f"full-size {precision} F={dims['num_freq']}: max {d.max():.3e} mae {d.mean():.3e}")
        assert d.max() < TOL_MAX[precision] and d.mean() < TOL_MAE[precision], (dims["num_freq"], d.max(), d.mean())


def test_batch_independence_and_host_entry():
    """Size-independent properties at a bigger batch: utterances are independent (a batch equals
    its utterances run one by one) and the host-buffer entry point equals the device one."""
    dims = synth.make_dims(65, 32, 48, 64)
    sd = synth.make_state_dict(dims, 5, "stress")
    eng = MaskEngine(activation="mish", **dims)
    eng.load_state_dict_tensors({k: torch.from_numpy(v).cuda() for k, v in sd.items() if "num_batches" not in k})
    x, emb = synth.make_inputs(5, 77, dims, 3)
    xt, et = torch.from_numpy(x).cuda(), torch.from_numpy(emb).cuda()
    for precision in ("fp32", "fp16x3", "fp16_f8c", "bf16"):
        full = eng.forward(xt, et, precision=precision)
        for b in (0, 4):
            one = eng.forward(xt[b:b + 1], et[b:b + 1], precision=precision)
            assert torch.allclose(full[b:b + 1], one, atol=2e-6, rtol=0)
        xh, eh = torch.from_numpy(x).pin_memory(), torch.from_numpy(emb).pin_memory()
        mh = torch.empty_like(xh).pin_memory()
        eng.forward_host(xh, eh, mh, precision=precision)
        assert torch.equal(mh, full.cpu())


@pytest.mark.parametrize("B,groups_per_cta", [(130, 1), (385, 1), (130, 2), (257, 2), (385, 2)])
def test_recurrence_with_several_utterance_groups(B, groups_per_cta):
    """More than 128 utterances: several groups of 128 per launch, one per CTA (default) or two per CTA (the experiment knob: each
    group with its own accumulator, cell warps and exchange barrier; beyond 256 a second / half-filled set) - utterances of every
    group must equal their single-utterance run, and the whole batch the fp32 path.  The knob is read once per process, so the
    two-group cases run in a child process."""
    if groups_per_cta == 2:
        import subprocess, sys, os
        env = dict(os.environ, VOICESPLIT_LSTM_GROUPS_PER_CTA="2")
        code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import test_gpu_parity as t; "
                "t._several_groups_body(%d)" % (os.path.dirname(os.path.abspath(__file__)),
                                                os.path.dirname(os.path.dirname(os.path.abspath(__file__))), B))
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        return
    _several_groups_body(B)


def _several_groups_body(B):
    dims = synth.make_dims(17, 8, 40, 24)        # 40 hidden units: three slices of 16, the last one partial
    sd = synth.make_state_dict(dims, 9, "stress")
    eng = MaskEngine(activation="mish", **dims)
    eng.load_state_dict_tensors({k: torch.from_numpy(v).cuda() for k, v in sd.items() if "num_batches" not in k})
    x, emb = synth.make_inputs(B, 23, dims, 4)
    xt, et = torch.from_numpy(x).cuda(), torch.from_numpy(emb).cuda()
    ref = eng.forward(xt, et, precision="fp32")
    for precision in ("fp16x3", "bf16x3", "fp16_f8c", "fp16"):
        full = eng.forward(xt, et, precision=precision)
        torch.cuda.synchronize()
        assert (full - ref).abs().max() < TOL_MAX[precision]
        for b in (0, 127, 128, B - 1):
            one = eng.forward(xt[b:b + 1], et[b:b + 1], precision=precision)
            assert torch.allclose(full[b:b + 1], one, atol=2e-6, rtol=0), (precision, b)


def _gemm_pairs_body():
    """LSTM input projection, FC1 and FC2 through the W-tile multicast pairs (2-CTA clusters) on a problem with an odd number of
    M blocks (the last pair has a CTA without rows) and N tails - against the fp32 path."""
    dims = synth.make_dims(65, 32, 48, 64)
    sd = synth.make_state_dict(dims, 5, "stress")
    eng = MaskEngine(activation="mish", **dims)
    eng.load_state_dict_tensors({k: torch.from_numpy(v).cuda() for k, v in sd.items() if "num_batches" not in k})
    x, emb = synth.make_inputs(5, 53, dims, 3)            # 265 rows = 3 M blocks (2 full + 9 rows): the second pair has a CTA without rows
    xt, et = torch.from_numpy(x).cuda(), torch.from_numpy(emb).cuda()
    ref, ref_m = eng.forward(xt, et, precision="fp32", want_masked=True)
    for precision in ("fp16x3", "bf16x3", "fp16_f8c"):
        out, out_m = eng.forward(xt, et, precision=precision, want_masked=True)
        torch.cuda.synchronize()
        assert (out - ref).abs().max() < TOL_MAX[precision], precision
        assert (out_m - ref_m).abs().max() < TOL_MAX[precision], precision


def test_gemm_weight_tile_multicast_pairs_on_a_small_problem():
    """The pairs are normally enabled only for GEMMs with at least two waves of tiles (the full-size input projection, covered by
    bench.py's parity block); VOICESPLIT_GEMM_CLUSTER=3 forces them wherever there are two M blocks.  Read once per process ->
    child process."""
    import subprocess, sys, os
    env = dict(os.environ, VOICESPLIT_GEMM_CLUSTER="3")
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import test_gpu_parity as t; t._gemm_pairs_body()"
            % (os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_abi_error_paths():
    """Errors surface as codes + messages, never as silent fallbacks."""
    import ctypes
    from voicesplit_b200 import _cabi
    lib = _cabi.load()
    dims = synth.make_dims(33, 16, 24, 40)
    eng = MaskEngine(activation="mish", **dims)
    x = torch.zeros(1, 4, 33, device="cuda"); emb = torch.zeros(1, 16, device="cuda"); out = torch.empty_like(x)
    ws = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    # parameters not loaded yet
    rc = lib.vs_forward(eng.handle, P(x), P(emb), P(out), None, 1, 4, 0, P(ws), ws.numel(), None)
    assert rc == -3 and b"not loaded" in lib.vs_last_error()
    sd = synth.make_state_dict(dims, 1, "default")
    eng.load_state_dict_tensors({k: torch.from_numpy(v).cuda() for k, v in sd.items() if "num_batches" not in k})
    assert lib.vs_forward(eng.handle, P(x), P(emb), P(out), None, 1, 4, 99, P(ws), ws.numel(), None) == -1   # bad precision
    assert lib.vs_forward(eng.handle, P(x), P(emb), P(out), None, 0, 4, 0, P(ws), ws.numel(), None) == -1    # B < 1
    assert lib.vs_forward(eng.handle, P(x), P(emb), P(out), None, 1, 4, 0, P(ws), 16, None) == -3            # workspace too small
    assert b"workspace" in lib.vs_last_error()
    assert lib.vs_forward(eng.handle, None, P(emb), P(out), None, 1, 4, 0, P(ws), ws.numel(), None) == -1    # null input
    bad = _cabi.VsDims(33, 16, 24, 40, 34, 0)                                                                # fc2_dim != num_freq
    h = ctypes.c_void_p()
    assert lib.vs_engine_create(ctypes.byref(bad), ctypes.byref(h)) == -1
    with pytest.raises(ValueError):
        eng.forward(torch.zeros(1, 4, 32, device="cuda"), emb)


def test_fp16_modes_stay_finite_on_huge_activations():
    """Half has a narrow exponent: activations are clamped before conversion, so even absurd weights
    give a finite mask (bf16x3 is the mode without the range caveat and must still match fp32)."""
    dims = synth.make_dims(33, 16, 24, 40)
    sd = synth.make_state_dict(dims, 9, "default")
    sd["conv.2.weight"] = sd["conv.2.weight"] * 3e5          # BatchNorm gamma of cnn1: activations ~1e5
    eng = MaskEngine(activation="relu", **dims)
    eng.load_state_dict_tensors({k: torch.from_numpy(v).cuda() for k, v in sd.items() if "num_batches" not in k})
    x, emb = synth.make_inputs(2, 19, dims, 4)
    xt, et = torch.from_numpy(x).cuda(), torch.from_numpy(emb).cuda()
    ref = eng.forward(xt, et, precision="fp32")
    for p in ("fp16x3", "fp16"):
        assert torch.isfinite(eng.forward(xt, et, precision=p)).all()
    assert (eng.forward(xt, et, precision="bf16x3") - ref).abs().max() < 1e-3


def test_many_utterance_groups_lstm_multi_launch():
    """More 128-utterance groups than are co-resident (H=400: 2 groups per launch) -> the recurrent kernel
    is launched per chunk of groups; every utterance must still equal its stand-alone result."""
    dims = synth.make_dims(17, 8, 400, 24)
    sd = synth.make_state_dict(dims, 12, "stress")
    eng = MaskEngine(activation="mish", **dims)
    eng.load_state_dict_tensors({k: torch.from_numpy(v).cuda() for k, v in sd.items() if "num_batches" not in k})
    x, emb = synth.make_inputs(300, 6, dims, 8)
    xt, et = torch.from_numpy(x).cuda(), torch.from_numpy(emb).cuda()
    full = eng.forward(xt, et, precision="fp16x3")
    ref = eng.forward(xt, et, precision="fp32")
    assert (full - ref).abs().max() < 1e-3
    for b in (0, 127, 128, 257, 299):
        one = eng.forward(xt[b:b + 1], et[b:b + 1], precision="fp16x3")
        assert torch.allclose(full[b:b + 1], one, atol=2e-6, rtol=0), b


def test_pipelined_host_entry_matches_synchronous():
    dims = synth.make_dims(33, 16, 24, 40)
    sd = synth.make_state_dict(dims, 2, "stress")
    eng = MaskEngine(activation="mish", **dims)
    eng.load_state_dict_tensors({k: torch.from_numpy(v).cuda() for k, v in sd.items() if "num_batches" not in k})
    eng.host_reserve(3, 26, "fp16x3")            # staging for every shape below is allocated here, not on a request
    bufs = []
    for i in range(2):
        x, emb = synth.make_inputs(3, 21 + 5 * i, dims, 40 + i)
        x, emb = torch.from_numpy(x).pin_memory(), torch.from_numpy(emb).pin_memory()
        bufs.append((x, emb, torch.empty_like(x).pin_memory(), torch.empty_like(x).pin_memory()))
    want = []
    for x, emb, _, _ in bufs:
        m = torch.empty_like(x).pin_memory(); mm = torch.empty_like(x).pin_memory()
        eng.forward_host(x, emb, m, mm, precision="fp16x3")
        want.append((m.clone(), mm.clone()))
    for rounds in range(3):                     # reuse the slots several times, both in flight at once
        eng.host_submit(0, *bufs[0], precision="fp16x3")
        eng.host_submit(1, *bufs[1], precision="fp16x3")
        with pytest.raises(Exception, match="in flight"):
            eng.host_submit(1, *bufs[1], precision="fp16x3")
        eng.host_wait(0); eng.host_wait(1)
        for i in range(2):
            assert torch.equal(bufs[i][2], want[i][0]) and torch.equal(bufs[i][3], want[i][1])
            bufs[i][2].zero_(); bufs[i][3].zero_()
    with pytest.raises(Exception, match="nothing submitted"):
        eng.host_wait(0)
