"""Diagnostic: snapshot the training workspace after vs_train_forward / vs_train_backward with flat vs 2-D cnn2 tiles and
locate the regions (train.cu: train_carve order) that differ."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from voicesplit_b200 import synth
from voicesplit_b200.engine import MaskEngine

F, B, T = 257, 2, 601
H, N1, E = 400, 600, 256
dims = synth.make_dims(F)
sd = synth.make_state_dict(dims, 15, "stress")
x, emb = synth.make_inputs(B, T, dims, 16)
gw = torch.from_numpy(np.random.default_rng(1).standard_normal((B, T, F)).astype(np.float32)).cuda()
xt, et = torch.from_numpy(x).cuda(), torch.from_numpy(emb).cuda()
Fp = (F + 2 + 7) // 8 * 8
M = B * T
plane = M * Fp * 64 * 4
al = lambda v: (v + 1023) // 1024 * 1024
regions, off = [], 0
def take(name, nbytes, kind="f32"):
    global off
    regions.append((name, off, nbytes, kind)); off += al(nbytes)
for i in range(7): take(f"z{i}", plane)
for n in ("P", "G1", "G2"): take(n, plane)
for n in ("Ahi", "Alo", "Dhi", "Dlo"): take(n, plane // 2, "u16")
for n in ("z7", "xcat", "dxcat"): take(n, M * 8 * F * 4)
take("gates", M * 8 * H * 4); take("bias_u", B * 8 * H * 4)
for n in ("hout", "cseq", "hprev", "dh"): take(n, M * 2 * H * 4)
take("y1", M * N1 * 4); take("dy1", M * N1 * 4); take("dz2", M * F * 4)
take("stat", 8 * 256 * 4); take("sums", 128 * 8, "f64"); take("dwp", 49 * 64 * 64 * 4); take("dsum", B * 8 * H * 4)

shapes = {k: tuple(np.asarray(v).shape) for k, v in sd.items() if np.asarray(v).dtype == np.float32 and "running" not in k}
snaps = {}
for name, v in (("flat", "0"), ("2d", "1")):
    os.environ["VOICESPLIT_CONV_TILE2D"] = v
    os.environ["VOICESPLIT_CONV_TILE2D_DGRAD"] = "0"
    eng = MaskEngine(activation="mish", **dims)
    eng.load_state_dict_tensors({k: torch.from_numpy(v_).cuda() for k, v_ in sd.items() if "num_batches" not in k})
    need = int(eng.lib.vs_train_workspace_bytes(eng.handle, B, T))
    junk = torch.full((need,), 0x7f, dtype=torch.uint8, device="cuda"); del junk
    mask, saved = eng.train_forward(xt, et, None)
    torch.cuda.synchronize()
    fwd = saved[0].clone()
    grads, gemb, gx = eng.train_backward(saved, mask, gw, shapes)
    torch.cuda.synchronize()
    snaps[name] = (fwd, saved[0].clone(), mask.clone(), {k: g.clone() for k, g in grads.items()})
    del eng
print("workspace bytes", snaps["flat"][0].numel(), "carved through dsum", off)
for stage, idx in (("after forward", 0), ("after backward", 1)):
    print("==", stage)
    a, b = snaps["flat"][idx], snaps["2d"][idx]
    for name, o, n, kind in regions:
        ra, rb = a[o:o + n], b[o:o + n]
        if torch.equal(ra, rb):
            continue
        if kind == "f32":
            fa, fb = ra.view(torch.float32), rb.view(torch.float32)
            d = (fa - fb).abs(); d = torch.where(torch.isfinite(d), d, torch.full_like(d, 1e30))
            scale = float(fa[torch.isfinite(fa)].abs().max()) if torch.isfinite(fa).any() else 1.0
            nb = int((d > 1e-3 * max(scale, 1e-30)).sum())
            first = int(torch.nonzero(d > 1e-3 * max(scale, 1e-30))[0]) if nb else -1
            print(f"   {name:8s} differs: max |diff| {float(d.max()):.3e} (scale {scale:.3e}), elements > 1e-3 scale: {nb}, first element {first}")
        else:
            print(f"   {name:8s} differs (bits)")
print("== gradients")
for k in snaps["flat"][3]:
    ga, gb = snaps["flat"][3][k], snaps["2d"][3][k]
    d = float((ga - gb).abs().max() / ga.abs().max().clamp(min=1e-30))
    if d > 1e-4 and not k.endswith(".bias"):
        print(f"   {k:28s} {d:.3e}")
