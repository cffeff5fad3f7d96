"""Diagnostic: training gradients with the 2-D paired cnn2 tiles vs flat tiles (VOICESPLIT_CONV_TILE2D), same inputs."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    from voicesplit_b200 import config, synth
    from models.voicesplit.model import VoiceSplit
    F, B, T = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    dims = synth.make_dims(F)
    sd = synth.make_state_dict(dims, 15, "stress")
    x, emb = synth.make_inputs(B, T, dims, 16)
    gw = np.random.default_rng(1).standard_normal((B, T, F)).astype(np.float32)
    m = VoiceSplit(config.AttrDict(synth.make_config_dict(dims)))
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    m = m.cuda().train()
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    mask = m(xt, torch.from_numpy(emb).cuda())
    (mask * torch.from_numpy(gw).cuda()).sum().backward()
    torch.cuda.synchronize()
    out = {k: p.grad.double().cpu().numpy() for k, p in m.named_parameters()}
    out["x"] = xt.grad.double().cpu().numpy(); out["mask"] = mask.detach().double().cpu().numpy()
    np.savez(sys.argv[5], **out)
    sys.exit(0)
import numpy as np
CASES = {"flat": dict(VOICESPLIT_CONV_TILE2D="0"), "2d": dict(VOICESPLIT_CONV_TILE2D="1"), "2d_again": dict(VOICESPLIT_CONV_TILE2D="1"),
         "fwd2d_only": dict(VOICESPLIT_CONV_TILE2D="1", VOICESPLIT_CONV_TILE2D_DGRAD="0"),
         "dgrad2d_only": dict(VOICESPLIT_CONV_TILE2D="0", VOICESPLIT_CONV_TILE2D_DGRAD="1"), "flat_again": dict(VOICESPLIT_CONV_TILE2D="0")}
for F, B, T in ((257, 2, 601),):
  for name, ev in CASES.items():
    if name == "flat":
        subprocess.check_call([sys.executable, __file__, "child", str(F), str(B), str(T), "/tmp/diag_t0.npz"], env=dict(os.environ, **ev))
        continue
    subprocess.check_call([sys.executable, __file__, "child", str(F), str(B), str(T), "/tmp/diag_t1.npz"], env=dict(os.environ, **ev))
    a, b = np.load("/tmp/diag_t0.npz"), np.load("/tmp/diag_t1.npz")
    print(f"--- F={F} B={B} T={T}: flat vs {name}: max |diff| / max |flat|")
    for k in a.files:
        d = np.abs(a[k] - b[k]).max() / max(np.abs(a[k]).max(), 1e-30)
        if (d > 1e-4 and not k.endswith(".bias")) or k in ("mask", "x", "fc2.weight", "fc1.bias"):
            extra = ""
            if k == "x":  # noqa
                dd = np.abs(a[k] - b[k]); idx = np.unravel_index(dd.argmax(), dd.shape); extra = f" argmax (b,t,f)={idx}"
            print(f"   {k:28s} {d:.3e}{extra}")
