"""TEST INFRASTRUCTURE - ctypes wrapper around oracle/voicesplit_oracle.c (see its header).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libvs_oracle.so")
_lib = None

CONV_IDX = (1, 5, 9, 13, 17, 21, 25, 28)          # nn.Sequential positions of the convs
BN_IDX = (2, 6, 10, 14, 18, 22, 26, 29)
ACT = {"mish": 0, "voicesplit": 0, "relu": 1, "voicefilter": 1}


_P = ctypes.POINTER(ctypes.c_float)


class _Params(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in
                ("num_freq", "emb_dim", "lstm_dim", "fc1_dim", "fc2_dim", "activation")] + \
               [(n, _P * 8) for n in ("conv_w", "conv_b", "bn_g", "bn_b", "bn_m", "bn_v")] + \
               [(n, _P * 2) for n in ("w_ih", "w_hh", "b_ih", "b_hh")] + \
               [(n, _P) for n in ("fc1_w", "fc1_b", "fc2_w", "fc2_b")]


def build(force=False):
    src = os.path.join(_HERE, "voicesplit_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.vs_oracle_forward.restype = ctypes.c_int
        _lib.vs_oracle_activation.restype = None
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _as_np(v):
    if hasattr(v, "detach"):
        v = v.detach().cpu().numpy()
    return np.ascontiguousarray(v, dtype=np.float32)


def forward(state_dict, dims, x, emb, activation="mish", want=(), dump_layer=-1):
    """Run the oracle.  Returns dict with 'mask' and any of want = ('masked','conv_out','lstm_out','dump')."""
    lib = _load()
    sd = {k: _as_np(v) for k, v in state_dict.items() if "num_batches_tracked" not in k}
    keep = [sd]  # keep arrays alive
    p = _Params()
    p.num_freq, p.emb_dim, p.lstm_dim = dims["num_freq"], dims["emb_dim"], dims["lstm_dim"]
    p.fc1_dim, p.fc2_dim, p.activation = dims["fc1_dim"], dims["fc2_dim"], ACT[activation]
    for l in range(8):
        p.conv_w[l] = _fp(sd[f"conv.{CONV_IDX[l]}.weight"]); p.conv_b[l] = _fp(sd[f"conv.{CONV_IDX[l]}.bias"])
        p.bn_g[l] = _fp(sd[f"conv.{BN_IDX[l]}.weight"]); p.bn_b[l] = _fp(sd[f"conv.{BN_IDX[l]}.bias"])
        p.bn_m[l] = _fp(sd[f"conv.{BN_IDX[l]}.running_mean"]); p.bn_v[l] = _fp(sd[f"conv.{BN_IDX[l]}.running_var"])
    for d, sfx in enumerate(("", "_reverse")):
        p.w_ih[d] = _fp(sd[f"lstm.weight_ih_l0{sfx}"]); p.w_hh[d] = _fp(sd[f"lstm.weight_hh_l0{sfx}"])
        p.b_ih[d] = _fp(sd[f"lstm.bias_ih_l0{sfx}"]); p.b_hh[d] = _fp(sd[f"lstm.bias_hh_l0{sfx}"])
    p.fc1_w, p.fc1_b = _fp(sd["fc1.weight"]), _fp(sd["fc1.bias"])
    p.fc2_w, p.fc2_b = _fp(sd["fc2.weight"]), _fp(sd["fc2.bias"])
    x = _as_np(x); emb = _as_np(emb)
    B, T, F = x.shape
    assert F == dims["num_freq"] and emb.shape == (B, dims["emb_dim"])
    H = dims["lstm_dim"]
    out = {"mask": np.empty((B, T, F), np.float32)}
    null = ctypes.POINTER(ctypes.c_float)()
    opt = {"masked": (B, T, F), "conv_out": (B, T, 8 * F), "lstm_out": (B, T, 2 * H)}
    ptrs = {}
    for k, shp in opt.items():
        if k in want:
            out[k] = np.empty(shp, np.float32); ptrs[k] = _fp(out[k])
        else:
            ptrs[k] = null
    dump = null
    if "dump" in want:
        cout = 8 if dump_layer == 7 else 64
        out["dump"] = np.empty((B, cout, T, F), np.float32); dump = _fp(out["dump"])
    rc = lib.vs_oracle_forward(ctypes.byref(p), _fp(x), _fp(emb), ctypes.c_int(B), ctypes.c_int(T),
                               _fp(out["mask"]), ptrs["masked"], ptrs["conv_out"], ptrs["lstm_out"],
                               ctypes.c_int(dump_layer), dump)
    if rc != 0:
        raise RuntimeError(f"vs_oracle_forward failed rc={rc}")
    del keep
    return out


def activation(x, kind="mish"):
    lib = _load()
    x = _as_np(x)
    y = np.empty_like(x)
    lib.vs_oracle_activation(_fp(x), _fp(y), ctypes.c_long(x.size), ctypes.c_int(ACT[kind]))
    return y
