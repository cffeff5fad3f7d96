"""ctypes binding of include/voicesplit_b200.h (the C ABI of libvoicesplit_sm100.so).

There is deliberately no fallback: if the shared library is missing or a symbol is absent the
import raises, so a GPU box can never silently run a different implementation.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvoicesplit_sm100.so")

VS_OK = 0
ACT_MISH, ACT_RELU = 0, 1
PREC_FP32, PREC_BF16X3, PREC_BF16, PREC_FP16X3, PREC_FP16, PREC_FP16_F8C = 0, 1, 2, 3, 4, 5
PRECISIONS = {"fp32": PREC_FP32, "bf16x3": PREC_BF16X3, "bf16": PREC_BF16, "fp16x3": PREC_FP16X3, "fp16": PREC_FP16,
              "fp16_f8c": PREC_FP16_F8C}

_fp = ctypes.POINTER(ctypes.c_float)


class VsDims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("num_freq", "emb_dim", "lstm_dim", "fc1_dim", "fc2_dim", "activation")]


class VsParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p * 8) for n in
                ("conv_w", "conv_b", "bn_gamma", "bn_beta", "bn_mean", "bn_var")] + \
               [(n, ctypes.c_void_p * 2) for n in ("w_ih", "w_hh", "b_ih", "b_hh")] + \
               [(n, ctypes.c_void_p) for n in ("fc1_w", "fc1_b", "fc2_w", "fc2_b")]


class VsAudioParams(ctypes.Structure):
    _fields_ = [("n_fft", ctypes.c_int32), ("hop_length", ctypes.c_int32), ("win_length", ctypes.c_int32),
                ("min_level_db", ctypes.c_float), ("ref_level_db", ctypes.c_float)]


class VsLossParams(ctypes.Structure):
    _fields_ = [("n_fft", ctypes.c_int32), ("hop_length", ctypes.c_int32), ("win_length", ctypes.c_int32),
                ("min_level_db", ctypes.c_float), ("ref_level_db", ctypes.c_float), ("phase_mode", ctypes.c_int32)]


class VsEncoderDims(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in ("num_mels", "lstm_layers", "lstm_hidden", "emb_dim", "window", "stride", "sample_rate")]


class VsEncoderParams(ctypes.Structure):
    _fields_ = [("w_ih", ctypes.c_void_p * 4), ("w_hh", ctypes.c_void_p * 4), ("b_ih", ctypes.c_void_p * 4), ("b_hh", ctypes.c_void_p * 4),
                ("proj_w", ctypes.c_void_p), ("proj_b", ctypes.c_void_p)]


class VsTrainState(ctypes.Structure):
    _fields_ = [("running_mean", ctypes.c_void_p * 8), ("running_var", ctypes.c_void_p * 8),
                ("num_batches_tracked", ctypes.c_void_p * 8), ("momentum", ctypes.c_float)]


class VsGrads(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p * 8) for n in ("conv_w", "conv_b", "bn_gamma", "bn_beta")] + \
               [(n, ctypes.c_void_p * 2) for n in ("w_ih", "w_hh", "b_ih", "b_hh")] + \
               [(n, ctypes.c_void_p) for n in ("fc1_w", "fc1_b", "fc2_w", "fc2_b")]


# name -> (restype, argtypes); must list every function include/voicesplit_b200.h declares
_VP, _I, _SZ = ctypes.c_void_p, ctypes.c_int32, ctypes.c_size_t
# data-parallel callbacks (vs_stat_allreduce_fn, vs_backward_hook_fn)
STAT_ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, _VP, _VP, _I, _VP)
BACKWARD_HOOK_FN = ctypes.CFUNCTYPE(ctypes.c_int, _VP, _I, _VP)
BWD_STAGE_LSTM_FC_DONE = 1
SIGNATURES = {
    "vs_abi_version": (ctypes.c_int, []),
    "vs_last_error": (ctypes.c_char_p, []),
    "vs_engine_create": (ctypes.c_int, [ctypes.POINTER(VsDims), ctypes.POINTER(_VP)]),
    "vs_engine_destroy": (ctypes.c_int, [_VP]),
    "vs_engine_load_params": (ctypes.c_int, [_VP, ctypes.POINTER(VsParams), _VP]),
    "vs_workspace_bytes": (_SZ, [_VP, _I, _I, _I]),
    "vs_forward": (ctypes.c_int, [_VP, _VP, _VP, _VP, _VP, _I, _I, _I, _VP, _SZ, _VP]),
    "vs_forward_host": (ctypes.c_int, [_VP, _VP, _VP, _VP, _VP, _I, _I, _I, _VP]),
    "vs_forward_host_submit": (ctypes.c_int, [_VP, _I, _VP, _VP, _VP, _VP, _I, _I, _I]),
    "vs_forward_host_wait": (ctypes.c_int, [_VP, _I]),
    "vs_forward_host_reserve": (ctypes.c_int, [_VP, _I, _I, _I]),
    "vs_engine_set_train_tensor_cores": (ctypes.c_int, [_VP, _I]),
    "vs_train_workspace_bytes": (_SZ, [_VP, _I, _I]),
    "vs_train_forward": (ctypes.c_int, [_VP, ctypes.POINTER(VsTrainState), _VP, _VP, _VP, _I, _I, _VP, _SZ, _VP]),
    "vs_train_backward": (ctypes.c_int, [_VP, _VP, _VP, _VP, _VP, ctypes.POINTER(VsGrads), _VP, _VP, _I, _I, _VP, _SZ, _VP]),
    "vs_engine_set_sync_bn": (ctypes.c_int, [_VP, STAT_ALLREDUCE_FN, _VP, _I]),
    "vs_engine_set_backward_hook": (ctypes.c_int, [_VP, BACKWARD_HOOK_FN, _VP]),
    "vs_audio_configure": (ctypes.c_int, [_VP, ctypes.POINTER(VsAudioParams), _VP]),
    "vs_audio_workspace_bytes": (_SZ, [_VP, _I, _I]),
    "vs_wav2spec": (ctypes.c_int, [_VP, _VP, _VP, _VP, _I, _I, _VP, _SZ, _VP]),
    "vs_spec2wav": (ctypes.c_int, [_VP, _VP, _VP, _VP, _I, _I, _VP, _SZ, _VP]),
    "vs_loss_configure": (ctypes.c_int, [_VP, ctypes.POINTER(VsLossParams), _VP]),
    "vs_loss_workspace_bytes": (_SZ, [_VP, _I, _I]),
    "vs_loss_spec2wav": (ctypes.c_int, [_VP, _VP, _VP, _VP, _I, _I, _VP, _SZ, _VP]),
    "vs_loss_spec2wav_backward": (ctypes.c_int, [_VP, _VP, _VP, _VP, _VP, _I, _I, _VP, _SZ, _VP]),
    "vs_sisnr_loss": (ctypes.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _VP, _SZ, _VP]),
    "vs_encoder_configure": (ctypes.c_int, [_VP, ctypes.POINTER(VsEncoderDims), _VP]),
    "vs_encoder_load_params": (ctypes.c_int, [_VP, ctypes.POINTER(VsEncoderParams), _VP]),
    "vs_encoder_workspace_bytes": (_SZ, [_VP, _I, _I, _I]),
    "vs_encoder_mel": (ctypes.c_int, [_VP, _VP, _VP, _I, _I, _VP, _SZ, _VP]),
    "vs_encoder_forward": (ctypes.c_int, [_VP, _VP, _VP, _I, _I, _VP, _SZ, _VP]),
    "vs_encoder_dvector": (ctypes.c_int, [_VP, _VP, _VP, _I, _I, _VP, _SZ, _VP]),
    "vs_sisnr_wav": (ctypes.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _VP]),
    "vs_sdr_workspace_bytes": (_SZ, [_I, _I]),
    "vs_sdr": (ctypes.c_int, [_VP, _VP, _VP, _VP, _I, _I, _VP, _SZ, _VP]),
    "vs_conv_stack": (ctypes.c_int, [_VP, _VP, _VP, _I, _I, _I, _VP, _SZ, _VP]),
    "vs_debug_conv_layer": (ctypes.c_int, [_VP, _I, _VP, _VP, _I, _I, _I, _VP]),
    "vs_debug_lstm_head": (ctypes.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _VP]),
    "vs_debug_lstm_timing": (ctypes.c_int, [_VP, ctypes.POINTER(ctypes.c_int64)]),
    "vs_last_launch_count": (ctypes.c_int, [_VP]),
    "vs_engine_set_profiling": (ctypes.c_int, [_VP, _I]),
    "vs_profile_read": (ctypes.c_int, [_VP, _I, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_float)]),
}

_lib = None


def load():
    """Load the shared library and bind every declared symbol (raises if anything is missing)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m voicesplit_b200.build` "
                "(there is no CPU or PyTorch fallback for the mask path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class VsError(RuntimeError):
    pass


def check(rc, what):
    if rc != VS_OK:
        msg = load().vs_last_error()
        raise VsError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
