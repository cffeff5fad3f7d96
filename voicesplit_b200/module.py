"""nn.Module front end of the B200 mask-estimation engine.

Mirrors the reference's module contract so that the reference's train.py /
test.py can import `models.voicesplit.model.VoiceSplit` / `models.voicefilter.model.VoiceFilter`
from this repo unchanged:

* constructor takes the config object and reads `config.audio[backend]['num_freq']` and
  `config.model[...]` exactly like /root/reference/models/voicesplit/model.py:10-13,58-64;
* `state_dict()` has the reference's keys, shapes and dtypes (the `conv.N` indices are the
  positions inside an nn.Sequential, so placeholders sit where the reference has ZeroPad2d and
  the activation);
* `forward(x[B,T,F], speaker_embedding[B,E]) -> mask[B,T,F]` (model.py:66-89).

The layers below only *hold* parameters - none of them is ever called.  The arithmetic runs in
hand-written sm_100a kernels behind the C ABI (include/voicesplit_b200.h), in eval mode (tensor-core
precision modes) and in train mode (fp32 kernels, autograd through vs_train_forward/backward).  There
is no CPU path and no PyTorch fallback: forward raises if the inputs are not on a CUDA device.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import synth
from .engine import MaskEngine


class _Slot(nn.Module):
    """Parameter-free placeholder keeping the reference's nn.Sequential indices (pad / activation)."""

    def __init__(self, what):
        super().__init__()
        self.what = what

    def extra_repr(self):
        return self.what

    def forward(self, *_):  # pragma: no cover - never called
        raise RuntimeError("placeholder module: the conv stack runs inside the CUDA engine")


def _conv_stack_modules(activation):
    mods = []
    for _idx, cin, cout, kh, kw, dil in synth.CONV_LAYERS:
        if kh * kw > 1:
            pad_t, pad_f = (kh - 1) // 2 * dil, (kw - 1) // 2
            mods.append(_Slot(f"zero pad T+-{pad_t} F+-{pad_f}"))
        mods.append(nn.Conv2d(cin, cout, kernel_size=(kh, kw), dilation=(dil, 1)))
        mods.append(nn.BatchNorm2d(cout))
        mods.append(_Slot(activation))
    return mods


class _MaskTrainFn(torch.autograd.Function):
    """Training-mode forward/backward through the engine (fp32 kernels, batch-statistics BatchNorm)."""

    @staticmethod
    def forward(ctx, module, x, emb, *params):
        eng = module._sync_engine(x.device)
        eng.set_train_tensor_cores(module.train_tensor_cores)
        buffers = {k: v for k, v in module.named_buffers()}
        mask, saved = eng.train_forward(x, emb, buffers, momentum=0.1)
        ctx.eng, ctx.saved = eng, saved
        ctx.shapes = {k: tuple(p.shape) for k, p in module.named_parameters()}
        ctx.save_for_backward(mask)
        return mask

    @staticmethod
    def backward(ctx, grad_mask):
        (mask,) = ctx.saved_tensors
        grads, gemb = ctx.eng.train_backward(ctx.saved, mask, grad_mask, ctx.shapes)
        return (None, None, gemb) + tuple(grads[k] for k in ctx.eng.PARAM_ORDER)


class MaskEstimator(nn.Module):
    ACTIVATION = "mish"

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.audio = self.config.audio[self.config.audio["backend"]]
        m = self.config.model
        self.dims = synth.make_dims(self.audio["num_freq"], m["emb_dim"], m["lstm_dim"], m["fc1_dim"], m["fc2_dim"])
        self.conv = nn.Sequential(*_conv_stack_modules(self.ACTIVATION))
        assert [i for i, mod in enumerate(self.conv) if isinstance(mod, nn.Conv2d)] == [c[0] for c in synth.CONV_LAYERS]
        self.lstm = nn.LSTM(8 * self.dims["num_freq"] + self.dims["emb_dim"], self.dims["lstm_dim"],
                            batch_first=True, bidirectional=True)
        self.fc1 = nn.Linear(2 * self.dims["lstm_dim"], self.dims["fc1_dim"])
        self.fc2 = nn.Linear(self.dims["fc1_dim"], self.dims["fc2_dim"])
        # arithmetic of the contractions: "fp16x3" / "bf16x3" (fp32-faithful split operands on tensor
        # cores), "fp16" / "bf16" (single pass, fast) or "fp32" (CUDA cores); see include/voicesplit_b200.h
        self.precision = os.environ.get("VOICESPLIT_PRECISION", "fp16x3")
        # training: conv forward / data gradient on tensor cores (fp16x3 / bf16x3); False = fp32 CUDA cores
        self.train_tensor_cores = os.environ.get("VOICESPLIT_TRAIN_FP32", "0") != "1"
        self._engine = None
        self._packed_sig = None

    # the engine handle is process-local: drop it when the module is pickled / deep-copied (it is rebuilt lazily)
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engine"] = None
        state["_packed_sig"] = None
        return state

    # ---- engine plumbing ----------------------------------------------------------------------
    def _signature(self):
        sig = []
        for t in list(self.parameters()) + list(self.buffers()):
            sig.append((t.data_ptr(), t._version))
        return tuple(sig)

    def _sync_engine(self, device):
        if self._engine is None or self._engine.device != device:
            with torch.cuda.device(device):
                self._engine = MaskEngine(activation=self.ACTIVATION, device=device, **self.dims)
            self._packed_sig = None
        sig = self._signature()
        if sig != self._packed_sig:  # optimizer step / load_state_dict / .cuda() invalidate the packing
            self._engine.load_state_dict_tensors({k: v for k, v in self.state_dict().items()
                                                  if v.dtype == torch.float32})
            self._packed_sig = sig
        return self._engine

    def engine(self, device=None):
        """The MaskEngine behind this module on `device` (default: where the parameters live), parameters packed and current.
        Other engine services - the audio front end, the loss chain (losses.SpecSiSNRLoss) - attach to it."""
        device = torch.device(device) if device is not None else next(self.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("voicesplit_b200 runs on sm_100a CUDA devices only: move the module with .cuda() first")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        return self._sync_engine(device)

    def _guard(self, x):
        if not x.is_cuda:
            raise RuntimeError("voicesplit_b200 runs on sm_100a CUDA devices only: move the module and inputs "
                               "with .cuda() (there is no CPU or PyTorch fallback)")

    # ---- the reference contract ---------------------------------------------------------------
    def forward(self, x, speaker_embedding):
        """x: [B, T, num_freq], speaker_embedding: [B, emb_dim] -> mask [B, T, num_freq]."""
        self._guard(x)
        if self.training:
            # train.py:84,94: BatchNorm uses batch statistics and updates its running buffers; the output
            # carries a grad_fn whose backward fills .grad of every parameter (fp32 kernels)
            if x.requires_grad:
                raise NotImplementedError("the gradient w.r.t. the input spectrogram is not provided")
            names = dict(self.named_parameters())
            params = [names[k] for k in MaskEngine.PARAM_ORDER]
            return _MaskTrainFn.apply(self, x, speaker_embedding.to(x.device), *params)
        eng = self._sync_engine(x.device)
        return eng.forward(x, speaker_embedding.to(x.device), precision=self.precision)

    @torch.no_grad()
    def forward_masked(self, x, speaker_embedding):
        """mask and mask*x (the caller-side apply of reference train.py:95) from one fused launch."""
        self._guard(x)
        eng = self._sync_engine(x.device)
        return eng.forward(x, speaker_embedding.to(x.device), precision=self.precision, want_masked=True)
