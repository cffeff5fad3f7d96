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


class _TrainCtx:
    """What the backward needs from a training forward: the engine workspace is consumed by the backward (gate activations and
    dX are overwritten in place), so a second backward through the same graph must raise, not return garbage."""

    def __init__(self, eng, saved):
        self.eng, self.saved = eng, saved

    def take(self):
        if self.saved is None:
            raise RuntimeError("voicesplit_b200: this graph's saved activations were consumed by a previous backward "
                               "(retain_graph / a second backward through the same forward is not supported)")
        saved, self.saved = self.saved, None
        return saved


class _MaskTrainFn(torch.autograd.Function):
    """Training-mode forward/backward through the engine (batch-statistics BatchNorm, full backward)."""

    @staticmethod
    def forward(ctx, module, x, emb, *params):
        eng = module._sync_engine(x.device)
        eng.set_train_tensor_cores(module.train_tensor_cores)
        module._install_dp_hooks(eng)
        momentum, track = module._bn_train_config()
        buffers = {k: v for k, v in module.named_buffers()} if track else None
        mask, saved = eng.train_forward(x, emb, buffers, momentum=momentum)
        if track:
            # the running statistics were updated through raw pointers (no tensor version bump): the eval-mode BatchNorm fold
            # held by the engine is stale even if no optimizer step follows
            module._packed_sig = None
        ctx.module, ctx.state = module, _TrainCtx(eng, saved)
        ctx.shapes = {k: tuple(p.shape) for k, p in module.named_parameters()}
        ctx.save_for_backward(mask)
        return mask

    @staticmethod
    def backward(ctx, grad_mask):
        (mask,) = ctx.saved_tensors
        module, eng = ctx.module, ctx.state.eng
        saved = ctx.state.take()
        flat, overlap_ok = module._grad_buffer(ctx.shapes, mask.device)
        module._dp_backward_begin(flat if overlap_ok else None)
        grads, gemb, gx = eng.train_backward(saved, mask, grad_mask, ctx.shapes, flat=flat, want_grad_x=ctx.needs_input_grad[1])
        out = tuple(grads[k] for k in eng.PARAM_ORDER)
        del grads          # no other reference to the views: autograd then adopts them as .grad without a copy
        return (None, gx, gemb if ctx.needs_input_grad[2] else None) + out


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
        # data parallel (enable_data_parallel): process-group module, SyncBN flag, flat gradient buffers
        self._dp = None
        self.sync_bn = False
        self._dp_overlap = False
        self._flat = [None, None]
        self._flat_active = None
        self._dp_pending = None

    # the engine handle is process-local: drop it when the module is pickled / deep-copied (it is rebuilt lazily)
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engine"] = None
        state["_packed_sig"] = None
        state["_dp"] = None
        state["_flat"] = [None, None]
        state["_flat_active"] = None
        state["_dp_pending"] = None
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

    # ---- training plumbing ----------------------------------------------------------------------
    def _bn_train_config(self):
        """(momentum, update running statistics?) from the BatchNorm2d holders, like the reference modules would use them.
        Settings the kernels do not implement raise instead of being silently ignored."""
        bns = [m for m in self.conv if isinstance(m, nn.BatchNorm2d)]
        moms = {m.momentum for m in bns}
        if len(moms) != 1 or None in moms:
            raise NotImplementedError("all BatchNorm2d layers must share one numeric momentum (cumulative averaging is not supported)")
        if any(abs(m.eps - 1e-5) > 1e-12 for m in bns):
            raise NotImplementedError("BatchNorm2d eps other than 1e-5 is not supported by the fused kernels")
        if any(not m.training for m in bns) or any(not m.affine for m in bns):
            raise NotImplementedError("partially frozen BatchNorm (a sub-module in eval mode while the model trains) is not supported: "
                                      "call model.eval() for running statistics or model.train() for batch statistics")
        track = {m.track_running_stats for m in bns}
        if len(track) != 1:
            raise NotImplementedError("track_running_stats must be the same for every BatchNorm2d")
        return float(moms.pop()), bool(track.pop())

    def enable_data_parallel(self, dist, sync_bn=False, overlap=True):
        """Data-parallel training (BASELINE config 4): `dist` is torch.distributed with an initialised default group (or None).
        Gradients land in ONE flat fp32 buffer (parameters' .grad are views of it), `voicesplit_b200.dist.allreduce_gradients`
        reduces it with a single collective; with overlap=True the LSTM / FC tail (97 % of the bytes) starts reducing while
        the conv-stack backward is still running.  sync_bn=True makes every BatchNorm use the statistics of the concatenated
        global batch (one 1 KB all-reduce per layer and direction) - exactly the reference's single-process semantics."""
        self._dp = dist
        self.sync_bn = bool(sync_bn)
        self._dp_overlap = bool(overlap)
        self._hooks_for = None
        return self

    def _install_dp_hooks(self, eng):
        key = (id(eng), id(self._dp), self.sync_bn, self._dp_overlap)
        if getattr(self, "_hooks_for", None) == key:
            return
        dist = self._dp
        world = dist.get_world_size() if dist is not None else 1
        if dist is not None and self.sync_bn and world > 1:
            eng.set_sync_bn(lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM), world)
        else:
            eng.set_sync_bn(None, 1)
        eng.set_backward_hook(self._dp_stage if (dist is not None and self._dp_overlap and world > 1) else None)
        self._hooks_for = key

    def _grad_buffer(self, shapes, device):
        """One of two flat gradient buffers: the one no current .grad aliases (autograd ADDS the returned gradient to an
        existing .grad, which must therefore live elsewhere).  Second value: True when no .grad exists yet, i.e. the buffer
        holds the final gradient as soon as the kernels have run (the condition for the overlapped all-reduce)."""
        from .engine import MaskEngine
        _, total = MaskEngine.grad_layout(shapes)
        live = [p.grad for p in self.parameters() if p.grad is not None]
        pick = 0
        for i in (0, 1):
            f = self._flat[i]
            if f is None or f.numel() != total or f.device != device:
                pick = i
                self._flat[i] = torch.empty(total, dtype=torch.float32, device=device)
                break
            lo, hi = f.data_ptr(), f.data_ptr() + f.numel() * 4
            if not any(lo <= g.data_ptr() < hi for g in live):
                pick = i
                break
        self._flat_active = self._flat[pick]
        return self._flat_active, not live

    def _dp_backward_begin(self, flat):
        self._dp_flat_for_hook = flat
        self._dp_pending = None

    def _dp_stage(self, stage):
        """Engine callback in the middle of vs_train_backward: the LSTM / FC gradients (the tail of the flat buffer) are
        enqueued on the current stream - start their all-reduce; NCCL runs it on its own stream, next to the conv backward."""
        flat, dist = self._dp_flat_for_hook, self._dp
        if stage != 1 or flat is None or dist is None:
            return
        from .dist import reduce_flat
        self._dp_pending = (flat, self.grad_tail_offset(), reduce_flat(flat[self.grad_tail_offset():], dist, async_op=True))

    def grad_tail_offset(self):
        """First element of the LSTM / FC part of the flat gradient buffer (everything before it is conv / BatchNorm)."""
        from .engine import MaskEngine
        lay, _ = MaskEngine.grad_layout({k: tuple(p.shape) for k, p in self.named_parameters()})
        return lay["lstm.weight_ih_l0"][0]

    def flat_gradient(self):
        """The flat fp32 buffer all current .grad tensors are views of, or None (e.g. gradients were accumulated or replaced)."""
        from .engine import MaskEngine
        names = dict(self.named_parameters())
        lay, total = MaskEngine.grad_layout({k: tuple(p.shape) for k, p in names.items()})
        for f in self._flat:       # either buffer: an accumulated gradient stays in the buffer of the first backward
            if f is None or f.numel() != total:
                continue
            base = f.data_ptr()
            if all(names[k].grad is not None and names[k].grad.data_ptr() == base + 4 * o and names[k].grad.is_contiguous()
                   for k, (o, _n) in lay.items()):
                return f
        return None

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
            # carries a grad_fn whose backward fills .grad of every parameter (and of x / the d-vector if they require it)
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
