"""Multi-GPU plumbing of the mask path: utterances are independent in the forward pass (eval-mode
BatchNorm uses running statistics, LSTM state is per utterance), so a batch is sharded across ranks
with NO data-path collective.  torch.distributed is used only to launch one
process per GPU, to barrier around the timed region and to take the max of the per-rank timings."""
from __future__ import annotations

import os

import torch


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend, device=None):
    """Initialise the default process group when WORLD_SIZE > 1; returns torch.distributed or None."""
    _, world, _ = env_rank()
    if world <= 1:
        return None
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return dist


def shard(global_batch, rank, world):
    """[start, stop) of the utterances rank `rank` owns when `global_batch` utterances are split as
    evenly as possible (the first `global_batch % world` ranks take one more)."""
    base, extra = divmod(global_batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_over_ranks(value, dist, device="cpu"):
    if dist is None:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, dist, device="cpu"):
    if dist is None:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(local_units, local_ms, dist, device="cpu"):
    """Whole-job throughput = units processed by all ranks / max-over-ranks time."""
    total = sum_over_ranks(local_units, dist, device)
    ms = max_over_ranks(local_ms, dist, device)
    return total / (ms / 1e3), ms


def reduce_flat(flat, dist, async_op=False):
    """Average `flat` over the ranks with ONE collective: NCCL averages inside the reduction (ReduceOp.AVG); backends without
    AVG (gloo) sum and divide.  Returns the work handle when async_op (its .wait() orders the current stream behind it)."""
    if dist.get_backend() == "nccl":
        return dist.all_reduce(flat, op=dist.ReduceOp.AVG, async_op=async_op)
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=async_op)
    if async_op:
        class _Div:
            def wait(self_inner):
                work.wait()
                flat.div_(dist.get_world_size())
        return _Div()
    flat.div_(dist.get_world_size())
    return None


def allreduce_gradients(module_or_parameters, dist, world=None):
    """Data-parallel step of BASELINE config 4: ONE all-reduce over a flat fp32 buffer holding every parameter gradient
    (18.9 M floats = 75.5 MB at the shipped config), averaged over the ranks (the loss is a batch mean).

    Fast path - a MaskEstimator whose .grad tensors are views of its flat gradient buffer (what its backward produces):
    the buffer is reduced in place, no gather and no copy-back; if `enable_data_parallel(..., overlap=True)` started the
    LSTM / FC tail during the backward, only the conv / BatchNorm head (2 MB) is reduced here and the tail is awaited.
    Generic path (any iterable of parameters): gather into a temporary flat buffer, reduce, scatter back.
    Returns the number of gradient elements reduced.  No-op for a single process."""
    if dist is None:
        return 0
    flat_of = getattr(module_or_parameters, "flat_gradient", None)
    if flat_of is not None:
        module = module_or_parameters
        flat = flat_of()
        if flat is not None:
            pending, module._dp_pending = module._dp_pending, None
            if pending is not None and pending[0] is flat:
                _f, tail, work = pending
                reduce_flat(flat[:tail], dist)
                work.wait()
            else:
                if pending is not None:
                    pending[2].wait()
                reduce_flat(flat, dist)
            return flat.numel()
        params = [p for p in module.parameters() if p.grad is not None]
    else:
        params = [p for p in module_or_parameters if p.grad is not None]
    if not params:
        return 0
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    reduce_flat(flat, dist)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat.numel()
