"""Training-step benchmark (BASELINE config 4 shape): forward (batch-stat BatchNorm) + the reference's loss chain
(both spectrograms through the differentiable iSTFT, then Si-SNR: train.py:95-108, one engine call) + backward + ONE
flat NCCL gradient all-reduce + Adam step, data-parallel over the ranks it is launched on.  --loss flat applies the
general-C torch criterion to the flattened spectrograms instead (no iSTFT), the round-1 early variant.

    python tools/train_bench.py --batch 8 --steps 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_bench.py --batch 8

Prints one JSON line (rank 0).  Not the headline metric (bench.py is); numbers go to DESIGN.md section 7.
BatchNorm statistics are per rank (what "a single NCCL all-reduce" implies)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voicesplit_b200 import config, dist as vdist, synth  # noqa: E402
from voicesplit_b200.losses import SpecSiSNRLoss, si_snr_with_pit  # noqa: E402


def eager_loss_chain(est, tgt, phase, lens, n_fft=1200, hop=160, win=400, min_db=-100.0, ref_db=20.0):
    """The reference's loss chain as stock torch ops on the GPU (torch_spec2wav with torch.istft standing in for the removed
    torchaudio.functional.istft, then the C = 1 criterion) - the eager baseline the fused engine call is timed against."""
    window = torch.hann_window(win, periodic=False, device=est.device)

    def spec2wav(spec):
        S = (torch.clamp(spec, 0.0, 1.0) - 1.0) * -min_db + ref_db
        mag = torch.pow(10.0, S * 0.05).transpose(2, 1)
        ph = phase.transpose(2, 1)
        z = torch.complex(mag * torch.exp(ph.cos()), mag * torch.exp(ph.sin()))
        return torch.istft(z, n_fft, hop_length=hop, win_length=win, window=window, center=True)
    B = est.shape[0]
    return si_snr_with_pit(spec2wav(est).view(B, 1, -1), spec2wav(tgt).view(B, 1, -1), lens.view(-1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8, help="utterances per GPU")
    ap.add_argument("--frames", type=int, default=301)
    ap.add_argument("--freq", type=int, default=601)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--loss", choices=("istft", "flat"), default="istft")
    ap.add_argument("--cpu-reference", action="store_true", help="also time PyTorch autograd on the host CPU (B=2)")
    args = ap.parse_args()
    rank, world, local = vdist.env_rank()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = vdist.init("nccl", dev)
    from models.voicesplit.model import VoiceSplit
    dims = synth.make_dims(args.freq, 256, 400, 600)
    model = VoiceSplit(config.AttrDict(synth.make_config_dict(dims)))
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_state_dict(dims, 0, "default").items()})
    model = model.to(dev).train().enable_data_parallel(dist, sync_bn=False, overlap=True)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    B, T, F = args.batch, args.frames, args.freq
    x, emb = synth.make_inputs(B, T, dims, 100 + rank)
    x, emb = torch.from_numpy(x).to(dev), torch.from_numpy(emb).to(dev)
    target = torch.rand(B, T, F, device=dev) * x
    lengths = torch.full((B,), T * F, device=dev)
    phase = (torch.rand(B, T, F, device=dev) * 2 - 1) * np.pi
    seq_len = torch.full((B, 1), 160 * (T - 1), device=dev, dtype=torch.int64)
    crit = SpecSiSNRLoss(model.engine(dev), dict(n_fft=2 * (F - 1), hop_length=160, win_length=400)) if args.loss == "istft" else None

    def criterion(mask):
        if crit is not None:
            return crit(mask * x, target, phase, seq_len)
        return si_snr_with_pit((mask * x).view(B, 1, -1), target.view(B, 1, -1), lengths)

    sections = {}

    def mark():
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    def step(timed=False):
        t0 = mark()
        opt.zero_grad(set_to_none=True)
        mask = model(x, emb)
        t1 = mark()
        loss = criterion(mask)
        loss.backward()
        t2 = mark()
        n = vdist.allreduce_gradients(model, dist)          # flat-buffer fast path (MaskEstimator.enable_data_parallel)
        t3 = mark()
        opt.step()
        t4 = mark()
        if timed:
            torch.cuda.synchronize()
            for name, a, b in (("forward", t0, t1), ("loss+backward", t1, t2), ("allreduce", t2, t3), ("adam", t3, t4)):
                sections[name] = sections.get(name, 0.0) + a.elapsed_time(b)
        return loss, n

    for _ in range(args.warmup):
        step()
    # per-kernel breakdown of one step (forward and backward are separate engine calls)
    eng = model._engine
    eng.set_profiling(True)
    opt.zero_grad(set_to_none=True)
    mask = model(x, emb)
    fwd = eng.profile_read()
    criterion(mask).backward()
    bwd = eng.profile_read()
    eng.set_profiling(False)
    breakdown = {}
    for tag, rows in (("fwd", fwd), ("bwd", bwd)):
        for name, ms in rows:
            breakdown[f"{tag}:{name}"] = round(breakdown.get(f"{tag}:{name}", 0.0) + ms, 3)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    losses = []
    for _ in range(args.steps):
        loss, nred = step()
        losses.append(loss.detach())
    losses = [float(v) for v in losses]
    e1.record()
    torch.cuda.synchronize()
    thr, ms = vdist.aggregate_throughput(B * args.steps, e0.elapsed_time(e1), dist, dev)
    for _ in range(2):
        step(timed=True)
    if rank == 0:
        chain = None
        if crit is not None:       # the loss chain alone: fused engine call vs the same chain as stock torch ops on this GPU
            est = (mask.detach() * x).requires_grad_(True)

            def t_ms(fn, n=5):
                fn(); torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(n):
                    fn()
                b.record(); torch.cuda.synchronize()
                return a.elapsed_time(b) / n
            fused = t_ms(lambda: crit(est, target, phase, seq_len).backward())
            eager = t_ms(lambda: eager_loss_chain(est, target, phase, seq_len).backward())
            est.grad = None
            l_f = crit(est, target, phase, seq_len); l_f.backward(); g_f = est.grad.clone(); est.grad = None
            l_e = eager_loss_chain(est, target, phase, seq_len); l_e.backward()
            chain = {"fused_ms": round(fused, 3), "eager_torch_ms": round(eager, 3), "loss_fused": float(l_f.detach()), "loss_eager": float(l_e.detach()),
                     "grad_max_rel_diff": float((g_f - est.grad).abs().max() / est.grad.abs().max())}
        out = {"metric": "training utterances/s (forward + iSTFT/Si-SNR loss chain + backward + grad all-reduce + Adam)"
               if crit is not None else "training utterances/s (forward + Si-SNR-PIT + backward + grad all-reduce + Adam)", "value": thr,
               "loss": args.loss, "loss_chain": chain,
               "n_gpus": world, "per_gpu_batch": B, "frames": T, "freq_bins": F, "ms_per_step": ms / args.steps,
               "allreduce_floats": nred, "losses": losses, "arithmetic": "conv forward / data gradient / weight gradient, LSTM input GEMMs and iSTFT GEMMs on tcgen05 (fp16x3 activations, bf16x3 gradients); "
                                                                        "BatchNorm, LSTM recurrence and head gradients fp32 CUDA cores",
               "bn_statistics": "per rank", "kernel_ms": breakdown,
               "section_ms": {k: round(v / 2, 2) for k, v in sections.items()}}
        if args.cpu_reference:
            from oracle import torch_port
            sd = {k: torch.from_numpy(np.array(v)) for k, v in synth.make_state_dict(dims, 0, "default").items()}
            for k, v in sd.items():
                if v.dtype == torch.float32 and "running" not in k:
                    v.requires_grad_(True)
            xc, ec = synth.make_inputs(2, T, dims, 100)
            xc, ec = torch.from_numpy(xc), torch.from_numpy(ec)
            t0 = time.perf_counter()
            m = torch_port.forward_train(sd, xc, ec)
            si_snr_with_pit((m * xc).view(2, 1, -1), (xc * 0.5).view(2, 1, -1), torch.full((2,), T * F)).backward()
            dt = time.perf_counter() - t0
            out["cpu_reference"] = {"value": 2 / dt, "unit": "utterances/s", "sample": f"B=2 forward+backward via torch autograd (oracle/torch_port.forward_train), {torch.get_num_threads()} threads, {dt:.1f} s"}
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
