"""bench.py - utterances/s of the speaker-conditioned mask-estimation forward pass on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp16x3|bf16x3|fp16|bf16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (CNN -> BiLSTM -> FC -> sigmoid mask -> mask * spectrogram)
over one batch of synthetic utterances.  Workload (BASELINE.json): `--batch` utterances per GPU of
601 frames x 257 bins + a random 256-d d-vector, random-init ("stress" flavour) weights of the
reference architecture.  Utterances are independent, so the batch is sharded across ranks with no
data-path collective (weak scaling; the only collective is the max-over-ranks of the timings).

One JSON line is printed by rank 0; keys are documented in DESIGN.md ("Measurement").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from voicesplit_b200 import synth  # noqa: E402

# algorithmic forward FLOPs (2 x MAC) per utterance, SURVEY.md section 8(d)
def flops_per_utt(T, F, E=256, H=400, N1=600):
    P = T * F
    conv = 2 * P * (64 * 7 + 64 * 64 * 7 + 5 * 64 * 64 * 25 + 64 * 8)
    lstm = 2 * T * (8 * F * 8 * H) + 2 * (E * 8 * H) + 2 * T * 2 * (H * 4 * H)
    fc = 2 * T * (2 * H * N1 + N1 * F)
    return dict(conv=conv, conv5x5_layer=2 * P * 64 * 64 * 25, lstm=lstm, fc=fc, total=conv + lstm + fc)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d["bf16_tflops_sustained"],
                    source="MEASURED_PEAKS.json (of measured)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="B200_PROFILING.md fallback (of fallback)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def _usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_throughput(dims, T, seconds_budget=20.0, steps=1, warmup=1):
    """The reference's CPU implementation of the path (the same torch.nn ATen ops the reference module
    issues, restated in oracle/torch_port.py; /root/reference itself does not exist on the GPU box)
    on a bounded sample of the workload: B_s utterances of the same T x F.  The thread count is the
    best of a short sweep (all cores is not always fastest for oneDNN on a many-core host)."""
    from oracle import torch_port
    cores = _usable_cores()
    sd = {k: torch.from_numpy(v) for k, v in synth.make_state_dict(dims, 0, "stress").items() if v.dtype == np.float32}
    xs, es = synth.make_inputs(1, max(16, T // 8), dims, 98)           # short probe for the thread sweep
    xs, es = torch.from_numpy(xs), torch.from_numpy(es)
    best, best_t = None, cores
    for nt in sorted({cores, max(1, cores // 2), min(cores, 32), min(cores, 16), min(cores, 8)}, reverse=True):
        torch.set_num_threads(nt)
        torch_port.forward(sd, xs, es)
        t0 = time.perf_counter()
        torch_port.forward(sd, xs, es)
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, best_t = dt, nt
    torch.set_num_threads(best_t)
    x, emb = synth.make_inputs(1, T, dims, 99)
    xt, et = torch.from_numpy(x), torch.from_numpy(emb)
    t0 = time.perf_counter()
    torch_port.forward(sd, xt, et)                      # warm-up (also sizes the sample)
    one = time.perf_counter() - t0
    bs = int(max(1, min(8, (seconds_budget / max(steps + warmup - 1, 1)) // max(one, 1e-3))))
    x, emb = synth.make_inputs(bs, T, dims, 99)
    xt, et = torch.from_numpy(x), torch.from_numpy(emb)
    for _ in range(max(warmup - 1, 0)):
        torch_port.forward(sd, xt, et)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        torch_port.forward(sd, xt, et)
        times.append(time.perf_counter() - t0)
    dt = float(np.sum(times))
    return dict(value=bs * steps / dt, unit="utterances/s", cores=best_t, kind="port",
                sample=f"{bs} utterance(s) x {steps} step(s) of {T}x{dims['num_freq']} through oracle/torch_port.py "
                       f"(the reference's own torch.nn CPU ops, fp32, {best_t} of {cores} host threads: best of a sweep), "
                       f"{dt:.1f} s"), dt / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("VOICESPLIT_PRECISION", "fp16x3"))
    ap.add_argument("--batch", type=int, default=256, help="utterances per GPU")
    ap.add_argument("--frames", type=int, default=601)
    ap.add_argument("--freq", type=int, default=257)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary config-4 / config-5 measurements")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dims = synth.make_dims(args.freq, 256, 400, 600)
    T, F, B = args.frames, args.freq, args.batch
    fl = flops_per_utt(T, F)
    workload = f"full forward CNN+BiLSTM+FC+mask apply, {B} utt/GPU x {T} frames x {F} bins + 256-d d-vector (BASELINE configs[2] shape)"
    config = {"workload": workload, "per_gpu_batch": B, "frames": T, "freq_bins": F, "global_batch": B * world,
              "parallelism": f"utterance-sharded x{world}, no data-path collective",
              "l2_policy": "inputs larger than L2 (x is %.0f MB per step)" % (B * T * F * 4 / 1e6),
              "weights": "random-init stress flavour (synth.make_state_dict seed 0)"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        cb, step_s = cpu_reference_throughput(dims, T, seconds_budget=60.0, steps=max(args.steps, 1), warmup=max(args.warmup, 1))
        line = {"impl": "reference", "metric": "utterances/s (601-frame, 257-bin spectrogram) masked", "value": cb["value"],
                "unit": "utterances/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    from voicesplit_b200 import dist as vdist
    from voicesplit_b200.engine import MaskEngine
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = vdist.init("nccl", dev)

    eng = MaskEngine(activation="mish", device=dev, **dims)
    sd = synth.make_state_dict(dims, 0, "stress")
    eng.load_state_dict_tensors({k: torch.from_numpy(v).to(dev) for k, v in sd.items() if v.dtype == np.float32})
    xh, eh = synth.make_inputs(B, T, dims, 1234 + rank)
    xh, eh = torch.from_numpy(xh).pin_memory(), torch.from_numpy(eh).pin_memory()
    x, emb = xh.to(dev), eh.to(dev)
    mask_h, masked_h = torch.empty_like(xh).pin_memory(), torch.empty_like(xh).pin_memory()
    prec = args.precision

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        barrier()
        return ms

    # ---- parity evidence: the timed precision against this repo's fp32 CUDA-core path (itself pinned to
    # the oracle/golden vectors by tests/) on the first two utterances of the bench batch
    nb = min(2, B)
    ref32 = eng.forward(x[:nb], emb[:nb], precision="fp32")
    got = eng.forward(x[:nb], emb[:nb], precision=prec)
    parity = {"vs": "fp32 CUDA-core path, first %d utterances" % nb, "mask_mae": float((got - ref32).abs().mean()),
              "mask_max_abs": float((got - ref32).abs().max())}

    # ---- device-resident throughput ("value")
    for _ in range(args.warmup):
        eng.forward(x, emb, precision=prec, want_masked=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.set_profiling(True)
    local_ms = timed(lambda: eng.forward(x, emb, precision=prec, want_masked=True), args.steps)
    kernel_ms = {}
    for name, ms in eng.profile_read():           # per-kernel times of the last timed step
        kernel_ms[name] = kernel_ms.get(name, 0.0) + ms
    launches_per_step = eng.last_launch_count()
    eng.set_profiling(False)
    value, dev_ms = vdist.aggregate_throughput(B * args.steps, local_ms, dist, dev)
    # ---- end to end through the host-buffer plugin call ("e2e"): every step copies its inputs from pinned
    # host memory, runs the forward and copies mask + masked back.  The serving form of the call is
    # used (vs_forward_host_submit / _wait, two slots), so the copies of step i+1 / i-1 overlap the
    # compute of step i; the synchronous vs_forward_host is timed as well and reported next to it.
    for _ in range(min(args.warmup, 2)):
        eng.forward_host(xh, eh, mask_h, masked_h, precision=prec)
    local_sync = timed(lambda: eng.forward_host(xh, eh, mask_h, masked_h, precision=prec), max(2, args.steps // 2))
    sync_value, _ = vdist.aggregate_throughput(B * max(2, args.steps // 2), local_sync, dist, dev)
    slots = [(xh, eh, mask_h, masked_h),
             (xh.clone().pin_memory(), eh.clone().pin_memory(), torch.empty_like(xh).pin_memory(), torch.empty_like(xh).pin_memory())]

    def pipelined(steps):
        for i in range(steps):
            eng.host_submit(i & 1, *slots[i & 1], precision=prec)
            if i > 0:
                eng.host_wait((i - 1) & 1)
        eng.host_wait((steps - 1) & 1)

    pipelined(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pipelined(args.steps)           # the last host_wait blocks until the last D2H has landed
    e1.record()
    torch.cuda.synchronize()
    local_e2e = e0.elapsed_time(e1)
    barrier()
    e2e_value, e2e_ms = vdist.aggregate_throughput(B * args.steps, local_e2e, dist, dev)
    clocks = sampler.stop() if rank == 0 else None
    # ---- the single-pass fast mode, reported next to its measured error (never as the headline)
    fast = None
    if prec in ("fp16x3", "bf16x3", "fp16_f8c"):
        fp = prec[:4]
        gotf = eng.forward(x[:nb], emb[:nb], precision=fp)
        for _ in range(2):
            eng.forward(x, emb, precision=fp, want_masked=True)
        fsteps = max(2, args.steps // 2)
        fms = timed(lambda: eng.forward(x, emb, precision=fp, want_masked=True), fsteps)
        fval, _ = vdist.aggregate_throughput(B * fsteps, fms, dist, dev)
        fast = {"precision": fp, "value": fval, "unit": "utterances/s", "mask_mae_vs_fp32_path": float((gotf - ref32).abs().mean()),
                "mask_max_abs_vs_fp32_path": float((gotf - ref32).abs().max()),
                "note": "single MMA pass, 11-bit (fp16) / 8-bit (bf16) operands; error measured on stress weights"}

    if rank == 0:
        peaks = load_peaks()
        # dominant kernel: the five 5x5 dilated conv layers (89.5 % of the algorithmic FLOPs)
        conv_ms = [kernel_ms.get(f"cnn{i}") for i in (3, 4, 5, 6, 7)]
        roof = None
        if all(v is not None for v in conv_ms):
            avg = float(np.mean(conv_ms))
            ach = fl["conv5x5_layer"] * B / (avg / 1e3) / 1e12
            passes = {"bf16x3": 3, "fp16x3": 3, "fp16_f8c": 2, "bf16": 1, "fp16": 1}.get(prec)
            peak = peaks["bf16_tflops_sustained"]
            traffic = None
            prof = os.path.join(ROOT, "profiles", "r01_conv_tc_summary.json")
            if os.path.exists(prof) and passes:
                pj = json.load(open(prof))
                if pj.get("precision") == prec and pj.get("frames") == T and pj.get("freq_bins") == F:
                    traffic = pj["dram_bytes_per_launch"] / pj["batch"] * B
            roof = {"bound": "tensor", "kernel": "k_conv_tc: dilated 5x5 conv 64->64 + BN + act (cnn3..cnn7)", "achieved": ach, "peak": peak,
                    "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                    "traffic_source": "profiles/r01_conv_tc_summary.json (ncu --set full, scaled per utterance)" if traffic else None,
                    "peak_source": peaks["source"] + ", sustained 16-bit dense (cuBLAS bf16; fp16 runs on the same pipe)",
                    "avg_launch_ms": avg, "algorithmic_flops_per_launch": fl["conv5x5_layer"] * B,
                    "algorithmic_bytes_per_launch": B * T * padded_f(F) * 64 * 2 * (2 if passes >= 2 else 1) * 2 if passes else None,
                    "mma_passes": passes, "tensor_pipe_frac_incl_passes": (ach * passes / peak) if passes else None,
                    # what the tensor pipe actually executes: algorithmic flops x passes x 6/5 (five filter taps occupy six M=128 slots)
                    "issued_tflops": (ach * passes * 1.2) if passes else None,
                    "issued_frac_of_peak": (ach * passes * 1.2 / peak) if passes else None,
                    "note": "fp32 mode runs on CUDA cores (no tensor pipe)" if prec == "fp32" else
                            "frac counts ALGORITHMIC flops; the faithful mode issues 3 MMA passes (hi*hi + lo*hi + hi*lo) and pads 5 taps to 6 slots"}
        line = {"metric": "utterances/s (601-frame, 257-bin spectrogram) masked", "value": value, "unit": "utterances/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"bf16x3": "bf16x3 (split-bf16 operands hi+lo, 3 MMAs, fp32 accumulate)",
                          "fp16x3": "fp16x3 (split-fp16 operands hi+lo, 3 MMAs, fp32 accumulate)",
                          "fp16_f8c": "fp16 + e4m3 correction (conv: 4 f16 + 4 f8f6f4 MMAs per tap pair, fp32 accumulate; LSTM/FC fp16x3)",
                          "bf16": "bf16", "fp16": "f16", "fp32": "f32"}[prec],
                "data": "synthetic", "config": config,
                "e2e": {"value": e2e_value, "unit": "utterances/s", "ms_per_step": e2e_ms / args.steps,
                        "api": "vs_forward_host_submit/_wait (2 slots, copies overlap the neighbouring step's compute)",
                        "synchronous_call_value": sync_value,
                        "h2d_bytes_per_step": int(xh.numel() * 4 + eh.numel() * 4),
                        "d2h_bytes_per_step": int(mask_h.numel() * 4 + masked_h.numel() * 4)},
                "gpu_launches": launches_per_step * args.steps * 2,   # device-resident + e2e timed regions
                "launches_per_step": launches_per_step,
                "clocks": clocks, "roofline": roof, "parity": parity, "fast_mode": fast,
                "kernel_ms_last_step": {k: round(v, 4) for k, v in kernel_ms.items()},
                "gflop_per_utt": {k: v / 1e9 for k, v in fl.items()},
                "tflops_total_algorithmic": fl["total"] * value / 1e12}
        if world == 1 and not args.no_cpu_baseline:
            cb, _ = cpu_reference_throughput(dims, T, seconds_budget=20.0)
            line["cpu_baseline"] = cb
        if world == 1 and not args.no_extras:
            # secondary measurements (never the headline): BASELINE config 4 (training step) and config 5 (waveform in,
            # separated waveform out) at the reference-native 301 x 601 shape; a failure here must not cost the headline
            try:
                del eng, x, emb, xh, eh, mask_h, masked_h, slots
                torch.cuda.empty_cache()
                line["extras"] = extra_measurements(dev)
            except Exception as ex:      # noqa: BLE001
                line["extras"] = {"error": repr(ex)[:300]}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def extra_measurements(dev):
    from voicesplit_b200 import config as vconfig
    from voicesplit_b200.engine import MaskEngine
    from voicesplit_b200.losses import SpecSiSNRLoss
    from models.voicesplit.model import VoiceSplit
    out = {}
    dims = synth.make_dims(601, 256, 400, 600)

    def timed(fn, n):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    # ---- config 4: forward (batch-stat BatchNorm) + loss chain of train.py:95-108 (both spectrograms through the Q1-faithful
    # differentiable iSTFT, then Si-SNR) + backward + Adam, B = 32, 301 x 601
    model = VoiceSplit(vconfig.AttrDict(synth.make_config_dict(dims)))
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_state_dict(dims, 0, "default").items()})
    model = model.to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    B, T, F = 32, 301, 601
    x, emb = synth.make_inputs(B, T, dims, 7)
    x, emb = torch.from_numpy(x).to(dev), torch.from_numpy(emb).to(dev)
    target = torch.rand(B, T, F, device=dev) * x
    phase = (torch.rand(B, T, F, device=dev) * 2 - 1) * np.pi
    seq_len = torch.full((B, 1), 160 * (T - 1), device=dev, dtype=torch.int64)
    crit = SpecSiSNRLoss(model.engine(dev), dict(n_fft=1200, hop_length=160, win_length=400), "q1")

    def step():
        opt.zero_grad(set_to_none=True)
        mask = model(x, emb)
        crit(mask * x, target, phase, seq_len).backward()
        opt.step()
    step(); step()
    ms = timed(step, 6)
    out["train_step_config4"] = {"value": B / (ms / 1e3), "unit": "utterances/s", "ms_per_step": ms, "per_gpu_batch": B, "frames": T, "freq_bins": F,
                                 "what": "forward (batch-stat BN) + differentiable iSTFT x2 + Si-SNR (one fused engine call) + backward + Adam; "
                                         "conv fwd/dgrad/wgrad, LSTM input GEMMs and the iSTFT GEMMs on tcgen05 (fp16x3/bf16x3), rest fp32"}
    del model, opt
    torch.cuda.empty_cache()
    # ---- config 5: 3 s @ 16 kHz waveform -> STFT -> mask -> iSTFT -> waveform, B = 64
    eng = MaskEngine(activation="mish", device=dev, **dims)
    eng.load_state_dict_tensors({k: torch.from_numpy(v).to(dev) for k, v in synth.make_state_dict(dims, 0, "stress").items() if v.dtype == np.float32})
    eng.configure_audio()
    Bw = 64
    wav = torch.randn(Bw, 48000, device=dev) * 0.05
    e2 = torch.randn(Bw, 256, device=dev)
    ms = timed(lambda: eng.separate(wav, e2), 3)
    out["audio_e2e_config5"] = {"value": Bw / (ms / 1e3), "unit": "utterances/s", "ms_per_step": ms, "batch": Bw, "samples": 48000,
                                "what": "waveform -> STFT (tcgen05 GEMM) -> CNN+BiLSTM+FC mask (fp16x3) -> mask*spec -> iSTFT (mixture phase) -> waveform"}
    # ---- d-vector extraction: 3 s reference clips -> GE2E embedding, B = 128
    from voicesplit_b200.speaker_encoder import SpeakerEncoder
    enc = SpeakerEncoder(engine=eng)
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_encoder_state_dict(1, "stress").items()})
    enc = enc.to(dev)
    ref_wav = torch.from_numpy(synth.make_reference_audio(128, 48000, 5)).to(dev)
    ms = timed(lambda: enc.embed_wav(ref_wav), 5)
    out["dvector_extract"] = {"value": 128 / (ms / 1e3), "unit": "utterances/s", "ms_per_step": ms, "batch": 128, "samples": 48000,
                              "what": "waveform -> |STFT|^2 -> 40 mel -> log10 -> 6 windows x 3 x LSTM(768) (fp16x3 tcgen05, persistent recurrent "
                                      "kernel) -> Linear(256) -> L2 normalise -> mean"}
    del eng, wav, e2, enc, ref_wav
    torch.cuda.empty_cache()
    # ---- the honest GPU bar: the reference's own stock torch ops (restated in oracle/torch_port.py; the
    # reference tree cannot travel) on the SAME B200 through cuDNN/cuBLAS eager, strict fp32 and TF32-allowed
    out["stock_torch_gpu_baseline"] = stock_torch_gpu_baseline(dev)
    return out


def stock_torch_gpu_baseline(dev, B=32, T=601):
    from oracle import torch_port
    dims = synth.make_dims(257, 256, 400, 600)
    sd = {k: torch.from_numpy(np.array(v)).to(dev) for k, v in synth.make_state_dict(dims, 0, "stress").items()}
    x, emb = synth.make_inputs(B, T, dims, 11)
    x, emb = torch.from_numpy(x).to(dev), torch.from_numpy(emb).to(dev)
    res = {"batch": B, "frames": T, "freq_bins": 257, "unit": "utterances/s",
           "what": "stock PyTorch eager (F.conv2d/batch_norm/Mish/LSTM/linear -> cuDNN/cuBLAS) on the same GPU, inputs resident"}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        for name, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            for _ in range(2):
                torch_port.forward(sd, x, emb)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                torch_port.forward(sd, x, emb)
            e1.record()
            torch.cuda.synchronize()
            res[name] = B * 3 / (e0.elapsed_time(e1) / 1e3)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
    return res


def padded_f(F):
    return (F + 2 + 7) // 8 * 8


if __name__ == "__main__":
    sys.exit(main())
