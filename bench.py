"""Benchmark of the SC-SfMLearner training hot path (BASELINE.json metric: train-step frames/sec at
256x832 ResNet-18 on 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun launches N>1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)

A "step" is one complete optimisation step (train.py:254-282): 3 DispResNet-18 + 4 PoseResNet-18
forward/backward calls, fused photometric/geometry/smoothness losses, gradient all-reduce (N>1), Adam,
on a synthetic KITTI-shaped batch of 4 frames per GPU (BASELINE config 2; weak scaling).
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "sc-sfmlearner-release_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

H, W, PER_GPU_BATCH, N_REF = 256, 832, 4, 2
# algorithmic work per frame (SURVEY.md section 8d / BASELINE.md section 4)
CONV_TRAIN_GFLOP_PER_FRAME = 453.4
WORKLOAD = "DispResNet18+PoseResNet18 full train step, batch 4 per GPU, 256x832 synthetic KITTI triplets (config 2)"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"], "tflops_sustained": d["bf16_tflops_sustained"],
                "source": "MEASURED_PEAKS.json (of measured)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "B200_PROFILING.md fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_batch(rank, pinned):
    from scsfm import synth
    tgt, refs, K = synth.triplet(1234 + rank, PER_GPU_BATCH, H, W, N_REF)
    if pinned:
        tgt, refs, K = tgt.pin_memory(), [r.pin_memory() for r in refs], K.pin_memory()
    return tgt, refs, K


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import models
    from scsfm import lib as L
    from scsfm import nnops
    from scsfm.trainer import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.load()          # fails loudly if libscsfm.so is missing
    nnops.CONFIG["conv_mode"] = args.conv_mode
    torch.manual_seed(0)
    disp, pose = models.DispResNet(18, False).to(dev).train(), models.PoseResNet(18, False).to(dev).train()
    trainer = Trainer(disp, pose, lr=1e-4, num_scales=1, with_ssim=1, with_mask=1, with_auto_mask=1, padding_mode="zeros",
                      w1=1.0, w2=0.1, w3=0.5, distributed=world > 1)
    h_tgt, h_refs, h_K = synthetic_batch(rank, pinned=True)
    d_tgt, d_refs, d_K = h_tgt.to(dev), [r.to(dev) for r in h_refs], h_K.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """device time of `steps` calls (L2 flushed before each), max over ranks, in ms"""
        evs = []
        barrier()
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        total = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.MAX)
        return float(total)

    # ---- warm-up with full per-family profiling: finds the dominant kernel family -----------------
    trainer.step(d_tgt, d_refs, d_K)          # first step unprofiled: lazy kernel loading, allocations, Adam state
    torch.cuda.synchronize()
    L.PROF.update(enabled=True, only=None, events=[])
    for _ in range(max(args.warmup, 3)):
        trainer.step(d_tgt, d_refs, d_K)
    torch.cuda.synchronize()
    fam = {}
    for family, work, e0, e1, _ in L.PROF["events"]:
        t = fam.setdefault(family, [0.0, 0.0, 0])
        t[0] += e0.elapsed_time(e1); t[1] += work; t[2] += 1
    dominant = max(fam, key=lambda k: fam[k][0])
    breakdown = {k: round(v[0] / max(args.warmup, 3), 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])}

    # ---- per-kernel roofline pass: the same step, eager, CUDA events only around the dominant family ------
    L.PROF.update(enabled=True, only={dominant}, events=[])
    ms_eager = timed(lambda: trainer.step(d_tgt, d_refs, d_K), args.steps)
    dom_ms = sum(e[2].elapsed_time(e[3]) for e in L.PROF["events"])
    dom_work = sum(e[1] for e in L.PROF["events"])
    dom_n = len(L.PROF["events"])
    L.PROF.update(enabled=False, only=None, events=[])

    # ---- timed region: device-resident inputs.  Single GPU: the whole step is one CUDA-graph replay ---------
    graphed = world == 1 and not args.no_graph
    if graphed:
        trainer.capture(d_tgt, d_refs, d_K)
    elif world > 1 and args.graph_ddp:
        # opt-in experiment: the data-parallel step (incl. the NCCL all-reduce on the side stream) as one CUDA graph
        try:
            trainer.capture(d_tgt, d_refs, d_K, allow_distributed=True)
            graphed = True
        except Exception as e:      # noqa: BLE001 -- fall back to the validated eager path
            print("graph capture of the data-parallel step failed, running eagerly: %r" % (e,), file=sys.stderr)
    launches0 = L.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(lambda: trainer.step(d_tgt, d_refs, d_K), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = trainer.launches_per_step * args.steps if graphed else (L.launch_count() - launches0)

    # ---- end to end: pinned host inputs copied in, loss read back, every step ----------------------
    result = torch.empty(4, dtype=torch.float32).pin_memory()

    def e2e_step():
        t = h_tgt.to(dev, non_blocking=True)
        r = [x.to(dev, non_blocking=True) for x in h_refs]
        k = h_K.to(dev, non_blocking=True)
        out = trainer.step(t, r, k)
        result.copy_(torch.stack(out), non_blocking=True)
        torch.cuda.current_stream().synchronize()       # the user reads the loss (train.py:277)

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    h2d = sum(t.numel() * 4 for t in [h_tgt, h_K] + h_refs)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    frames = PER_GPU_BATCH * world * args.steps
    is_conv = dominant.startswith("conv")
    if is_conv:
        achieved = dom_work / (dom_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": round(achieved, 2), "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                "frac": round(achieved / peaks["tflops_sustained"], 5), "traffic": None}
    else:
        achieved = dom_work / (dom_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": round(achieved / peaks["hbm_gbs"], 5), "traffic": None}
    try:        # measured DRAM traffic per launch of the family's kernels (ncu capture summarised by tools/ncu_summarise.py)
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_by_family.json")))
        if dominant in tr:
            roof["traffic"] = tr[dominant]["dram_bytes_per_launch"]
            roof["traffic_source"] = tr["_source"]
            roof["algorithmic_bytes_or_flops_per_launch"] = round(dom_work / max(dom_n, 1))
    except (OSError, ValueError, KeyError):
        pass
    roof.update(kernel=dominant, launches_timed=dom_n, avg_launch_us=round(1e3 * dom_ms / max(dom_n, 1), 2),
                share_of_step=round(dom_ms / ms_eager, 4), peak_source=peaks["source"],
                note="achieved = algorithmic FLOPs (2*M*N*K per conv pass) or bytes of the family / its CUDA-event time over the "
                     "same K steps run eagerly (events cannot be recorded inside the replayed CUDA graph); sustained bf16 "
                     "peak is the denominator because the kernel runs inside a long step")
    line = {
        "metric": "train-step frames/sec at 256x832 ResNet18 (DispResNet18+PoseResNet18, fwd+bwd+losses+Adam)",
        "value": round(frames / (ms * 1e-3), 3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32" if args.conv_mode == "fp32" else "tf32",
        "data": "synthetic", "impl": "ours",
        "config": {"workload": WORKLOAD, "global_batch": PER_GPU_BATCH * world, "height": H, "width": W, "n_ref": N_REF,
                   "parallelism": "dp%d" % world, "conv_mode": args.conv_mode, "l2": "flushed (256 MiB write) before every step",
                   "cuda_graph": graphed, "eager_ms_per_step": round(ms_eager / args.steps, 3),
                   "loss_flags": "num_scales=1 ssim=1 mask=1 auto_mask=1 zeros"},
        "e2e": {"value": round(frames / (ms_e2e * 1e-3), 3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 16, "ms_per_step": round(ms_e2e / args.steps, 3)},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
        "step_breakdown_ms": breakdown,
        "conv_gflop_per_frame_train": CONV_TRAIN_GFLOP_PER_FRAME,
        "step_tflops": round(CONV_TRAIN_GFLOP_PER_FRAME * frames / (ms * 1e-3) / 1e3, 2),
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference(steps=2, warmup=1, budget_s=40.0)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the step (oracle port), all host threads
# --------------------------------------------------------------------------------------------------
def usable_cores():
    """Host cores this process may actually use: affinity mask and cgroup CPU quota, not just os.cpu_count()
    (the GPU box reports 128 CPUs but oversubscribing a quota-limited container makes torch crawl)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except (OSError, ValueError):
        pass
    for q, per in (("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"),):
        try:
            quota, period = int(open(q).read()), int(open(per).read())
            if quota > 0:
                n = min(n, max(1, quota // period))
        except (OSError, ValueError):
            pass
    return max(1, n)


def pick_threads():
    """All usable host cores, unless a smaller torch thread count is measurably faster on this box (a 128-CPU
    host running a 3x3 conv + its backward: oversubscribed intra-op pools can be orders of magnitude slower)."""
    limit = usable_cores()
    cands = sorted({c for c in (8, 16, 32, 64, 128) if c <= limit} | {limit})
    x = torch.randn(2, 64, 64, 208, requires_grad=True)
    conv = torch.nn.Conv2d(64, 64, 3, padding=1)
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        conv(x).sum().backward()
        t0 = time.perf_counter()
        for _ in range(3):
            conv(x).sum().backward()
        dt = time.perf_counter() - t0
        if dt < best_t * 0.9:
            best, best_t = c, dt
    return best, limit


def cpu_reference(steps, warmup, budget_s):
    from oracle import geometry as OGEO
    from oracle import nets as N
    from oracle import step as OS
    from scsfm import synth
    OGEO.USE_LIBRARY_KERNELS = True        # F.grid_sample / F.avg_pool2d, exactly what the reference calls on CPU
    cores, usable = pick_threads()
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    disp, pose = N.DispResNet(18).train(), N.PoseResNet(18).train()
    opt = OS.make_optimizer(disp, pose, lr=1e-4)
    batch = PER_GPU_BATCH
    tgt, refs, K = synth.triplet(1234, batch, H, W, N_REF)

    def one(b):
        t0 = time.perf_counter()
        OS.train_step(disp, pose, opt, tgt[:b], [r[:b] for r in refs], K[:b], num_scales=1, with_ssim=1, with_mask=1,
                      with_auto_mask=1, padding_mode="zeros")
        return time.perf_counter() - t0
    # bounded sample: probe with one frame, then pick the largest per-step batch and step count that fit the budget
    probe = one(1)
    while batch > 1 and probe * batch * (steps + max(warmup - 1, 0)) > budget_s:
        batch //= 2
    steps = max(1, min(steps, int(budget_s / max(probe * batch, 1e-6)) - max(warmup - 1, 0)))
    if probe * batch * (steps + 1) > budget_s:
        warmup = 1                      # the probe step is the only warm-up that fits
    for _ in range(max(warmup - 1, 0)):
        one(batch)
    times = [one(batch) for _ in range(steps)]
    total = sum(times)
    return {"value": round(batch * steps / total, 4), "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": "%d timed step(s) of the oracle port of train.py:259-282 at batch %d (DispResNet18+PoseResNet18, "
                      "256x832, 2 refs, fp32, torch CPU with %d threads; %d usable host cores, thread count auto-picked)"
                      % (steps, batch, cores, usable),
            "ms_per_step": round(1e3 * total / steps, 1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_reference(steps=args.steps, warmup=args.warmup, budget_s=150.0)
    line = {
        "metric": "train-step frames/sec at 256x832 ResNet18 (DispResNet18+PoseResNet18, fwd+bwd+losses+Adam)",
        "value": base["value"], "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic", "impl": "reference",
        "config": {"workload": WORKLOAD, "global_batch": PER_GPU_BATCH, "height": H, "width": W, "n_ref": N_REF,
                   "parallelism": "cpu"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--conv-mode", choices=["fp32", "tf32"], default="tf32",
                    help="tf32 = tcgen05 tensor-core convolutions (the reference's own cuDNN default arithmetic on a GPU); fp32 = exact CUDA-core parity mode")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph-ddp", action="store_true", help="N>1: try to capture the data-parallel step in a CUDA graph (experimental)")
    ap.add_argument("--no-graph", action="store_true", help="time the eager step instead of the CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
