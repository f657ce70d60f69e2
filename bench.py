#!/usr/bin/env python
"""bench.py -- stage-1 AFI-GAN training throughput on synthetic R-50-FPN features (BASELINE.json `metric`).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one AFIGAN_Trainer.run_step on one batch of pre-extracted features (reference
afigan/engine/stage1_trainer.py:305-435 minus the guide-model forward): 2 G forwards, 1 G backward, 4 D forwards,
2 D backwards, BCE/L1 losses, 2 SGD updates, and (N > 1) the G and D gradient all-reduces.  Workload = BASELINE.json
configs[1] per GPU: batch 2, 800x1333-image R-50-FPN shapes, five levels, weak scaling.

One JSON line on stdout (rank 0):  value = whole-job img/s with the features already in HBM; e2e = the same step
through the public Stage1Step API fed from pinned HOST buffers (H2D of the features and D2H of the losses inside the
timed region); roofline = the dominant kernel (tcgen05 implicit-GEMM conv) timed per launch with CUDA events in an
extra instrumented step; cpu_baseline = the oracle port (oracle/afigan_oracle.py, torch CPU fp32, all host threads)
on a bounded sample of the same workload.  --impl reference prints the reference arm (same CPU port, more steps).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "afi-gan_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "stage1_afigan_train_throughput"
UNIT = "img/s"
PER_GPU_BATCH = 2
WORKLOAD = ("stage-1 AFI-GAN step (G fwd x2 + bwd, D fwd x4 + bwd x2, BCE/L1, 2 SGD) on synthetic Mask R-CNN R-50-FPN 1x "
            "features of an 800x1333 image and its 0.5x copy, p2-p6, batch 2 per GPU")


# ---- workload definition (config 1/2 of BASELINE.json).  Kept here so that the measured arm never touches oracle/.
C1_LR_SHAPES = ((104, 168), (52, 84), (26, 42), (13, 21), (7, 11))       # p2..p6 of the 0.5x image (400x666 -> padded 416x672)
C1_HR_SHAPES = ((200, 336), (100, 168), (50, 84), (25, 42), (13, 21))    # p2..p6 of the 800x1333 image (padded 800x1344)
G_FWD_FLOP_PER_INPUT_PX = 19_206_144
D_FWD_FLOP_PER_PX = 30_689_280


def synthetic_features(batch, rank, lr_shapes=C1_LR_SHAPES, hr_shapes=C1_HR_SHAPES, channels=256, seed=1234):
    """N(0,1) fp32 features, torch.Generator seeded with 1234 + rank; all HR levels are drawn first, then all LR levels."""
    import torch
    gen = torch.Generator().manual_seed(seed + rank)
    hr = [torch.randn(batch, channels, h, w, generator=gen) for h, w in hr_shapes]
    lr = [torch.randn(batch, channels, h, w, generator=gen) for h, w in lr_shapes]
    return lr, hr


def stage1_step_flops(lr_px, hr_px, g_forwards=2):
    """2 G fwd + 1 G bwd (no input grad) + 4 D fwd + 2 D bwd (no input grad): SURVEY.md §8d.  g_forwards=1: what the step EXECUTES when the
    two bit-identical G(lr) evaluations of the reference (same weights, same input, no state in G) share one forward pass."""
    g, d = G_FWD_FLOP_PER_INPUT_PX, D_FWD_FLOP_PER_PX
    return g_forwards * g * lr_px + (2 * g - 1_179_648) * lr_px + 4 * d * hr_px + 2 * (2 * d - 2_359_296) * hr_px


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], None, set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2]); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "samples": len(sm)}


def cpu_reference_step_time(sample_levels, steps, warmup, batch=PER_GPU_BATCH):
    """Times the CPU port of the reference step (oracle) on the given pyramid levels; returns (seconds per step, cores).
    This function (the cpu_baseline / --impl reference legs) is the ONLY place bench.py touches oracle/."""
    import torch
    from oracle import afigan_oracle as O
    torch.set_num_threads(os.cpu_count())
    g_sd, d_sd = O.init_states(0)
    lr_shapes = [O.C1_LR_SHAPES[i] for i in sample_levels]
    hr_shapes = [O.C1_HR_SHAPES[i] for i in sample_levels]
    lr_f, hr_f = O.synthetic_features(batch, 0, lr_shapes, hr_shapes)
    g_mom, d_mom = {}, {}
    for _ in range(warmup):
        O.stage1_step(g_sd, d_sd, lr_f, hr_f, lr=1e-3, g_mom=g_mom, d_mom=d_mom)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.stage1_step(g_sd, d_sd, lr_f, hr_f, lr=1e-3, g_mom=g_mom, d_mom=d_mom)
    dt = (time.perf_counter() - t0) / steps
    return dt, torch.get_num_threads()


def sample_flop_fraction(sample_levels):
    full = stage1_step_flops(sum(h * w for h, w in C1_LR_SHAPES), sum(h * w for h, w in C1_HR_SHAPES))
    part = stage1_step_flops(sum(C1_LR_SHAPES[i][0] * C1_LR_SHAPES[i][1] for i in sample_levels),
                             sum(C1_HR_SHAPES[i][0] * C1_HR_SHAPES[i][1] for i in sample_levels))
    return part / full


def cpu_baseline(sample_levels=(2, 3, 4), steps=1, warmup=0):
    """Bounded sample: the full reference step restricted to levels p4-p6 (6.2 % of the step's FLOPs), scaled by FLOPs."""
    import torch
    frac = sample_flop_fraction(sample_levels)
    # warm the thread pool / allocator on the two smallest levels first (cheap)
    cpu_reference_step_time((3, 4), 1, 0)
    dt, cores = cpu_reference_step_time(sample_levels, steps, warmup)
    full_step_s = dt / frac
    return {"value": PER_GPU_BATCH / full_step_s, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle/afigan_oracle.py stage1_step (torch {torch.__version__} CPU fp32, {cores} threads) on levels "
                      f"p{sample_levels[0] + 2}-p{sample_levels[-1] + 2} of the same batch-2 workload = {100 * frac:.1f}% of the "
                      f"step's FLOPs, {steps} timed step(s) of {dt:.2f} s, scaled by FLOPs to the full step",
            "sample_seconds": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: p4-p6 if (steps + warm-up) of it fit in ~3 minutes on this host, else p5-p6
    t_small, _ = cpu_reference_step_time((3, 4), 1, 0)
    n_steps, n_warm = max(1, args.steps), max(0, min(args.warmup, 1))
    levels = (2, 3, 4)
    if t_small * sample_flop_fraction(levels) / sample_flop_fraction((3, 4)) * (n_steps + n_warm) > 180.0:
        levels = (3, 4)
    frac = sample_flop_fraction(levels)
    dt, cores = cpu_reference_step_time(levels, n_steps, n_warm)
    import torch
    full_step_s = dt / frac
    val = PER_GPU_BATCH / full_step_s
    sample = (f"reference CPU path = oracle port (torch {torch.__version__} CPU fp32, {cores} threads): the unmodified reference cannot "
              f"travel to the GPU box (/root/reference absent, detectron2 not installed); each step = levels p{levels[0] + 2}-p6 "
              f"({100 * frac:.1f}% of the FLOPs), scaled by FLOPs")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": full_step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "per_gpu_batch": PER_GPU_BATCH},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from afigan import native
    from afigan.engine import Stage1Step
    from afigan.modeling import Discriminator, Generator

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200 (the hot path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    precision = args.precision
    torch.manual_seed(0)                                     # same random-init weights as the reference under seed 0
    G = Generator(n_residual_dense_blocks=3, precision=precision).to(dev)
    D = Discriminator(precision=precision).to(dev)
    if world > 1:
        for p in list(G.parameters()) + list(D.parameters()):
            dist.broadcast(p.data, 0)
    step = Stage1Step(G, D, lr=1e-3, momentum=0.9, weight_decay=1e-4, precision=precision)
    lr_h, hr_h = synthetic_features(PER_GPU_BATCH, rank)     # N(0,1) fp32, seed 1234 + rank
    lr_h, hr_h = [t.pin_memory() for t in lr_h], [t.pin_memory() for t in hr_h]
    lr_d, hr_d = [t.to(dev) for t in lr_h], [t.to(dev) for t in hr_h]
    h2d_bytes = sum(t.numel() * 4 for t in lr_h + hr_h)
    lib = native.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def hbm_step():
        step.run_step(lr_d, hr_d)

    host_losses = torch.empty(4, 8, dtype=torch.float32).pin_memory()

    # end-to-end: every step's features come from pinned HOST memory (one H2D copy of all ten tensors per step, issued on a copy
    # stream one step ahead = a prefetching loader) and the step's losses are read back to the host.
    from afigan.engine import FeaturePrefetcher
    pre = FeaturePrefetcher(dev)
    nl = len(lr_h)
    state = {"slot": pre.submit(lr_h + hr_h)}

    def e2e_step():
        feats = pre.get(state["slot"])
        state["slot"] = pre.submit(lr_h + hr_h)          # next step's H2D copy overlaps this step's compute
        losses = step.run_step(feats[:nl], feats[nl:])
        host_losses.copy_(losses, non_blocking=True)

    for _ in range(max(args.warmup, 3)):
        hbm_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.afi_launch_count(1)
    ms_total = timed(hbm_step, args.steps)
    launches = int(lib.afi_launch_count(0))
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * PER_GPU_BATCH / (ms_step / 1e3)

    e2e_step()
    e2e_ms = timed(e2e_step, args.steps) / args.steps
    torch.cuda.synchronize()
    assert torch.isfinite(host_losses[:2, :5]).all(), "non-finite loss"
    e2e_value = world * PER_GPU_BATCH / (e2e_ms / 1e3)

    # ---- roofline leg: one extra instrumented step, every implicit-GEMM launch bracketed by CUDA events on its stream
    roof = None
    if rank == 0:
        native.check(lib.afi_profile_begin(4096))
    step.overlap = False            # single stream for this step only: per-launch event times must not overlap another stream's kernels
    hbm_step()                      # every rank runs the step (it contains the gradient all-reduces); only rank 0 records events
    step.overlap = True
    barrier()
    if rank == 0:
        n = C.c_int()
        native.check(lib.afi_profile_end(C.byref(n)))
        agg = {}
        kind, fl, ms, cin, cout, px = C.c_int(), C.c_double(), C.c_float(), C.c_int(), C.c_int(), C.c_longlong()
        top = None
        for i in range(n.value):
            lib.afi_profile_get(i, C.byref(kind), C.byref(fl), C.byref(ms), C.byref(cin), C.byref(cout), C.byref(px))
            a = agg.setdefault(kind.value, [0, 0.0, 0.0])
            a[0] += 1; a[1] += fl.value; a[2] += ms.value
            if kind.value in (0, 2, 4, 5) and (top is None or ms.value > top[0]):
                top = (ms.value, fl.value, cin.value, cout.value, px.value)
        pk = peaks()
        names = {0: "k_conv_tc (tcgen05 implicit-GEMM conv/dgrad, per-tap tiles: narrow, 1x1 and non-3x3 layers)", 1: "k_wgrad_tc (tcgen05 weight gradient)",
                 2: "k_conv_simt (fp32 FFMA implicit GEMM)", 3: "k_wgrad_simt (fp32 FFMA weight gradient)",
                 4: "k_conv_halo<PAIR> (tcgen05 cta_group::2 implicit-GEMM conv/dgrad, halo tiles on CTA pairs)",
                 5: "k_conv_halo (tcgen05 implicit-GEMM conv/dgrad, halo tiles on single CTAs)"}
        dom = max(agg, key=lambda k: agg[k][2])
        cnt, flops, tms = agg[dom]
        achieved = flops / (tms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")   # DRAM bytes of the dominant kernel's largest launch (ncu --set full)
        traffic_detail = None
        if os.path.exists(tpath) and dom == 4:
            traffic_detail = json.load(open(tpath))
            traffic = traffic_detail["bytes_per_launch"]
        roof = {"bound": "tensor", "kernel": names[dom], "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_sustained"], "frac_of_burst_peak": achieved / pk["bf16_burst"], "peak_source": pk["source"] +
                " sustained bf16 (kernel timed inside a long step)", "traffic": traffic, "traffic_detail": traffic_detail, "launches_per_step": cnt,
                "flops_per_launch_avg": flops / cnt, "ms_per_launch_avg": tms / cnt, "share_of_step": tms / ms_step,
                "per_kernel": {names[k]: {"launches": v[0], "ms": v[2], "tflops": v[1] / max(v[2], 1e-9) / 1e9} for k, v in agg.items()},
                "largest_launch": None if top is None else {"ms": top[0], "tflops": top[1] / top[0] / 1e9, "cin": top[2], "cout": top[3],
                                                            "pixels": top[4]}}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    lr_px, hr_px = PER_GPU_BATCH * sum(h * w for h, w in C1_LR_SHAPES), PER_GPU_BATCH * sum(h * w for h, w in C1_HR_SHAPES)
    flops_ref = stage1_step_flops(lr_px, hr_px)                                            # the reference's op count
    flops_step = stage1_step_flops(lr_px, hr_px, 1 if step.reuse_g_forward else 2)        # what this step executes
    cpu = cpu_baseline() if (world == 1 and not args.no_cpu_baseline) else None
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if precision.startswith("bf16") else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": PER_GPU_BATCH, "global_batch": world * PER_GPU_BATCH,
                       "parallelism": f"dp{world}", "operand_mode": precision,
                       "l2": "no flush needed: per-step inputs (231 MB fp32) and activations (several GB) exceed the 126 MB L2"},
            "step_tflops_per_gpu": flops_step / (ms_step * 1e-3) / 1e12,
            "step_flops": {"executed": flops_step, "reference_op_count": flops_ref,
                           "note": "the reference evaluates G(lr) twice per step with identical weights and inputs (detached for the D phase, with a graph "
                                   "for the G phase); this step evaluates it once and uses the bit-identical result in both phases"
                                   if step.reuse_g_forward else "literal: two G(lr) evaluations per step"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4 * 8 * 4},
            "gpu_launches": launches, "roofline": roof}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("AFIGAN_PRECISION", "bf16"), choices=["bf16", "fp32", "bf16_simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29517", os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps),
               "--warmup", str(args.warmup), "--precision", args.precision]
        sys.exit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
