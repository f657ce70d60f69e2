#!/usr/bin/env python
"""bench.py -- stage-1 AFI-GAN training throughput on synthetic R-50-FPN features (BASELINE.json `metric`).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|split|fp32]
                  [--workload stage1|g_only|stage2_c3|pafpn_c4|infer_c5]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one AFIGAN_Trainer.run_step on one batch of pre-extracted features (reference
afigan/engine/stage1_trainer.py:305-435 minus the guide-model forward): 2 G forwards, 1 G backward, 4 D forwards,
2 D backwards, BCE/L1 losses, 2 SGD updates, and (N > 1) the G and D gradient all-reduces.  Workload = BASELINE.json
configs[1] per GPU: batch 2, 800x1333-image R-50-FPN shapes, five levels, weak scaling.

One JSON line on stdout (rank 0):  value = whole-job img/s with the features already in HBM; e2e = the same step
through the public Stage1Step API fed from pinned HOST buffers (H2D of the features and D2H of the losses inside the
timed region); roofline = the dominant kernel (tcgen05 implicit-GEMM conv) timed per launch with CUDA events in an
extra instrumented step; parity = the timed operand mode's measured errors against the full-size golden fixture made
from the unmodified reference (tests/golden/stage1_full.npz); parity_mode = throughput + errors of the split-precision
mode (the one that meets north_star's tolerance) in the same run; extra = G and D fwd+bwd TFLOP/s (the metric's second
half); cpu_baseline = ONE whole five-level step of the reference's own modules (oracle/_ref through oracle/ref_runner.py,
torch CPU fp32, all host threads; the oracle port if oracle/_ref was not staged).  --impl reference prints the reference
arm: whole steps of the same CPU path, as many as fit a ~150 s budget.  --workload selects the other BASELINE.json
configs (one JSON line each; the default line is config 1/2).
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


# ---------------------------------------------------------------------------------------------------------------------
# CPU reference legs.  These functions (cpu_baseline / --impl reference) are the ONLY place bench.py touches oracle/.
# ---------------------------------------------------------------------------------------------------------------------
class _CpuStep:
    """One trainer process of the reference's CPU path: oracle/_ref (the reference's own Generator / Discriminator module files driven by
    the restated run_step of oracle/ref_runner.py) when it was staged, else the oracle port (oracle/afigan_oracle.py)."""

    def __init__(self):
        import torch
        torch.set_num_threads(os.cpu_count())
        self.cores = torch.get_num_threads()
        from oracle import ref_runner
        if ref_runner.available():
            self.kind = "reference"
            self.runner = ref_runner.ReferenceStage1(lr=1e-3, momentum=0.9, weight_decay=1e-4)
            self.what = (f"the reference's own generator_rdb.py / feature_patch_discriminator.py (oracle/_ref, unmodified) under the run_step of "
                         f"stage1_trainer.py:334-433 restated in oracle/ref_runner.py, torch {torch.__version__} CPU fp32, {self.cores} threads")
        else:
            from oracle import afigan_oracle as O
            self.kind = "port"
            self.O = O
            self.g_sd, self.d_sd = O.init_states(0)
            self.g_mom, self.d_mom = {}, {}
            self.what = f"oracle port oracle/afigan_oracle.py stage1_step (oracle/_ref not staged), torch {torch.__version__} CPU fp32, {self.cores} threads"

    def step(self, lr_f, hr_f):
        if self.kind == "reference":
            return self.runner.run_step(lr_f, hr_f)
        return self.O.stage1_step(self.g_sd, self.d_sd, lr_f, hr_f, lr=1e-3, g_mom=self.g_mom, d_mom=self.d_mom)


def _cpu_time_steps(cpu, n_steps, levels=None):
    lr_shapes = C1_LR_SHAPES if levels is None else [C1_LR_SHAPES[i] for i in levels]
    hr_shapes = C1_HR_SHAPES if levels is None else [C1_HR_SHAPES[i] for i in levels]
    lr_f, hr_f = synthetic_features(PER_GPU_BATCH, 0, lr_shapes, hr_shapes)
    times = []
    for _ in range(n_steps):
        t0 = time.perf_counter()
        cpu.step(lr_f, hr_f)
        times.append(time.perf_counter() - t0)
    return times


def cpu_baseline():
    """Bounded sample of the default line: ONE whole five-level config-1 step (batch 2) of the CPU path, timed after a warm-up step on the
    two smallest levels (thread pool, allocator, oneDNN primitives).  Nothing is scaled or extrapolated."""
    cpu = _CpuStep()
    _cpu_time_steps(cpu, 1, levels=(3, 4))
    t = _cpu_time_steps(cpu, 1)[0]
    return {"value": PER_GPU_BATCH / t, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind,
            "sample": f"1 whole five-level step (batch 2, all of p2-p6) of {cpu.what}; {t:.1f} s, after one warm-up step on p5-p6",
            "sample_seconds": t, "steps_timed": 1}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu = _CpuStep()
    budget = float(os.environ.get("AFIGAN_REFERENCE_BUDGET_S", "150"))
    warm = _cpu_time_steps(cpu, 1)                      # ONE whole-step warm-up, which also sizes the timed region
    n_steps = int(max(1, min(args.steps, (budget - warm[0]) // max(warm[0], 1e-3))))
    times = _cpu_time_steps(cpu, n_steps)
    dt = sum(times) / len(times)
    val = PER_GPU_BATCH / dt
    sample = (f"{n_steps} whole five-level step(s) (batch 2, all of p2-p6; --steps {args.steps} capped by a {budget:.0f} s budget) after 1 "
              f"whole-step warm-up, of {cpu.what}")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": n_steps, "steps_requested": args.steps,
            "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "per_gpu_batch": PER_GPU_BATCH, "global_batch": PER_GPU_BATCH,
                                                            "note": "the CPU path runs on rank 0's host cores only; it does not scale with --gpus"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind, "sample": sample, "steps_timed": n_steps,
                             "step_seconds": [round(t, 3) for t in times]},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def _dist_setup():
    import torch
    import torch.distributed as dist
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
    return rank, world, local_rank, dev


def _timing_tools(world, dev):
    import torch
    import torch.distributed as dist

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

    return barrier, timed


def _dist_teardown(world):
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


def _models(precision, dev):
    import torch
    from afigan.modeling import Discriminator, Generator
    torch.manual_seed(0)                                     # same random-init weights as the reference under seed 0 (G first, then D)
    G = Generator(n_residual_dense_blocks=3, precision=precision).to(dev)
    D = Discriminator(precision=precision).to(dev)
    D.Discriminators[0].train()
    return G, D


NORTH_STAR_TOL = {"features_and_gradients_rel": 1e-3, "loss_abs": 1e-4}


def golden_parity(step, G, D, lr_d, hr_d):
    """Errors of ONE step (fresh seed-0 weights, no optimiser update) against tests/golden/stage1_full.npz: the unmodified reference modules
    run at this very size on these very inputs (tests/golden/make_golden_full.py).  Not oracle/: a committed fixture."""
    import numpy as np
    import torch
    path = os.path.join(ROOT, "tests", "golden", "stage1_full.npz")
    if not os.path.exists(path):
        return {"unavailable": "tests/golden/stage1_full.npz missing"}
    fx = np.load(path)
    assert tuple(map(tuple, fx["hr_shapes"])) == C1_HR_SHAPES and int(fx["seed"]) == 1234 and int(fx["batch"]) == PER_GPU_BATCH
    step.run_step(lr_d, hr_d, apply_updates=False)
    m = step.metrics(5)
    d_loss = np.array([m[f"d_loss_p{l}"] for l in range(2, 7)])
    g_loss = np.array([m[f"g_loss_p{l}"] for l in range(2, 7)])

    def sample(t, n=257):
        f = t.detach().reshape(-1)
        idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long().to(f.device)
        return f[idx].double().cpu().numpy()

    out = {"d_loss_rel_worst": float(np.max(np.abs(d_loss - fx["d_loss"]) / fx["d_loss"])), "d_loss_abs_worst": float(np.max(np.abs(d_loss - fx["d_loss"]))),
           "g_loss_abs_worst": float(np.max(np.abs(g_loss - fx["g_loss"])))}
    for tag, mod, pre in (("d", D.Discriminators[0], "dgrad"), ("g", G.Generators[0], "ggrad")):
        wn = ws = 0.0
        for name, p in mod.named_parameters():
            if tag == "d" and name.endswith("0.bias") and not name.startswith("3."):
                continue                                     # true gradient 0 (bias in front of a batch-statistics BatchNorm)
            rn = float(fx[f"{pre}_norm/" + name])
            wn = max(wn, abs(float(p.grad.norm()) - rn) / rn)
            ref = fx[f"{pre}_sample/" + name].astype(np.float64)
            ws = max(ws, float(np.linalg.norm(sample(p.grad) - ref) / np.linalg.norm(ref)))
        out[f"{tag}_grad_norm_rel_worst"], out[f"{tag}_grad_sample_rel_worst"] = wn, ws
    out["meets_north_star"] = bool(max(out["d_grad_sample_rel_worst"], out["g_grad_sample_rel_worst"]) <= NORTH_STAR_TOL["features_and_gradients_rel"]
                                   and max(out["d_loss_abs_worst"], out["g_loss_abs_worst"]) <= NORTH_STAR_TOL["loss_abs"])
    out["against"] = ("tests/golden/stage1_full.npz = the UNMODIFIED reference modules (torch CPU fp32) on the same config-1 batch (seed 1234) "
                      "and seed-0 weights; worst over the 5 levels / the 33 parameter tensors (norm, and a 257-point strided sample each)")
    out["north_star_tolerance"] = NORTH_STAR_TOL
    return out


def phase_extras(step, lr_d, hr_d, pk, n=2):
    """G and D fwd+bwd throughput (BASELINE.json metric, second half) from CUDA events around the phases of `n` instrumented steps."""
    t = {}
    for _ in range(n):
        step.run_step(lr_d, hr_d, timers=t)
    lr_px, hr_px = PER_GPU_BATCH * sum(h * w for h, w in C1_LR_SHAPES), PER_GPU_BATCH * sum(h * w for h, w in C1_HR_SHAPES)
    g_f, d_f = G_FWD_FLOP_PER_INPUT_PX, D_FWD_FLOP_PER_PX
    g_flops = (g_f + 2 * g_f - 1_179_648) * lr_px                    # one forward + one backward (no input gradient)
    d_flops = (2 * d_f + 2 * (2 * d_f - 2_359_296)) * hr_px          # the D phase: two forwards + two backwards (no input gradient)
    g_ms = (t["g_forward"] + t["l1_g_backward"]) / n
    d_ms = t["d_forward_backward"] / n
    g_tf, d_tf = g_flops / g_ms / 1e9, d_flops / d_ms / 1e9
    return {"g_fwd_bwd_tflops": g_tf, "g_frac_of_sustained_peak": g_tf / pk["bf16_sustained"], "g_fwd_bwd_ms": g_ms,
            "d_fwd_bwd_tflops": d_tf, "d_frac_of_sustained_peak": d_tf / pk["bf16_sustained"], "d_fwd_bwd_ms": d_ms,
            "phase_ms": {k: v / n for k, v in t.items()},
            "note": "AF interpolator: 1 forward (5 levels, grouped) + L1 + 1 backward; discriminator: the D phase = 2 forwards + 2 backwards over 10 "
                    "grouped calls incl. BatchNorm, head, BCE; algorithmic FLOPs of SURVEY.md §8d over CUDA-event time, every elementwise pass included"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from afigan import native
    from afigan.engine import FeaturePrefetcher, Stage1Step

    rank, world, local_rank, dev = _dist_setup()
    barrier, timed = _timing_tools(world, dev)
    precision = args.precision
    G, D = _models(precision, dev)
    step = Stage1Step(G, D, lr=1e-3, momentum=0.9, weight_decay=1e-4, precision=precision)   # broadcasts rank 0's parameters (DDP semantics)
    lr_h, hr_h = synthetic_features(PER_GPU_BATCH, rank)     # N(0,1) fp32, seed 1234 + rank
    lr_h, hr_h = [t.pin_memory() for t in lr_h], [t.pin_memory() for t in hr_h]
    lr_d, hr_d = [t.to(dev) for t in lr_h], [t.to(dev) for t in hr_h]
    h2d_bytes = sum(t.numel() * 4 for t in lr_h + hr_h)
    lib = native.lib()
    pk = peaks()
    extras = world == 1 and not args.no_extras

    parity = golden_parity(step, G, D, lr_d, hr_d) if extras else None      # before any update: seed-0 weights
    if parity is not None:
        parity["operand_mode"] = precision

    def hbm_step():
        step.run_step(lr_d, hr_d)

    host_losses = torch.empty(4, 8, dtype=torch.float32).pin_memory()

    # end-to-end: every step's features come from pinned HOST memory (one H2D copy of all ten tensors per step, issued on a copy
    # stream one step ahead = a prefetching loader) and the step's losses are read back to the host.
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

    # ---- roofline leg: three extra instrumented steps, every implicit-GEMM launch bracketed by CUDA events on its stream; per launch the
    # MEDIAN of the three (one step alone can catch a power-cap clock excursion in its largest launch)
    roof = None
    step.overlap = False            # single stream for these steps only: per-launch event times must not overlap another stream's kernels
    runs = []
    for _ in range(3):
        if rank == 0:
            native.check(lib.afi_profile_begin(4096))
        hbm_step()                  # every rank runs the step (it contains the gradient all-reduces); only rank 0 records events
        barrier()
        if rank == 0:
            n = C.c_int()
            native.check(lib.afi_profile_end(C.byref(n)))
            kind, fl, ms, cin, cout, px = C.c_int(), C.c_double(), C.c_float(), C.c_int(), C.c_int(), C.c_longlong()
            rec = []
            for i in range(n.value):
                lib.afi_profile_get(i, C.byref(kind), C.byref(fl), C.byref(ms), C.byref(cin), C.byref(cout), C.byref(px))
                rec.append((kind.value, fl.value, ms.value, cin.value, cout.value, px.value))
            runs.append(rec)
    step.overlap = True
    if rank == 0:
        assert len({len(r) for r in runs}) == 1 and all(a[:2] == b[:2] for a, b in zip(runs[0], runs[1])), "instrumented steps differ"
        agg = {}
        top = None
        for i, (kind_v, fl_v, _, cin_v, cout_v, px_v) in enumerate(runs[0]):
            ms_v = sorted(r[i][2] for r in runs)[1]
            a = agg.setdefault(kind_v, [0, 0.0, 0.0])
            a[0] += 1; a[1] += fl_v; a[2] += ms_v
            if kind_v in (0, 2, 4, 5) and (top is None or ms_v > top[0]):
                top = (ms_v, fl_v, cin_v, cout_v, px_v)
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
        if os.path.exists(tpath) and dom == 4 and precision in ("bf16", "split"):
            traffic_detail = json.load(open(tpath))
            if precision == "split":
                traffic_detail = traffic_detail.get("split")
            traffic = traffic_detail["bytes_per_launch"] if traffic_detail else None
        roof = {"bound": "tensor", "kernel": names[dom], "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_sustained"], "frac_of_burst_peak": achieved / pk["bf16_burst"], "peak_source": pk["source"] +
                " sustained bf16 (kernel timed inside a long step)", "traffic": traffic, "traffic_detail": traffic_detail, "launches_per_step": cnt,
                "flops_per_launch_avg": flops / cnt, "ms_per_launch_avg": tms / cnt, "share_of_step": tms / ms_step,
                "per_kernel": {names[k]: {"launches": v[0], "ms": v[2], "tflops": v[1] / max(v[2], 1e-9) / 1e9} for k, v in agg.items()},
                "largest_launch": None if top is None else {"ms": top[0], "tflops": top[1] / top[0] / 1e9, "cin": top[2], "cout": top[3],
                                                            "pixels": top[4]}}
        if roof["frac"] > 1.0:
            roof["note"] = ("frac > 1: the kernel's average over the step exceeds the driver-measured SUSTAINED cuBLAS bf16 figure used as `peak` "
                            "(MEASURED_PEAKS.json); against the burst figure it is frac_of_burst_peak")
        if precision == "split":
            roof["note"] = ("split-precision mode: achieved counts ALGORITHMIC FLOPs; the tensor cores execute 3 (backward, forward-only) to 6 "
                            "(training forward) bf16 MMA passes per algorithmic product, so the MMA pipe runs at 3-6x this figure")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    extra = phase_extras(step, lr_d, hr_d, pk) if extras else None

    # ---- the split-precision (parity) mode in the same run: its errors against the same golden and its throughput
    parity_mode = None
    if extras and precision != "split":

        G2, D2 = _models("split", dev)
        step2 = Stage1Step(G2, D2, lr=1e-3, momentum=0.9, weight_decay=1e-4, precision="split")
        par2 = golden_parity(step2, G2, D2, lr_d, hr_d)
        for _ in range(2):
            step2.run_step(lr_d, hr_d)
        ns = max(2, min(args.steps, 5))
        ms2 = timed(lambda: step2.run_step(lr_d, hr_d), ns) / ns
        parity_mode = {"operand_mode": "split", "value": PER_GPU_BATCH / (ms2 / 1e3), "unit": UNIT, "ms_per_step": ms2, "steps": ns, "warmup": 2,
                       "step_tflops_per_gpu": stage1_step_flops(PER_GPU_BATCH * sum(h * w for h, w in C1_LR_SHAPES),
                                                                PER_GPU_BATCH * sum(h * w for h, w in C1_HR_SHAPES), 1) / (ms2 * 1e-3) / 1e12,
                       "what": "AFI_PREC_SPLIT: fp32 storage; every GEMM on the tcgen05 tensor cores as bf16x3 plane-pair products with fp32 TMEM "
                               "accumulation (6 pairs for the sign-critical training forwards, 3 for backward / forward-only GEMMs): the "
                               "library's DEFAULT mode, the one whose results meet north_star's tolerance",
                       "parity": par2}
        del step2, G2, D2
        torch.cuda.empty_cache()

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
    if parity is not None:
        line["parity"] = parity
    if parity_mode is not None:
        line["parity_mode"] = parity_mode
    if extra is not None:
        line["extra"] = extra
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
# The other BASELINE.json configs (one JSON line each): --workload g_only | stage2_c3 | pafpn_c4 | infer_c5
# ---------------------------------------------------------------------------------------------------------------------
def _line(metric, value, unit, world, args, ms_step, precision, workload, **more):
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if precision.startswith("bf16") else "f32", "data": "synthetic",
            "config": {"workload": workload, "operand_mode": precision, "parallelism": f"dp{world}"}}
    line.update(more)
    return line


def run_g_only(args):
    """AF interpolator alone at config-1 shapes: ONE grouped forward over the five LR levels (batch 2) + ONE backward (weight gradients, no
    input gradient -- stage 1 detaches the inputs): the metric's 'AF-interp fwd+bwd TFLOP/s vs peak'."""
    import torch
    from afigan import native
    from afigan.engine import Stage1Step
    rank, world, local_rank, dev = _dist_setup()
    barrier, timed = _timing_tools(world, dev)
    G, D = _models(args.precision, dev)
    step = Stage1Step(G, D, precision=args.precision, distributed=False)
    lr_h, hr_h = synthetic_features(PER_GPU_BATCH, rank)
    lr_d, hr_d = [t.to(dev) for t in lr_h], [t.to(dev) for t in hr_h]
    trs = step._g_forward(lr_d, hr_d, True, "g")
    dys = [torch.randn_like(t) / t.numel() for t in trs]
    lib = native.lib()

    def one():
        native.check(lib.afi_zero(step.g_acc.data_ptr(), step.g_acc.numel(), native.stream_ptr()))
        step._g_forward(lr_d, hr_d, True, "g")
        step._g_backward(lr_d, hr_d, dys, "g")

    for _ in range(max(args.warmup, 3)):
        one()
    lib.afi_launch_count(1)
    ms = timed(one, args.steps) / args.steps
    launches = int(lib.afi_launch_count(0))
    lr_px = PER_GPU_BATCH * sum(h * w for h, w in C1_LR_SHAPES)
    flops = (3 * G_FWD_FLOP_PER_INPUT_PX - 1_179_648) * lr_px
    tf = flops / ms / 1e9
    pk = peaks()
    if rank == 0:
        print(json.dumps(_line("afi_generator_fwd_bwd_tflops", world * tf, "TFLOP/s", world, args, ms, args.precision,
                               "AF interpolator fwd + bwd (no input gradient), five LR levels of config 1 grouped, batch 2 per GPU",
                               img_per_s=world * PER_GPU_BATCH / (ms * 1e-3), gpu_launches=launches,
                               roofline={"bound": "tensor", "achieved": tf, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": tf / pk["bf16_sustained"],
                                         "frac_of_burst_peak": tf / pk["bf16_burst"], "traffic": None,
                                         "note": "whole fwd+bwd incl. layout conversions and elementwise passes; algorithmic FLOPs (SURVEY.md §8d)"})))
    _dist_teardown(world)


def run_pafpn_c4(args):
    """BASELINE config 4: the PAFPN / FPN top-down path with the shared AF interpolator, three merges (25x42 -> 50x84 -> 100x168 with
    1024 / 512 / 256-channel laterals), batch 16 per GPU, forward + FULL backward (interpolator, lateral and input gradients) through the
    autograd modules (`Generator.merge`, the call the necks make)."""
    import torch
    from afigan import native
    rank, world, local_rank, dev = _dist_setup()
    barrier, timed = _timing_tools(world, dev)
    N = 16
    G, _ = _models(args.precision, dev)
    G.deferred_weight_grads = True          # one packed gradient accumulator for the three calls of a backward pass
    gen = torch.Generator().manual_seed(33 + rank)
    c5 = torch.randn(N, 256, 13, 21, generator=gen).to(dev).requires_grad_(True)
    lat_in = [torch.randn(N, c, h, w, generator=gen).to(dev).requires_grad_(True) for c, h, w in ((1024, 25, 42), (512, 50, 84), (256, 100, 168))]
    lat_w = [(torch.randn(256, c, 1, 1, generator=gen) * (1.0 / c) ** 0.5).to(dev).requires_grad_(True) for c in (1024, 512, 256)]
    lat_b = [torch.zeros(256, device=dev, requires_grad=True) for _ in range(3)]
    dy = torch.randn(N, 256, 100, 168, generator=gen).to(dev) / (N * 256 * 100 * 168)
    lib = native.lib()

    upcast = False

    def one():
        cur = c5.float() if upcast else c5
        for x_, w_, b_ in zip(lat_in, lat_w, lat_b):
            cur = G.merge(cur, x_.float() if upcast else x_, w_, b_, "sum")
        cur.backward(dy)
        for t in [c5] + lat_in + lat_w + lat_b + list(G.parameters()):
            t.grad = None

    for _ in range(max(args.warmup, 3)):
        one()
    lib.afi_launch_count(1)
    ms = timed(one, args.steps) / args.steps
    launches = int(lib.afi_launch_count(0))
    # the same step fed with bf16 channels_last bottom-up maps (what an autocast backbone hands the neck): the boundary takes them as they
    # are (afi_view4.dtype, any strides) -- no up-cast copy, no transpose
    fp32_inputs = (c5, lat_in)
    c5 = c5.detach().bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    lat_in = [t.detach().bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True) for t in lat_in]
    for _ in range(3):
        one()
    ms_bf16_in = timed(one, args.steps) / args.steps
    upcast = True                           # ... against the same bf16 tensors up-cast by torch in front of every call (the ABI <= 3 path)
    for _ in range(3):
        one()
    ms_bf16_upcast = timed(one, args.steps) / args.steps
    upcast = False
    c5, lat_in = fp32_inputs
    afi_px = N * (13 * 21 + 25 * 42 + 50 * 84)
    flops = 3 * G_FWD_FLOP_PER_INPUT_PX * afi_px + 3 * 2 * 256 * sum(c * h * w for c, h, w in ((1024, 25, 42), (512, 50, 84), (256, 100, 168))) * N
    tf = flops / ms / 1e9
    pk = peaks()
    if rank == 0:
        print(json.dumps(_line("pafpn_topdown_afi_fwd_bwd_img_per_s", world * N / (ms * 1e-3), "img/s", world, args, ms, args.precision,
                               "config 4: PAFPN top-down path, 3 AF-interpolator merges + 1x1 laterals, batch 16 per GPU, fwd + full bwd via autograd",
                               per_gpu_batch=N, step_tflops_per_gpu=tf, gpu_launches=launches,
                               bf16_channels_last_inputs={"ms_per_step": ms_bf16_in, "img_per_s": world * N / (ms_bf16_in * 1e-3),
                                                          "ms_per_step_with_torch_upcast": ms_bf16_upcast,
                                                          "note": "bf16 leaves also receive bf16 gradients (autograd casts the fp32 input gradients)"},
                               roofline={"bound": "tensor", "achieved": tf, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": tf / pk["bf16_sustained"],
                                         "traffic": None, "note": "AFI fwd + 2x bwd + laterals fwd/dgrad/wgrad, algorithmic FLOPs over the whole autograd step"})))
    _dist_teardown(world)


C3_LEVELS = ((64, 96), (32, 48), (16, 24), (8, 12), (4, 6))             # p3 .. p7 of a 512x768 (0.5x) image
C3_D_SIZES = ((56, 88), (28, 44), (14, 22), (7, 11), (3, 5))


def run_stage2_c3(args):
    """BASELINE config 3 (stage-2 shapes on a BiFPN pyramid): 7 layers x 4 AF-interpolator fusion sites (p7 -> p3) with autograd incl. the
    input gradients, then the stage-2 loss block (stage2_trainer.py:298-384) on five levels: D phase (2 D forwards + backward per level),
    G phase (L1 + 2 D forwards), one backward through the whole top-down path."""
    import torch
    from afigan import native
    from afigan.engine import Stage2Step
    from afigan.modeling import bifpn_feature_fusion
    rank, world, local_rank, dev = _dist_setup()
    barrier, timed = _timing_tools(world, dev)
    N = PER_GPU_BATCH
    G, D = _models(args.precision, dev)
    G.deferred_weight_grads = True          # one packed gradient accumulator for the 28 calls of a backward pass (no per-call un-pack / adds)
    gen = torch.Generator().manual_seed(35 + rank)
    feats = [torch.randn(N, 256, h, w, generator=gen).to(dev).requires_grad_(True) for h, w in C3_LEVELS]
    wts = [torch.tensor([0.7, 1.3], device=dev, requires_grad=True) for _ in range(28)]
    guide = [torch.randn(N, 256, 2 * h + 1, 2 * w, generator=gen).to(dev) for h, w in C3_D_SIZES]
    lib = native.lib()
    s2 = Stage2Step(D, lr=1e-3, momentum=0.9, weight_decay=1e-4, precision=args.precision, distributed=False)     # grouped D / G loss block

    def one():
        outs = []
        k = 0
        for layer in range(7):
            top = feats[4]
            for l in (3, 2, 1, 0):
                top = bifpn_feature_fusion(G, feats[l], top, wts[k], swish=True); k += 1      # fusion + swish: one pass each way
                if layer == 6:
                    outs.append(top)
        model = [o[:, :, :h, :w] for o, (h, w) in zip(outs[::-1], C3_D_SIZES[:4])] + [feats[4][:, :, :3, :5]]
        s2.d_phase(guide, model)                                  # D forward x2 + backward over the five levels (grouped), all-reduce, SGD
        g_losses = s2.g_losses(guide, model)                      # adv (grouped D forwards, no gradient) + L1 with autograd
        sum(g_losses.values()).backward()
        for t in feats + wts + list(G.parameters()):
            t.grad = None

    for _ in range(max(args.warmup, 3)):
        one()
    lib.afi_launch_count(1)
    ms = timed(one, args.steps) / args.steps
    launches = int(lib.afi_launch_count(0))
    afi_px = N * 7 * sum(h * w for h, w in C3_LEVELS[1:])
    d_px = N * sum(h * w for h, w in C3_D_SIZES)
    flops = 3 * G_FWD_FLOP_PER_INPUT_PX * afi_px + (4 * D_FWD_FLOP_PER_PX + 2 * (2 * D_FWD_FLOP_PER_PX - 2_359_296)) * d_px
    tf = flops / ms / 1e9
    pk = peaks()
    if rank == 0:
        print(json.dumps(_line("stage2_bifpn_afi_step_img_per_s", world * N / (ms * 1e-3), "img/s", world, args, ms, args.precision,
                               "config 3: 28 AF-interpolator fusion sites (7 BiFPN layers x 4) fwd+bwd with autograd on a 64x96..4x6 pyramid + stage-2 "
                               "D/G loss block on 56x88..3x5, batch 2 per GPU",
                               per_gpu_batch=N, step_tflops_per_gpu=tf, gpu_launches=launches,
                               roofline={"bound": "tensor", "achieved": tf, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": tf / pk["bf16_sustained"],
                                         "traffic": None, "note": "small maps: launch / tile-latency bound (14 280 AFI pixels and 6 560 D pixels per image)"})))
    _dist_teardown(world)


def run_infer_c5(args):
    """BASELINE config 5: inference sweep of the BiFPN top-down path -- 28 AF-interpolator fusion sites per image (7 layers x 4), batch 1 per
    GPU, short side 400..1200 (long = short x 1333/800, padded to 128), images sharded over the GPUs (no collective)."""
    import torch
    from afigan import native
    from afigan.modeling import bifpn_feature_fusion
    rank, world, local_rank, dev = _dist_setup()
    barrier, timed = _timing_tools(world, dev)
    from afigan.modeling import Generator
    G, _ = _models(args.precision, dev)
    G.eval()
    # IN_FLIGHT - 1 further lanes (own workspace, own graphs, own stream, the same weights) for the "several images in flight" leg
    IN_FLIGHT = max(2, int(os.environ.get("AFIGAN_IN_FLIGHT", "4")))
    lanes = [G]
    for _ in range(IN_FLIGHT - 1):
        g2 = Generator(n_residual_dense_blocks=3, precision=args.precision).to(dev).eval()
        g2.load_state_dict(G.state_dict())
        lanes.append(g2)
    streams = [torch.cuda.Stream(device=dev) for _ in range(IN_FLIGHT)]
    wts = torch.tensor([0.7, 1.3], device=dev)
    lib = native.lib()
    sweep, tot_ms, tot_flops, tot_ms_k = [], 0.0, 0.0, 0.0
    for short in range(400, 1201, 100):
        long_ = (short * 1333 // 800 + 127) // 128 * 128
        short_p = (short + 127) // 128 * 128
        levels = [(short_p // s, long_ // s) for s in (8, 16, 32, 64, 128)]          # p3 .. p7
        gen = torch.Generator().manual_seed(36 + rank)
        feats = [torch.randn(1, 256, h, w, generator=gen).to(dev) for h, w in levels]

        def one(g=G):
            with torch.no_grad():
                for _ in range(7):
                    top = feats[4]
                    for l in (3, 2, 1, 0):
                        top = bifpn_feature_fusion(g, feats[l], top, wts)
            return top

        for _ in range(max(args.warmup, 3)):
            one()
        ms_calls = timed(one, args.steps) / args.steps        # one CUDA graph per interpolator call (what Generator.fuse does by itself)
        # ... and the whole neck pass of an image as ONE graph (what a serving loop with static shapes captures): no per-call input / output
        # copies, no graph-launch gaps between the 28 dependent calls
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            out = one()
        for _ in range(3):
            g.replay()
        ms = timed(g.replay, args.steps) / args.steps
        flops = 7 * G_FWD_FLOP_PER_INPUT_PX * sum(h * w for h, w in levels[1:])
        # ... and IN_FLIGHT images at once, one whole-image graph per lane on its own stream: a call on a coarse pyramid level occupies a
        # handful of SMs and is bound by the latency of its 21 dependent kernels, so independent images fill the rest of the GPU
        lane_graphs, lane_outs = [g], [out]
        for k in range(1, IN_FLIGHT):
            for _ in range(3):
                one(lanes[k])
            torch.cuda.synchronize()
            gk = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gk, stream=streams[k]):
                lane_outs.append(one(lanes[k]))
            lane_graphs.append(gk)
            gk.replay()
        torch.cuda.synchronize()
        assert all(torch.equal(o, out) for o in lane_outs[1:]), "lanes disagree"
        fork, joins = torch.cuda.Event(), [torch.cuda.Event() for _ in range(IN_FLIGHT)]

        def in_flight():
            fork.record()
            for k in range(IN_FLIGHT):
                streams[k].wait_event(fork)
                with torch.cuda.stream(streams[k]):
                    lane_graphs[k].replay()
                    joins[k].record()
            for k in range(IN_FLIGHT):
                torch.cuda.current_stream().wait_event(joins[k])

        for _ in range(3):
            in_flight()
        ms_k = timed(in_flight, args.steps) / args.steps / IN_FLIGHT
        sweep.append({"short_side": short, "padded": [short_p, long_], "ms_per_image": ms, "ms_per_image_graph_per_call": ms_calls,
                      "img_per_s": world * 1e3 / ms, "tflop_per_image": flops / 1e12, "tflops_per_gpu": flops / ms / 1e9,
                      "ms_per_image_%d_in_flight" % IN_FLIGHT: ms_k, "img_per_s_%d_in_flight" % IN_FLIGHT: world * 1e3 / ms_k,
                      "tflops_per_gpu_%d_in_flight" % IN_FLIGHT: flops / ms_k / 1e9})
        del g, out, lane_graphs, lane_outs
        tot_ms += ms; tot_flops += flops; tot_ms_k += ms_k
    pk = peaks()
    tf = tot_flops / tot_ms / 1e9
    if rank == 0:
        print(json.dumps(_line("bifpn_afi_inference_img_per_s", world * len(sweep) * 1e3 / tot_ms, "img/s", world, args, tot_ms / len(sweep), args.precision,
                               "config 5: 28 AF-interpolator fusion sites per image (7 BiFPN layers x 4), batch 1 per GPU, short side 400..1200, "
                               "images sharded over the GPUs; value = images of the whole sweep / time; each image's 28 calls replayed as one CUDA graph",
                               sweep=sweep, step_tflops_per_gpu=tf,
                               in_flight={"images": IN_FLIGHT, "img_per_s": world * len(sweep) * 1e3 / tot_ms_k, "tflops_per_gpu": tot_flops / tot_ms_k / 1e9,
                                          "note": "same sweep with %d independent images in flight per GPU (one stream + whole-image graph each); "
                                                  "`value` is the strict one-image-at-a-time reading of config 5" % IN_FLIGHT},
                               roofline={"bound": "tensor", "achieved": tf, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": tf / pk["bf16_sustained"],
                                         "traffic": None})))
    _dist_teardown(world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("AFIGAN_BENCH_PRECISION", "bf16"), choices=["bf16", "split", "fp32", "bf16_simt"])
    ap.add_argument("--workload", default="stage1", choices=["stage1", "g_only", "stage2_c3", "pafpn_c4", "infer_c5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the parity / parity_mode / extra blocks (N = 1 only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29517", os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps),
               "--warmup", str(args.warmup), "--precision", args.precision, "--workload", args.workload]
        sys.exit(subprocess.call(cmd))
    {"stage1": run_ours, "g_only": run_g_only, "stage2_c3": run_stage2_c3, "pafpn_c4": run_pafpn_c4, "infer_c5": run_infer_c5}[args.workload](args)


if __name__ == "__main__":
    main()
