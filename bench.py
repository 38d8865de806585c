#!/usr/bin/env python
"""Benchmark of the mean-shift + ellipsoid-fit hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg4|cfg1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A step = forward + backward of the fitting loss over one batch of synthetic shapes
(normalise x2 -> bandwidth -> T mean-shift iterations -> NMS -> K fp32 seed trajectories -> membership
-> ellipsoid fit -> SDF loss -> backward to the un-normalised embeddings).  Per-GPU work is fixed
(24 shapes of 2048 points on every rank: weak scaling; 8 GPUs = the 192-shape config), the only
collective is the 8-byte loss all-reduce.

Prints ONE JSON line on rank 0 (see README / DESIGN.md for the fields).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (shapes per GPU, points, quantile, iterations, max clusters, planted clusters)
    "cfg1": (1, 2048, 0.05, 10, 25, 16),
    "cfg2": (24, 2048, 0.05, 10, 25, 16),
    "cfg4": (16, 10000, 0.05, 10, 50, 16),
}
D = 128
N_SETS = 8          # rotating input sets: 8 x 25 MB of embeddings > 126 MB of L2


def algorithmic_flops_per_shape(N, T, K, passes=1):
    """G*2N^2 d(2T+2) + 12 T K N d + 8 K N d."""
    return passes * 2.0 * N * N * D * (2 * T + 2) + 12.0 * T * K * N * D + 8.0 * K * N * D


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of this workload (profiles/ncu_traffic.json); None when no capture exists."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(path))[workload][kernel]["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        return None


def hbm_stages(stage_events, B, N, K, hbm_peak_gbs):
    """Achieved HBM GB/s of the reduction stages (north_star: "achieved HBM GB/s against ~8 TB/s for the reductions"):
    CUDA events around every C-ABI call of the eager steps, ALGORITHMIC bytes per call (inputs read once + outputs written
    once; M = N, K = clusters found) / mean duration, against the measured copy bandwidth."""
    if not stage_events:
        return None
    d4 = D * 4
    alg = {
        "prifit_normalize_fwd": 2 * B * N * d4,
        "prifit_normalize_bwd": 3 * B * N * d4,
        "prifit_membership_fwd": B * (N * d4 + K * d4 + K * N * 4),
        "prifit_membership_bwd": B * (3 * N * d4 + 2 * K * N * 4 + K * d4),
        "prifit_fit_fwd": B * (K * N * 4 + N * 12),
        "prifit_fit_bwd": B * (2 * K * N * 4 + N * 12),
        "prifit_sdf_loss_fwd": B * (N * 12 + N * 8),
        "prifit_sdf_loss_bwd": B * (N * 12 + N * 4),
    }
    agg = {}
    for name, e0, e1 in stage_events:
        if name in alg:
            agg.setdefault(name, []).append(e0.elapsed_time(e1) * 1e3)
    out = {}
    for name, us in agg.items():
        mean_us = sum(us) / len(us)
        gbs = alg[name] / (mean_us * 1e-6) / 1e9
        out[name.replace("prifit_", "")] = {"us": round(mean_us, 2), "bytes": int(alg[name]), "GBps": round(gbs, 1),
                                             "frac": round(gbs / hbm_peak_gbs, 4), "calls": len(us)}
    out["_note"] = ("CUDA events around each C-ABI call (launch gaps of a few us included) on the eager path, 24-shape batch; "
                    "peak = measured copy bandwidth %.0f GB/s; these stages move <= 100 MB and are latency-, not bandwidth-bound" % hbm_peak_gbs)
    return out


def sub_bench(extra_args, keys):
    """Runs another workload of this script in a child process (fresh CUDA context, its own graphs) and returns the chosen
    keys of its JSON line -- so that the default line, the one the driver records, also carries cfg4 and cfg5."""
    try:
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--no-extras"] + extra_args, capture_output=True,
                             text=True, timeout=420)
        for ln in reversed(res.stdout.splitlines()):
            if ln.startswith("{"):
                d = json.loads(ln)
                return {k: d.get(k) for k in keys if k in d}
        return {"error": (res.stderr or res.stdout)[-300:]}
    except Exception as e:                                   # a sub-measurement must not take the headline line down
        return {"error": str(e)[:300]}


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            top = sorted(sm)[len(sm) // 2:]          # samples under load = upper half
            out = {"sm_mhz": statistics.median(top), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def bind_to_gpu_cpus(local):
    """Multi-GPU runs: bind this rank to the CPUs NVML reports as local to its GPU BEFORE any pinned buffer is allocated, so
    that the staging memory is first-touched on the GPU's NUMA node and the per-step H2D copy does not cross sockets (round 1:
    the copy time doubled at 4-8 ranks).  Best effort: a box without NVML, or one NUMA node, changes nothing."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else local
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        bind_to_gpu_cpus(local)
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


# ------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    from prifit_b200 import _lib, dist as pdist, graph_step, ops, pipeline, synthetic
    import prifit_b200.convex_loss as cl
    import torch.distributed as dist

    rank, world, local = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU arm")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    _lib.load()
    B, N, q, T, kmax, kc = WORKLOADS[args.workload]
    strong = args.global_shapes is not None
    if strong:
        # strong scaling (SURVEY 8d's extra): a fixed global batch cut over the ranks, one call per rank and step
        if args.global_shapes % world:
            raise SystemExit("--global-shapes must be a multiple of the number of GPUs")
        B = args.global_shapes // world
    engine = ops.MS_FP32_SIMT if args.engine == "fp32" else ops.MS_F16_TCGEN05
    ops.DEFAULT_ENGINE = engine

    # rotating synthetic input sets; rank r, set s uses shapes seeded (s*world + r) * B + b
    host_E, host_P = [], []
    n_sets = N_SETS if N <= 4096 else 2          # cfg4: two sets of 82 MB of embeddings already exceed the 126 MB L2
    if strong:
        n_sets = max(2, min(N_SETS, (160 << 20) // (B * N * D * 4) + 1))
    for s in range(n_sets):
        E, P, _ = synthetic.planted_shapes(B, n_points=N, n_clusters=kc, seed=1000 + (s * world + rank) * B)
        host_E.append(E.pin_memory()); host_P.append(P.pin_memory())
    dev_E = [e.to(dev) for e in host_E]
    dev_P = [p.to(dev) for p in host_P]
    # channel-first copies for the reference-shaped public API (convex_loss takes [B,128,N] / [B,3,N])
    host_Xcf = [e.permute(0, 2, 1).contiguous().pin_memory() for e in host_E]
    host_Pcf = [p.permute(0, 2, 1).contiguous().pin_memory() for p in host_P]
    torch.manual_seed(1234 + rank)

    ms_events = []
    last = {}

    use_graph = not args.no_graph
    if not use_graph:
        os.environ["PRIFIT_GRAPH"] = "0"                    # the public convex_loss() of the end-to-end arm follows it

    def step_resident(i, timed, graph=None):
        E = dev_E[i % n_sets].detach().requires_grad_(True)
        graph = use_graph if graph is None else graph
        ops.TIMING = ms_events if (timed and not graph) else None
        out = pipeline.fit_loss(E, dev_P[i % n_sets], quantile=q, iterations=T, max_num_clusters=kmax, graph=graph,
                                dist_reduce=world > 1)
        ops.TIMING = None
        L, Lb = pdist.global_loss(out)
        Lb.backward()
        return L, out

    # End-to-end arm: every step copies its inputs host -> device from pinned memory and reads its loss back.
    # Like a data loader would, the copy of step i+1 is issued on a side stream while step i computes, and the
    # loss of step i is read (pinned D2H) while step i+1 is being enqueued; both stay inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    read_stream = torch.cuda.Stream(device=dev)
    staged = {}
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    pending = {}

    h2d_events = []

    # two device-side input slots, filled alternately (no allocator traffic in the loop); a slot is refilled only after the
    # backward of the step that read it (two steps earlier) is through
    in_slots = [(torch.empty_like(host_Xcf[0], device=dev), torch.empty_like(host_Pcf[0], device=dev)) for _ in range(2)]
    slot_free = [None, None]

    def stage_inputs(i):
        Xb, Pb = in_slots[i & 1]
        with torch.cuda.stream(copy_stream):
            graph_step.gate_on_cluster_stage(copy_stream)   # behind the running step's all-seed kernel (L2), whatever the host's timing
            if slot_free[i & 1] is not None:
                copy_stream.wait_event(slot_free[i & 1])
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record(copy_stream)
            Xb.copy_(host_Xcf[i % n_sets], non_blocking=True)
            Pb.copy_(host_Pcf[i % n_sets], non_blocking=True)
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(copy_stream)
        h2d_events.append((e0, ev))
        staged[i] = (Xb, Pb, ev)

    def drain_loss():
        if "ev" in pending:
            pending.pop("ev").synchronize()
            last["e2e_loss"] = float(loss_host[pending.pop("slot")])

    def read_back_previous():
        """Device -> host read of the PREVIOUS step's loss (4 bytes, pinned), on a stream of its own.  Called from the
        enqueued hook of the next step (or at the end of the loop): the ~70 us of host work (events, stream switch, copy,
        the wait for the copy before it) then sit in the shadow of the device's cluster stage instead of between one step's
        backward and the next step's launch, where the device only has ~0.5 ms of latency chains queued."""
        drain_loss()                                 # the read issued one call earlier has long landed
        if "L" in pending:
            L, ev_l, slot = pending.pop("L"), pending.pop("ev_l"), pending.pop("slot_next")
            with torch.cuda.stream(read_stream):
                read_stream.wait_event(ev_l)
                loss_host[slot:slot + 1].copy_(L.detach().reshape(1), non_blocking=True)
                ev2 = torch.cuda.Event()
                ev2.record(read_stream)
            pending["ev"], pending["slot"] = ev2, slot

    host_marks = {"n": 0, "stage+wait": 0.0, "convex_loss": 0.0, "backward": 0.0, "hook": 0.0}

    def step_e2e(i, last_step):
        t0 = time.perf_counter()
        if i not in staged:
            stage_inputs(i)
        X, pts, ev = staged.pop(i)
        torch.cuda.current_stream().wait_event(ev)
        X = X.detach().requires_grad_(True)
        # Prefetch of the next step's inputs, like a data loader would: issued from graph_step.enqueued_hook, i.e. once this
        # step's device work and its own small host->device copy (the staged noise draws) are enqueued and before the host
        # blocks on the guard inputs.  The H2D copy engine serves one queue: 25 MB in front of the 27 KB noise copy would
        # stall the step by ~0.2 ms, and a prefetch issued after convex_loss() returns starts two thirds of a step late and
        # is still running when the next step wants its input.
        fired = []

        def hook():
            th = time.perf_counter()
            fired.append(1)
            if not last_step:
                stage_inputs(i + 1)
            read_back_previous()
            host_marks["hook"] += time.perf_counter() - th
        graph_step.enqueued_hook = hook
        t1 = time.perf_counter()
        total, l, params, labels = cl.convex_loss(pts, pts, X, quantile=q, iterations=T, max_num_clusters=kmax,
                                                  dist_reduce=world > 1, full_chamfer=False)
        graph_step.enqueued_hook = None
        if not fired:                                # eager path (a guard redo, PRIFIT_GRAPH=0): no hook call
            hook()
        t2 = time.perf_counter()
        total.backward()
        t3 = time.perf_counter()
        host_marks["n"] += 1; host_marks["stage+wait"] += t1 - t0; host_marks["convex_loss"] += t2 - t1; host_marks["backward"] += t3 - t2
        ev_l = torch.cuda.Event()
        ev_l.record()
        slot_free[i & 1] = ev_l                      # this step's input slot may be refilled once its backward is through
        # l = the global mean (one 8-byte all-reduce inside convex_loss when N > 1)
        pending["L"], pending["ev_l"], pending["slot_next"] = l, ev_l, i & 1
        if last_step:
            read_back_previous()
            drain_loss()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(fn, steps, warmup):
        for i in range(warmup):
            fn(i, False)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            fn(warmup + i, True)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1), wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1])

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.reset_launch_count()

    def resident(i, timed):
        L, out = step_resident(i, timed)
        last["L"], last["out"] = L, out

    ms_total, _ = timed_region(resident, args.steps, args.warmup)
    launches_total = _lib.launch_count()
    # The dominant kernel is timed with CUDA events around its launch.  Inside a graph replay (parallel branches)
    # a single kernel's duration is not observable, so when the timed region ran as graphs the kernel is timed in a
    # second region of the same steps on the eager path (same kernel, same inputs, one stream).
    eager_ms = None
    stage_events = []
    if use_graph:
        def eager(i, timed):
            _lib.TIMING = stage_events if timed else None      # CUDA events around every C-ABI call of the eager steps
            L, out = step_resident(i, timed, graph=False)
            _lib.TIMING = None
        eager_total, _ = timed_region(eager, min(args.steps, 10), 3)
        eager_ms = eager_total / min(args.steps, 10)
    n_ms_launch = len(ms_events)
    ms_kernel_in_step = sum(a.elapsed_time(b) for a, b in ms_events) / max(n_ms_launch, 1)
    # The bracket inside an eager step also holds whatever the host needs between the fp16 conversion launch and the kernel
    # launch (tensor-map encode, Python); the launch duration proper is taken from back-to-back launches of the same C-ABI
    # call on the last step's own unit embeddings and bandwidths, host running ahead, CUDA events on the launching stream.
    Xu, bwu = last["out"]["X"], last["out"]["cluster"].bw
    for _ in range(3):
        ops.meanshift(Xu, bwu, T, engine)
    torch.cuda.synchronize()
    kev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n_ms_launch = 10
    kev[0].record()
    for _ in range(n_ms_launch):
        ops.meanshift(Xu, bwu, T, engine)
    kev[1].record()
    torch.cuda.synchronize()
    ms_kernel = kev[0].elapsed_time(kev[1]) / n_ms_launch
    # diagnostic: the same public-API step on inputs already resident in HBM (what the H2D staging adds on top)
    dev_Xcf = [x.to(dev) for x in host_Xcf[:2]]
    dev_Pcf = [p.to(dev) for p in host_Pcf[:2]]

    def step_api_resident(i, timed):
        X = dev_Xcf[i % 2].detach().requires_grad_(True)
        total, l, params, labels = cl.convex_loss(dev_Pcf[i % 2], dev_Pcf[i % 2], X, quantile=q, iterations=T, max_num_clusters=kmax,
                                                  full_chamfer=False)
        total.backward()

    api_ms = None
    if world == 1:
        api_total, _ = timed_region(step_api_resident, min(args.steps, 10), 3)
        api_ms = api_total / min(args.steps, 10)
    e2e_last = args.warmup + args.steps - 1
    e2e_ms, _ = timed_region(lambda i, timed: step_e2e(i, i == e2e_last or i == args.warmup - 1), args.steps, args.warmup)
    if args.trace_e2e and world == 1:                # (one rank tracing alone would leave the others' collectives hanging)
        # diagnostics only (after every measurement): device timeline of the end-to-end loop under CUPTI
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        from device_timeline import print_timeline
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for i in range(8):
                step_e2e(i, i == 7)
            torch.cuda.synchronize()
        with open(args.trace_e2e, "w") as f:
            print_timeline(prof, out=f, first_kernel="normalize_cf_kernel<false>", last_kernel="normalize_cf_kernel<true>", which=4)
    clocks = sampler.stop() if rank == 0 else None

    # sustained: >= 3 s of back-to-back steps (the 20-step figure above is a burst at full clocks)
    sustained = None
    if not args.no_extras:
        s2 = ClockSampler(local)
        if rank == 0:
            s2.start()
        n_sus = max(50, int(3.2 / max(ms_total / args.steps * 1e-3, 1e-4)))
        sus_ms, _ = timed_region(resident, n_sus, 3)
        c2 = s2.stop() if rank == 0 else None
        sustained = {"value": round(world * B * n_sus / (sus_ms * 1e-3), 2), "unit": "shapes/s", "steps": n_sus,
                     "seconds": round(sus_ms * 1e-3, 2), "ms_per_step": round(sus_ms / n_sus, 4), "clocks": c2}

    res = last["out"]["cluster"]
    K_mean = sum(res.K_host) / len(res.K_host)
    passes = sum(res.passes) / len(res.passes)
    ms_per_step = ms_total / args.steps
    shapes_per_s = world * B / (ms_per_step * 1e-3)
    e2e_sps = world * B / (e2e_ms / args.steps * 1e-3)
    pk = peaks()
    # dominant kernel: the all-seed mean-shift pass (2 GEMMs x T x N^2 d per shape)
    flops_launch = 4.0 * N * N * D * T * B
    # kind::f16 runs at the bf16 rate.  The timed region is tens of milliseconds at full clocks (see `clocks`), so the
    # like-for-like denominator is the BURST cuBLAS figure; the fraction of the sustained one is reported beside it.
    tc_peak = pk["bf16_tflops"]
    achieved = flops_launch / (ms_kernel * 1e-3) / 1e12 if ms_kernel > 0 else 0.0
    line = {
        "metric": "shapes/sec mean-shift+ellipsoid fit fwd+bwd (2048 pts)" if N == 2048 else
                  "shapes/sec mean-shift+ellipsoid fit fwd+bwd (%d pts)" % N,
        "value": round(shapes_per_s, 2), "unit": "shapes/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f32 (all-seed mean-shift GEMMs: %s)" % ("f16 operands / f32 accumulate, tcgen05" if engine == ops.MS_F16_TCGEN05 else "f32 simt"),
        "data": "synthetic",
        "config": {"workload": "%s: %d shapes x %d pts x %d-d per GPU, T=%d, quantile=%g, max_num_clusters=%d, "
                               "%d planted clusters (S1)" % (args.workload, B, N, D, T, q, kmax, kc),
                   "shapes_per_gpu": B, "global_shapes": world * B, "clusters_found_mean": K_mean,
                   "guard_passes_mean": passes, "loss": float(last["L"]),
                   "l2": "rotating %d input sets (%.0f MB of embeddings) > 126 MB L2" % (n_sets, n_sets * B * N * D * 4 / 1e6),
                   "parallelism": "shapes sharded %d/GPU, one 8-byte NCCL all-reduce per step" % B,
                   "execution": ("1 CUDA graph per step (every branch's forward with its speculative backward behind it%s) + the two normalisation kernels, %d parallel branches of shapes"
                                 % ("; multi-GPU: forward graph | all-reduce beside the backward graph" if world > 1 else "", graph_step.default_branches())) if use_graph
                                else "eager launches on one stream",
                   "eager_ms_per_step": None if eager_ms is None else round(eager_ms, 4)},
        "e2e": {"value": round(e2e_sps, 2), "unit": "shapes/s", "ms_per_step": round(e2e_ms / args.steps, 4),
                "h2d_bytes_per_step": int(host_Xcf[0].numel() * 4 + host_Pcf[0].numel() * 4), "d2h_bytes_per_step": 4,
                "api_resident_ms_per_step": None if api_ms is None else round(api_ms, 4),
                "h2d_copy_ms_per_step": round(statistics.median(a.elapsed_time(b) for a, b in h2d_events[-args.steps:]), 4),
                "host_us_per_step": {k: round(v / max(host_marks["n"], 1) * 1e6, 1) for k, v in host_marks.items() if k != "n"},
                "api": "prifit_b200.convex_loss.convex_loss(points[B,3,N], chamfer[B,3,N], X[B,128,N], full_chamfer=False) + backward; "
                       "full_chamfer=False = the SDF half of the fitting loss, the term the metric is defined on (SURVEY 8d); the 25 MB input "
                       "gradient stays on the device (training), only the 4-byte loss is read back"},
        "gpu_launches": int(round(launches_total / (args.steps + args.warmup) * args.steps)),   # kernels + memset nodes, graph nodes included
        "gpu_launches_per_step": round(launches_total / (args.steps + args.warmup), 1),
        "roofline": {"bound": "tensor", "kernel": "meanshift_fwd (%s)" % args.engine, "achieved": round(achieved, 2),
                     "peak": round(tc_peak, 1), "unit": "TFLOP/s", "frac": round(achieved / tc_peak, 4),
                     "frac_of_sustained_peak": round(achieved / pk["bf16_tflops_sustained"], 4),
                     "traffic": ncu_traffic(args.workload, "meanshift_tc_kernel") if engine == ops.MS_F16_TCGEN05 else None,
                     "traffic_source": "static: dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu --set full capture "
                                       "(profiles/ncu_traffic.json), not measured in this run",
                     "kernel_ms": round(ms_kernel, 4), "launches_timed": n_ms_launch,
                     "kernel_timed_in": "10 back-to-back launches of prifit_meanshift_fwd (fp16 conversion + kernel) on the last step's unit "
                                        "embeddings and bandwidths, after the timed region (graph replays run it in parallel branches)",
                     "kernel_ms_in_eager_step": round(ms_kernel_in_step, 4),
                     "flops_per_launch": flops_launch,
                     "peak_source": "%s dense bf16 GEMM, burst (%.0f TF/s; sustained %.0f)" % (pk["source"], pk["bf16_tflops"], pk["bf16_tflops_sustained"]),
                     "whole_step_tflops": round(shapes_per_s / world * algorithmic_flops_per_shape(N, T, K_mean, passes) / 1e12, 2)},
        "clocks": clocks,
    }
    if sustained is not None:
        line["sustained"] = sustained
    hbm = hbm_stages(stage_events, B, N, K_mean, pk["hbm_gbs"])
    if hbm:
        line["roofline"]["hbm"] = hbm
    if rank == 0 and world == 1 and not args.no_extras and args.workload == "cfg2":
        line["cfg4"] = sub_bench(["--workload", "cfg4", "--steps", "10", "--warmup", "3"],
                                 ("value", "ms_per_step", "config", "roofline", "clocks"))
        line["cfg5"] = sub_bench(["--workload", "cfg5", "--steps", "10", "--warmup", "3"],
                                 ("value", "ms_per_step", "config", "e2e", "clocks"))
    if rank == 0:
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload, shapes=min(B, 12))
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------- cfg5
def run_cfg5(args):
    """BASELINE.json configs[4]: one self-supervised training step of the reference's OWN PointNet++ MSG part-segmentation
    model (models/pointnet2_part_seg_msg.get_model(50), unmodified, from the hosted reference tree) with the fitting loss --
    forward, backward, Adam step -- on synthetic ShapeNet-shaped batches of 24 shapes per GPU: points[24,3,2048] +
    chamfer_points[24,3,5000], quantile 0.05, 10 iterations, <= 25 clusters (README run).  The model calls convex_loss exactly
    as the reference does (no extension argument): the objective is the reference's complete analytic chamfer distance.
    N > 1: DistributedDataParallel over NCCL (7 MB of gradients), one process per GPU."""
    import prifit_b200.reference_host as host
    import torch.distributed as dist

    rank, world, local = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if host.find_tree() is None:
        if rank == 0:
            print(json.dumps({"metric": "shapes/sec PointNet++ MSG part-seg self-supervised step with the fitting loss", "value": None,
                              "unavailable": "no copy of the reference tree (baseline/_ref is made by __graft_entry__.build() where /root/reference exists)"}))
        return
    B, q, T, kmax = 24, 0.05, 10, 25
    model, module = host.build_partseg_model(dev, seed=1 + rank)
    net = model
    if world > 1:
        for p_ in model.parameters():                        # identical initial weights on every rank
            dist.broadcast(p_.data, 0)
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-08, weight_decay=1e-4)   # train_partseg_shapenet.py:277-283
    net.train()
    sets = [host.synthetic_partseg_batch(B, seed=100 + 10 * s + rank) for s in range(4)]
    host_sets = [tuple(t.pin_memory() for t in st) for st in sets]
    fit_events, last = [], {}
    orig_cl = module.convex_loss

    def timed_cl(*a, **k):                                   # CUDA events around the fitting loss inside the model's forward
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_cl(*a, **k)
        e1.record()
        fit_events.append((e0, e1))
        return out

    module.convex_loss = timed_cl

    def step(i, timed):
        pts, ch, cls = (t.to(dev, non_blocking=True) for t in host_sets[i % 4])        # pinned H2D inside the timed region
        out, loss = host.partseg_selfsup_step(net, opt, pts, ch, cls, quantile=q, msc_iterations=T, max_num_clusters=kmax)
        last["loss"] = float(loss)                           # device -> host read of the step's result, every step (the training
        last["K"] = [len(p) for p in out[6]]                 # loop of the reference logs it every step too)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i, False)
    barrier()
    fit_events.clear()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i, True)
    loss_host = last["loss"]
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = float(ms[0]) / args.steps
    fit_ms = sum(a.elapsed_time(b) for a, b in fit_events) / max(len(fit_events), 1)
    sps = world * B / (ms_step * 1e-3)
    if rank == 0:
        print(json.dumps({
            "metric": "shapes/sec PointNet++ MSG part-seg self-supervised step with the fitting loss (2048 pts)",
            "value": round(sps, 2), "unit": "shapes/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (backbone: torch fp32; fitting loss: this package)", "data": "synthetic",
            "config": {"workload": "cfg5: reference models/pointnet2_part_seg_msg.get_model(50) (1.76 M parameters, unmodified, from the hosted "
                                   "reference tree) + convex_loss as the reference calls it, 24 shapes x 2048 pts (+ 5000 chamfer pts) per GPU, "
                                   "T=10, quantile=0.05, max_num_clusters=25, Adam step",
                       "fitting_loss_forward_ms": round(fit_ms, 3), "backbone_and_backward_ms": round(ms_step - fit_ms, 3),
                       "clusters_found": last["K"][:4], "loss": loss_host,
                       "objective": "reference analytic_chamfer_distance (sampled-surface half + SDF half), PRIFIT_FULL_CHAMFER=%s" % os.environ.get("PRIFIT_FULL_CHAMFER", "1"),
                       "parallelism": "DistributedDataParallel, %d ranks" % world if world > 1 else "single GPU",
                       "l2": "4 rotating host batches, copied host -> device every step"},
            "e2e": {"value": round(sps, 2), "unit": "shapes/s", "ms_per_step": round(ms_step, 3),
                    "h2d_bytes_per_step": int(sum(t.numel() * 4 for t in host_sets[0])), "d2h_bytes_per_step": 4,
                    "api": "reference model forward(include_convex_loss=True) + backward + optimizer.step(), host tensors in"},
            "clocks": clocks}))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------ reference arm
def _cpu_step():
    """-> (step(E, P, q, T, kmax), kind, restore()).  kind "reference": the UNMODIFIED reference's own functions (a copy
    of its tree is reachable: /root/reference in the build container, baseline/_ref on the benchmark box -- put there by
    __graft_entry__.build()), run on the CPU with its hard-coded .cuda() calls made the identity; kind "port": the oracle
    restatement (bit-identical in fp32 to the reference, oracle/make_golden.py) when no copy is there."""
    from oracle import ref_loader
    if ref_loader.available():
        try:
            orig = torch.Tensor.cuda
            ns = ref_loader.load(force_cpu=True)
            from oracle import ref_runner

            def restore():
                torch.Tensor.cuda = orig

            return (lambda E, P, q, T, kmax: ref_runner.ref_fit_loss(ns, E, P, q, T, kmax)), "reference", restore
        except Exception as e:                                  # a broken copy must not take the bench line down
            sys.stderr.write("reference tree found but not importable (%s); timing the oracle port\n" % e)
    from oracle import restatement as R
    return (lambda E, P, q, T, kmax: R.fit_loss(E, P, q, T, kmax)), "port", (lambda: None)


def cpu_baseline(workload, shapes, budget_s=12.0):
    """The reference's CPU implementation of the path (see _cpu_step) timed on the host cores on a bounded sample of the
    same workload: whole passes over `shapes` shapes of one step, fwd+bwd, repeated until ~budget_s seconds of CPU work
    have been measured."""
    from prifit_b200 import synthetic
    B, N, q, T, kmax, kc = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if N > 4096:
        shapes = 1
    step, kind, restore = _cpu_step()
    try:
        E, P, _ = synthetic.planted_shapes(shapes, n_points=N, n_clusters=kc, seed=1000)
        step(E[:1], P[:1], q, T, kmax)                             # warm-up
        done, t0 = 0, time.perf_counter()
        while True:
            step(E, P, q, T, kmax)
            done += shapes
            dt = time.perf_counter() - t0
            if dt >= budget_s:
                break
    finally:
        restore()
    return {"value": round(done / dt, 3), "unit": "shapes/s", "cores": cores, "threads": torch.get_num_threads(),
            "kind": kind, "sample": "%d shapes (%d of the %d shapes of one step, %d passes), fwd+bwd, %.1f s" % (
                done, shapes, B, done // shapes, dt)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (see _cpu_step), all host threads.  A step =
    fwd+bwd over the step's shapes (cfg2: the full 24-shape batch when the whole run fits a few minutes, else a bounded
    sample; cfg4: one 10000-point shape); rank 0 only."""
    from prifit_b200 import synthetic
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, N, q, T, kmax, kc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind, _ = _cpu_step()
    E, P, _ = synthetic.planted_shapes(max(B, 2), n_points=N, n_clusters=kc, seed=1000)
    t0 = time.perf_counter()
    step(E[:1], P[:1], q, T, kmax)                                 # first call: also tells how long one shape takes
    step(E[:1], P[:1], q, T, kmax)
    per_shape = (time.perf_counter() - t0) / 2
    budget = 150.0                                                 # seconds for warm-up + timed steps
    per_step = int(max(1, min(B, budget / max(per_shape * (args.steps + args.warmup), 1e-9))))
    for i in range(args.warmup):
        step(E[:per_step], P[:per_step], q, T, kmax)
    t0 = time.perf_counter()
    for i in range(args.steps):
        o = (i * per_step) % max(E.shape[0] - per_step + 1, 1)
        step(E[o:o + per_step], P[o:o + per_step], q, T, kmax)
    dt = time.perf_counter() - t0
    sps = args.steps * per_step / dt
    sample = "%d shapes per step (of %d), fwd+bwd, %s on %d threads" % (
        per_step, B, "the unmodified reference (baseline/_ref)" if kind == "reference" else "oracle port", torch.get_num_threads())
    print(json.dumps({
        "impl": "reference", "metric": "shapes/sec mean-shift+ellipsoid fit fwd+bwd (%d pts)" % N,
        "value": round(sps, 3), "unit": "shapes/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        # the same workload string as the CUDA arm's line; what the CPU arm actually timed per step is in `sample`
        "config": {"workload": "%s: %d shapes x %d pts x %d-d per GPU, T=%d, quantile=%g, max_num_clusters=%d, "
                               "%d planted clusters (S1)" % (args.workload, B, N, D, T, q, kmax, kc),
                   "shapes_per_gpu": B, "sample": sample},
        "cpu_baseline": {"value": round(sps, 3), "unit": "shapes/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(sps, 3), "unit": "shapes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS) + ["cfg5"])
    ap.add_argument("--global-shapes", type=int, default=None, metavar="G",
                    help="strong scaling: G shapes in total, G / gpus per rank and step (default: the workload's per-GPU batch, weak scaling)")
    ap.add_argument("--no-extras", action="store_true", help="skip the sustained / cfg4 / cfg5 sub-measurements of the default line")
    ap.add_argument("--engine", default="tcgen05", choices=["tcgen05", "fp32"])
    ap.add_argument("--trace-e2e", default=None, metavar="FILE", help="write a device timeline of the end-to-end loop (diagnostics)")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replays")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="diagnostics (A/B runs of two builds): skip the 12 s CPU leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        if args.workload == "cfg5":
            args.workload = "cfg2"       # the reference's full convex_loss needs trimesh on the CPU: its hot path is what can be timed
        run_reference(args)
    elif args.workload == "cfg5":
        run_cfg5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
