#!/usr/bin/env python
"""bench.py -- loss fwd+bwd pairs/s of the sparsify-clip contrastive-loss hot path on B200.

Workload (BASELINE.json metric, config c3): experiment_3 composition
    loss = anchor(I, T, tau=0.1) + lalign(I, T) + (lunif(I) + lunif(T)) / 2
forward + backward (gradients w.r.t. I and T materialised), global B = 32768, D = 512, bf16
unit-norm synthetic embeddings.  At N > 1 the rows are sharded over the ranks ("strong" scaling:
the global batch is fixed, as BASELINE.json quotes it), one process per GPU under torchrun.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One JSON line on stdout (rank 0).  `value` = global B / device time per step with inputs resident in
HBM; `e2e` = the same step through the public API from pinned HOST buffers (H2D copy of I and T and
D2H read of the loss inside the timed region); `roofline` = algorithmic FLOPs of the B x B passes /
their summed CUDA-event duration against the measured bf16 peak; `cpu_baseline` = the torch port of
the reference op sequence (oracle/torch_port.py) on the host cores, on a bounded sample.
`--impl reference` times that CPU path alone and prints the same line shape.
"""
import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "loss_fwd_bwd_pairs_per_s"
UNIT = "pairs/s"
WEIGHTS_EXP3 = dict(anchor=1.0, align=1.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0, alpha=0.0, beta=0.0)
TAU = 0.1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32768, help="global batch B")
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--cpu-sample-batch", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tc-flags", type=int, default=None, help="debug: scb_set_tc_flags value")
    ap.add_argument("--graph", action="store_true", help="replay the step from a captured CUDA graph (measured: no gain "
                                                         "at c3, the step is kernel-bound; useful at small B)")
    return ap.parse_args()


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        k = d["per_launch_dram_bytes"][d["dominant"]]
        return float(k["read"] + k["write"])
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
    return 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            t_end = time.time() + 5.0          # NVML initialisation done = first sample delivered
            while not self.rows and time.time() < t_end and self.proc.poll() is None:
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self, first_row=0):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[first_row:]:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_reference_step_time(batch, dim, steps, warmup):
    """torch port of the reference op sequence, fp32, all host threads; returns (median s/step, threads)."""
    import torch
    from oracle import torch_port
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(42)
    I = torch.nn.functional.normalize(torch.randn(batch, dim, generator=g), dim=-1)
    T = torch.nn.functional.normalize(I + 0.5 * torch.randn(batch, dim, generator=g), dim=-1)
    w = (1.0, 1.0, 0.5, 0.5, 0.0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        torch_port.fwd_bwd(I, T, TAU, w)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return statistics.median(times), torch.get_num_threads()


def cpu_baseline_record(args, steps=3, warmup=1):
    bs = min(args.cpu_sample_batch, args.batch)
    t, threads = cpu_reference_step_time(bs, args.dim, steps, warmup)
    t_full = t * (args.batch / bs) ** 2          # every term but L_align is O(B^2 D)
    return {"value": args.batch / t_full, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"oracle/torch_port.py exp-3 fwd+bwd fp32 at B={bs}, D={args.dim}: median {t:.3f} s/step over "
                      f"{steps} steps = {bs / t:.1f} pairs/s measured; value extrapolated to B={args.batch} by (B/{bs})^2"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    bs = min(args.cpu_sample_batch, args.batch)
    t, threads = cpu_reference_step_time(bs, args.dim, steps, warmup)
    t_full = t * (args.batch / bs) ** 2
    value = args.batch / t_full
    sample = (f"each step = oracle/torch_port.py exp-3 fwd+bwd fp32 on a B={bs} sample (median {t:.3f} s); "
              f"value and ms_per_step extrapolated to B={args.batch} by (B/{bs})^2")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": t_full * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"c3: experiment_3 anchor+lalign+(lunif(img)+lunif(txt))/2, B={args.batch}, D={args.dim}, "
                                   f"tau={TAU}, CPU torch port of the reference ops"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import sparsify_clip_b200 as scb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    be = scb.get_backend()
    if args.tc_flags is not None:
        be.lib.scb_set_tc_flags(args.tc_flags)

    B, D = args.batch, args.dim
    assert B % world == 0, "global batch must divide over the ranks"
    n = B // world
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    I0 = torch.nn.functional.normalize(torch.randn(n, D, generator=g, device=dev), dim=-1)
    T0 = torch.nn.functional.normalize(I0 + 0.5 * torch.randn(n, D, generator=g, device=dev), dim=-1)
    I = I0.to(torch.bfloat16).requires_grad_(True)
    T = T0.to(torch.bfloat16).requires_grad_(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step(Iv, Tv):
        Iv.grad = None
        Tv.grad = None
        loss = scb.weighted_loss(Iv, Tv, TAU, WEIGHTS_EXP3, group=group)
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # nvidia-smi is started BEFORE the warm-up: its first start on a fresh box (NVML initialisation) was measured to stall
    # the GPU for ~0.1 s, which used to land inside the timed region of the first bench process (2x outliers).
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step(I, T)
    barrier()
    # Settling: on a fresh box the first process was measured up to 2x slow for its first ~0.2 s (clock / power-state
    # ramp), which a fixed handful of warm-up steps does not cover.  Keep stepping, untimed, until five consecutive
    # steps agree within 3 % (at most ~2 s), identically on every rank.
    hist = []
    for _ in range(160):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(I, T)
        e1.record()
        torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        hist.append(tt.item())
        if len(hist) >= 8 and (max(hist[-5:]) - min(hist[-5:])) <= 0.03 * min(hist[-5:]):
            break
    barrier()

    # The whole step (about 55 launches, most of them tiny) is captured once into a CUDA graph and replayed: the
    # library is capturable by contract (no allocation, no host sync, caller's stream).  Collectives stay eager.
    graphed = None
    if args.graph and world == 1:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step(I, T)
            torch.cuda.current_stream().wait_stream(side)
            gr = torch.cuda.CUDAGraph()
            I.grad = None
            T.grad = None
            with torch.cuda.graph(gr):
                g_loss = scb.weighted_loss(I, T, TAU, WEIGHTS_EXP3, group=group)
                g_loss.backward()
            graphed = (gr, g_loss)
            gr.replay()
            torch.cuda.synchronize()
        except Exception as exc:      # fall back to eager timing, and say so
            sys.stderr.write(f"bench.py: CUDA-graph capture failed ({exc}); timing eager launches\n")
            graphed = None

    def timed_step():
        if graphed is None:
            return step(I, T)
        graphed[0].replay()
        return graphed[1]

    # One untimed back-to-back burst of the same length as the timed loop: with the CPU running ahead of the GPU several
    # steps' worth of buffers are alive at once, and the caching allocator must have grown to that peak BEFORE the timed
    # loop (a cudaMalloc inside it synchronises the device: measured as sporadic +30 % outliers).
    # A third source, seen as ONE step of 64 ms (or +2 ms) right after a barrier, when the launch queue is empty and the
    # host cannot hide anything: a full collection of the Python garbage collector over the import-time heap (torch,
    # numpy, ...).  Collect now and freeze the survivors, so that collections inside the timed region only look at
    # the objects the steps themselves create (what `timeit` achieves by switching the collector off).
    gc.collect()
    gc.freeze()
    for _ in range(args.steps):
        flush.zero_()
        step(I, T)
    barrier()

    # ---- timed region: K steps, each bracketed by CUDA events, L2 flushed (untimed) in between
    first_row = len(sampler.rows)      # only samples taken from here on (the timed region) are reported
    gc0 = [g["collections"] for g in gc.get_stats()]
    evs = []
    launches0 = be.launches
    barrier()
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = timed_step()
        e1.record()
        evs.append((e0, e1))
    barrier()
    launches = be.launches - launches0
    gc_in_region = [g["collections"] - a for g, a in zip(gc.get_stats(), gc0)]
    # A short timed region yields too few nvidia-smi samples (100 ms period; denser polling was measured to stall the
    # GPU: outlier steps of 2x): keep the same load running, untimed, for ~0.6 s more.  Rank 0 owns the sampler and
    # decides; the step count is broadcast so that every rank runs the same collectives.
    topup = torch.zeros(1, device=dev, dtype=torch.int64)
    if rank == 0 and len(sampler.rows) - first_row < 3:
        local_ms = sum(a.elapsed_time(b) for a, b in evs) / max(1, args.steps)
        topup[0] = max(1, int(600.0 / max(0.05, local_ms)))
    if world > 1:
        dist.broadcast(topup, src=0)
    for _ in range(int(topup.item())):
        step(I, T)
    torch.cuda.synchronize()
    if graphed is not None:            # replays do not pass through the Python launch counter: count one eager step
        l0 = be.launches
        step(I, T)
        launches = (be.launches - l0) * args.steps
    clocks = sampler.stop(first_row) if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in evs]          # this rank's per-step intervals (reported for transparency)
    total_ms = sum(step_ms)
    tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_per_step = tt.item() / args.steps
    value = B / (ms_per_step * 1e-3)
    loss_val = float(loss.item())

    # ---- per-pass CUDA-event timing (separate instrumented steps) for the roofline line
    be.pass_events = []
    prof_steps = 3
    for _ in range(prof_steps):
        flush.zero_()
        step(I, T)
    torch.cuda.synchronize()
    per = {}
    for name, a, b in be.pass_events:
        per.setdefault(name, []).append(a.elapsed_time(b))
    be.pass_events = None
    pass_ms = sum(sum(v) for v in per.values()) / prof_steps
    pm = torch.tensor([pass_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(pm, op=dist.ReduceOp.MAX)
    pass_ms = pm.item()
    flops_alg = 14.0 * B * B * D                       # anchor 6 + 2 x lunif 4 (B^2 D each)
    peak, peak_src = peaks()
    achieved = flops_alg / world / (pass_ms * 1e-3) / 1e12      # per-GPU TFLOP/s of the B x B passes

    # ---- end to end through the public API from pinned host memory
    hI = I0.to(torch.bfloat16).cpu().pin_memory()
    hT = T0.to(torch.bfloat16).cpu().pin_memory()
    # Input pipeline: the H2D copy of step k+1 runs on a copy stream into the other of two device buffers while step k
    # computes (what a training loop's prefetcher does); every step's copy, its compute and the D2H read of its loss lie
    # inside one timed region that spans all steps, the first copy fully exposed.
    dbuf = [(torch.empty(n, D, dtype=torch.bfloat16, device=dev), torch.empty(n, D, dtype=torch.bfloat16, device=dev))
            for _ in range(2)]
    hloss = torch.empty((), dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream()
    ev_in = [torch.cuda.Event(), torch.cuda.Event()]
    ev_free = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(k):
        bsel = k & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[bsel])          # the step that last read this buffer has finished
            dbuf[bsel][0].copy_(hI, non_blocking=True)
            dbuf[bsel][1].copy_(hT, non_blocking=True)
            ev_in[bsel].record(copy_stream)

    def e2e_run(ksteps):
        cur = torch.cuda.current_stream()
        for e in ev_free:
            e.record(cur)
        prefetch(0)
        for k in range(ksteps):
            if k + 1 < ksteps:
                prefetch(k + 1)
            cur.wait_event(ev_in[k & 1])
            Iv = dbuf[k & 1][0].detach().requires_grad_(True)
            Tv = dbuf[k & 1][1].detach().requires_grad_(True)
            l = step(Iv, Tv)
            ev_free[k & 1].record(cur)
            hloss.copy_(l.detach(), non_blocking=True)

    e2e_run(3)
    barrier()
    ke = max(3, min(args.steps, 10))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(ke)
    e1.record()
    barrier()
    e2e_ev = [(e0, e1)]
    et = torch.tensor([sum(a.elapsed_time(b) for a, b in e2e_ev) / ke], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
    e2e_value = B / (et.item() * 1e-3)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"c3: experiment_3 anchor+lalign+(lunif(img)+lunif(txt))/2 fwd+bwd, global B={B}, D={D}, "
                                   f"tau={TAU}, bf16 unit-norm rows, rows sharded over {world} GPU(s)",
                       "l2": "256 MiB buffer written between timed iterations (untimed); per-step scratch also exceeds the 126 MB L2",
                       "timing": "sum of per-step CUDA-event intervals, max over ranks",
                       "launch": "one CUDA-graph replay per step" if graphed is not None else "eager launches"},
            "loss": loss_val,
            "step_ms_rank0": [round(x, 3) for x in step_ms],
            "python_gc_collections_in_timed_region": gc_in_region,      # per generation, rank 0
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * n * D * 2, "d2h_bytes_per_step": 4,
                    "ms_per_step": et.item(),
                    "how": "public API on device buffers filled from pinned host memory; the copy of step k+1 overlaps "
                           "the compute of step k (double buffer, copy stream); one CUDA-event region over all steps"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": measured_traffic(), "peak_source": peak_src,
                         "kernel": "the five B x B sweep launches of a step: k_tc_pass<LSE2> (row + column LSE in one sweep), "
                                   "2 x k_tc_pair<anchor-grad>, 2 x k_tc_pair<lunif> (dominant: k_tc_pair<lunif>; `traffic` = "
                                   "its DRAM bytes per launch)",
                         "dominant_kernel": {"name": "k_tc_pair<M_LUNIF_GRAD>", "algorithmic_flops_per_launch": 4.0 * B * B * D / world,
                                             "ms_per_launch": per.get("lunif", [0.0]) and sum(per["lunif"]) / len(per["lunif"]),
                                             "achieved_tflops": (4.0 * B * B * D / world) / max(1e-9, sum(per["lunif"]) / len(per["lunif"]) * 1e-3) / 1e12
                                             if per.get("lunif") else None},
                         "algorithmic_flops_per_step": flops_alg, "passes_ms_per_step": pass_ms,
                         "per_pass_ms": {k: sum(v) / prof_steps for k, v in per.items()},
                         "whole_step_frac": flops_alg / world / (ms_per_step * 1e-3) / 1e12 / peak},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_record(args)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
