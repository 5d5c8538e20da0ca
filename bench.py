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
D2H read of the loss inside the timed region); `roofline.frac` = algorithmic FLOPs of the step /
(ms_per_step x measured bf16 peak), with the sweep-only and dominant-kernel figures under their own
keys; `cpu_baseline` = the torch port of the reference op sequence (oracle/torch_port.py) on the host
cores, on a bounded B = 8192 sample (measured sample time printed, `value` extrapolated and marked);
`eager_gpu_baseline` = the same port run as eager PyTorch on THIS GPU at B = 4096 (the largest size torch.pdist's
backward fits), with this library on the same sample next to it (SURVEY.md §8d's second comparison);
`parity` (N > 1, untimed) = sharded result vs the same kernels unsharded and vs the fp64 closed form.
`--config c2|c3|c4` selects another BASELINE.json workload.  `--impl reference` times the CPU path alone.
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
TAU = 0.1
# BASELINE.json configs that are loss-only workloads.  `contractions` = algorithmic B x B x D
# contractions per step (anchor 3, each L_unif call 2): FLOPs = 2 B^2 D each.
CONFIGS = {
    "c2": dict(batch=4096, dim=512, learn_tau=False, contractions=5,
               name="c2: experiment_4 anchor+lalign+lunif(centroids)",
               w=dict(anchor=1.0, align=1.0, unif_img=0.0, unif_txt=0.0, unif_cen=1.0)),
    "c3": dict(batch=32768, dim=512, learn_tau=False, contractions=7,
               name="c3: experiment_3 anchor+lalign+(lunif(img)+lunif(txt))/2",
               w=dict(anchor=1.0, align=1.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0)),
    # experiment_10 at step 0.6 t_total: alpha = get_alpha(..)=1.2, beta = get_beta(..)=0.2 (sparsify_clip.py:879-902)
    "c4": dict(batch=65536, dim=768, learn_tau=True, contractions=5,
               name="c4: experiment_10 anchor+1.2*lalign+0.2*lunif(centroids), learnable tau",
               w=dict(anchor=1.0, align=1.2, unif_img=0.0, unif_txt=0.0, unif_cen=0.2)),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS), help="BASELINE.json workload (default: the headline c3)")
    ap.add_argument("--batch", type=int, default=None, help="override the global batch B of the config")
    ap.add_argument("--dim", type=int, default=None, help="override D of the config")
    ap.add_argument("--cpu-sample-batch", type=int, default=8192, help="B of the bounded CPU sample (BASELINE.md §4)")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed sharded-parity block at N > 1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tc-flags", type=int, default=None, help="debug: scb_set_tc_flags value")
    ap.add_argument("--launch", default="auto", choices=["auto", "graph", "eager"],
                    help="how a step is launched: 'graph' = the whole step (library kernels and, at N > 1, the NCCL collectives) "
                         "captured once into a CUDA graph and replayed; 'eager' = ~50 separate launches; 'auto' (default) = "
                         "graph when the capture succeeds (it cannot with a CPU-resident learnable temperature: c4), else eager")
    ap.add_argument("--graph", action="store_true", help="same as --launch graph")
    a = ap.parse_args()
    if a.graph:
        a.launch = "graph"
    cfg = CONFIGS[a.config]
    a.batch = a.batch or cfg["batch"]
    a.dim = a.dim or cfg["dim"]
    a.cfg = cfg
    return a


TRAFFIC_FILES = ("r02_traffic.json", "r01_traffic.json")


def measured_traffic(config="c3", world=1):
    """(DRAM bytes per launch of the dominant kernel, provenance) from the newest committed ncu --set full capture under
    profiles/ -- a number read from a file, NOT measured in this run -- or (None, why).  The captures are of fixed shapes:
    c3's sweeps on one GPU, and c4's 8-GPU shard (8192 rows x 65536 columns, D = 768); other runs carry null."""
    if config == "c4":
        if world != 8:
            return None, "no ncu capture of this shape (profiles/r03_traffic_c4_shard.json is c4's 8-GPU shard)"
        files = ("r03_traffic_c4_shard.json",)
    elif config == "c3" and world == 1:
        files = TRAFFIC_FILES
    else:
        return None, "no ncu capture of this shape (the committed captures are c3 on one GPU and c4's 8-GPU shard)"
    for name in files:
        p = os.path.join(ROOT, "profiles", name)
        try:
            with open(p) as f:
                d = json.load(f)
            k = d["per_launch_dram_bytes"][d["dominant"]]
            return float(k["read"] + k["write"]), f"profiles/{name} (ncu --set full of {d['dominant']}, {d.get('date', 'round 1')}; not measured in this run)"
        except Exception:
            continue
    return None, "no ncu capture committed"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
    return 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"


def sustained_peak():
    """MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS back to back for 4 s), or None."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops_sustained"])
    except Exception:
        return None


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
def _weights_tuple(w):
    return (w["anchor"], w["align"], w["unif_img"], w["unif_txt"], w["unif_cen"])


def cpu_reference_step_time(batch, dim, steps, warmup, cfg):
    """torch port of the reference op sequence, fp32, all host threads; returns (median s/step, threads)."""
    import torch
    from oracle import torch_port
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(42)
    I = torch.nn.functional.normalize(torch.randn(batch, dim, generator=g), dim=-1)
    T = torch.nn.functional.normalize(I + 0.5 * torch.randn(batch, dim, generator=g), dim=-1)
    tau = torch.tensor(TAU) if cfg["learn_tau"] else TAU
    w = _weights_tuple(cfg["w"])
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        torch_port.fwd_bwd(I, T, tau, w)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return statistics.median(times), torch.get_num_threads()


def _cpu_sample(args, steps, warmup):
    """One bounded CPU sample: the config's composition at B = min(cpu-sample-batch, B) (B = 8192 is the
    largest size that is always run; B = 32768 needs > 128 GB of host RAM for autograd's saved B x B tensors)."""
    bs = min(args.cpu_sample_batch, args.batch)
    t, threads = cpu_reference_step_time(bs, args.dim, steps, warmup, args.cfg)
    extrapolated = bs != args.batch
    t_full = t * (args.batch / bs) ** 2          # every term but L_align is O(B^2 D)
    return dict(bs=bs, t=t, threads=threads, extrapolated=extrapolated, value=args.batch / t_full, sample_value=bs / t)


def cpu_baseline_record(args, steps=2, warmup=1):
    r = _cpu_sample(args, steps, warmup)
    return {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port", "extrapolated": r["extrapolated"],
            "measured_sample": {"batch": r["bs"], "s_per_step": r["t"], "pairs_per_s": r["sample_value"]},
            "sample": f"oracle/torch_port.py {args.cfg['name']} fwd+bwd fp32 at B={r['bs']}, D={args.dim}: median {r['t']:.3f} s/step "
                      f"over {steps} steps = {r['sample_value']:.1f} pairs/s MEASURED"
                      + (f"; value extrapolated to B={args.batch} by (B/{r['bs']})^2 (pairs/s falls as 1/B)" if r["extrapolated"] else "")}


def eager_gpu_baseline_record(args, scb, torch, dev, steps=5, warmup=2):
    """Part of the baseline leg ("also time the reference's eager GPU path on one B200 where it fits"): the
    torch port of the reference op sequence run on THIS GPU, fp32 without autocast (the parity target of SURVEY.md §A.3),
    on a bounded sample -- torch.pdist's backward keeps B(B-1)/2 x D fp32 of scratch (17 GB at B = 4096, D = 512; 68 GB at
    B = 8192), so B = 4096 is the largest size that always fits.  Next to it: this library on the same sample (bf16,
    eager launches), timed the same way.  Reported, never the target; untimed with respect to the bench line."""
    from oracle import torch_port
    bs = min(4096, args.batch)
    w = _weights_tuple(args.cfg["w"])
    g = torch.Generator(device=dev).manual_seed(42)
    I = torch.nn.functional.normalize(torch.randn(bs, args.dim, generator=g, device=dev), dim=-1)
    T = torch.nn.functional.normalize(I + 0.5 * torch.randn(bs, args.dim, generator=g, device=dev), dim=-1)

    def timed(fn):
        ts = []
        for i in range(warmup + steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if i >= warmup:
                ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    try:
        tau = torch.tensor(TAU, device=dev) if args.cfg["learn_tau"] else TAU
        torch.cuda.reset_peak_memory_stats(dev)
        base = torch.cuda.memory_allocated(dev)
        ms_ref = timed(lambda: torch_port.fwd_bwd(I, T, tau, w))
        ref_mem = torch.cuda.max_memory_allocated(dev) - base
        Ib, Tb = I.to(torch.bfloat16).requires_grad_(True), T.to(torch.bfloat16).requires_grad_(True)
        tp = torch.nn.Parameter(torch.tensor(TAU)) if args.cfg["learn_tau"] else TAU

        def ours():
            Ib.grad = Tb.grad = None
            scb.weighted_loss(Ib, Tb, tp, args.cfg["w"]).backward()
        ms_ours = timed(ours)
    except Exception as exc:
        torch.cuda.empty_cache()
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
    torch.cuda.empty_cache()
    return {"value": bs / (ms_ref * 1e-3), "unit": UNIT, "kind": "port", "device": torch.cuda.get_device_name(dev),
            "ms_per_step": ms_ref, "peak_scratch_bytes": int(ref_mem),
            "ours_same_sample": {"value": bs / (ms_ours * 1e-3), "ms_per_step": ms_ours, "launch": "eager launches, bf16"},
            "sample": f"oracle/torch_port.py {args.cfg['name']} fwd+bwd, fp32 eager PyTorch on the GPU (mm + cross_entropy, "
                      f"pdist and its backward, as the reference runs them), B={bs}, D={args.dim}: median of {steps} steps after "
                      f"{warmup} warm-ups, CUDA events; larger B does not fit torch.pdist's backward scratch"}


def run_reference(args):
    """--impl reference: the reference's CPU path (torch port of its op sequence: the reference is a script whose imports
    are not installable here) on a bounded sample; `ms_per_step` is the MEASURED time of one sample step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    r = _cpu_sample(args, steps, warmup)
    sample = (f"each step = oracle/torch_port.py {args.cfg['name']} fwd+bwd fp32 on a B={r['bs']} sample: median "
              f"{r['t']:.3f} s MEASURED = {r['sample_value']:.1f} pairs/s at B={r['bs']}"
              + (f"; `value` is that sample extrapolated to B={args.batch} by (B/{r['bs']})^2" if r["extrapolated"] else ""))
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": r["t"] * 1e3, "ms_per_step_is": f"measured, one B={r['bs']} sample step",
            "extrapolated": r["extrapolated"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.cfg['name']}, B={args.batch}, D={args.dim}, tau={TAU}, CPU torch port of the "
                                   f"reference ops on a bounded B={r['bs']} sample"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": sample,
                             "extrapolated": r["extrapolated"],
                             "measured_sample": {"batch": r["bs"], "s_per_step": r["t"], "pairs_per_s": r["sample_value"]}},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- sharded parity (untimed, N > 1)
def sharded_parity(scb, dist, torch, dev, rank, world):
    """Row-sharded result over `world` ranks against (a) the same kernels unsharded on rank 0 at B = 8192 and (b) the fp64
    closed form of the reference (oracle/closed_form.py, the checker) at B = 1024.  Untimed; printed as `parity`."""
    out = {}
    w = dict(anchor=1.0, align=1.0, unif_img=0.5, unif_txt=0.5, unif_cen=0.0)
    wc = dict(anchor=1.0, align=1.2, unif_img=0.0, unif_txt=0.0, unif_cen=0.2)
    for key, (B, D, tau, ww, oracle) in {"exp3_B8192_vs_unsharded": (8192, 512, 0.1, w, False),
                                         "exp3_B1024_vs_fp64_closed_form": (1024, 512, 0.1, w, True),
                                         "exp10_B2048_D768_vs_fp64_closed_form": (2048, 768, 0.1, wc, True)}.items():
        g = torch.Generator(device=dev).manual_seed(1234)          # the same full batch on every rank
        I = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device=dev), dim=-1)
        T = torch.nn.functional.normalize(I + 0.5 * torch.randn(B, D, generator=g, device=dev), dim=-1)
        I, T = I.to(torch.bfloat16).float(), T.to(torch.bfloat16).float()
        n = B // world
        prev = scb.set_fp32_mode("bf16")
        try:
            Il = I[rank * n:(rank + 1) * n].clone().requires_grad_(True)
            Tl = T[rank * n:(rank + 1) * n].clone().requires_grad_(True)
            tp = torch.nn.Parameter(torch.tensor(tau))
            loss = scb.weighted_loss(Il, Tl, tp, ww, group=True)
            loss.backward()
            gI, gT = torch.empty(B, D, device=dev), torch.empty(B, D, device=dev)
            dist.all_gather_into_tensor(gI, Il.grad.contiguous())
            dist.all_gather_into_tensor(gT, Tl.grad.contiguous())
            if rank == 0:
                rec = {"B": B, "D": D, "world": world, "loss": loss.item()}
                if oracle:
                    from oracle import closed_form as cf      # the checker, never the thing measured
                    ref, dI, dT, dtau, _ = cf.weighted_loss(I.cpu().numpy(), T.cpu().numpy(), tau, *_weights_tuple(ww))
                    import numpy as np
                    rec.update(loss_rel=abs(loss.item() - ref) / abs(ref),
                               dI_rel=float(np.linalg.norm(gI.double().cpu().numpy() - dI) / np.linalg.norm(dI)),
                               dT_rel=float(np.linalg.norm(gT.double().cpu().numpy() - dT) / np.linalg.norm(dT)),
                               dtau_rel=abs(tp.grad.item() - dtau) / abs(dtau))
                    rec["ok"] = rec["loss_rel"] <= 1e-5 and max(rec["dI_rel"], rec["dT_rel"], rec["dtau_rel"]) <= 1e-3
                else:
                    If, Tf = I.clone().requires_grad_(True), T.clone().requires_grad_(True)
                    tf_ = torch.nn.Parameter(torch.tensor(tau))
                    full = scb.weighted_loss(If, Tf, tf_, ww)
                    full.backward()
                    rec.update(loss_rel=abs(loss.item() - full.item()) / abs(full.item()),
                               dI_rel=((gI - If.grad).norm() / If.grad.norm()).item(),
                               dT_rel=((gT - Tf.grad).norm() / Tf.grad.norm()).item(),
                               dtau_rel=abs(tp.grad.item() - tf_.grad.item()) / abs(tf_.grad.item()))
                    rec["ok"] = rec["loss_rel"] <= 2e-6 and max(rec["dI_rel"], rec["dT_rel"]) <= 2e-4 and rec["dtau_rel"] <= 1e-4
                out[key] = rec
        finally:
            scb.set_fp32_mode(prev)
    return out


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
    cfg = args.cfg
    W = dict(cfg["w"], alpha=0.0, beta=0.0)
    assert B % world == 0 and 8 % world == 0, "global batch must divide over 1, 2, 4 or 8 ranks"
    n = B // world
    # The global batch is generated in 8 row chunks with seeds 42 .. 49 (seed 42 + r at 8 ranks), so the
    # SAME global batch is sharded whatever N is: the printed loss must agree between the N = 1, 2, 4, 8 runs.
    chunks = []
    for c in range(rank * (8 // world), (rank + 1) * (8 // world)):
        g = torch.Generator(device=dev).manual_seed(42 + c)
        Ic = torch.nn.functional.normalize(torch.randn(B // 8, D, generator=g, device=dev), dim=-1)
        Tc = torch.nn.functional.normalize(Ic + 0.5 * torch.randn(B // 8, D, generator=g, device=dev), dim=-1)
        chunks.append((Ic, Tc))
    I0 = torch.cat([c[0] for c in chunks])
    T0 = torch.cat([c[1] for c in chunks])
    del chunks
    I = I0.to(torch.bfloat16).requires_grad_(True)
    T = T0.to(torch.bfloat16).requires_grad_(True)
    # learnable temperature as the reference creates it: a 0-dim fp32 nn.Parameter on the CPU (sparsify_clip.py:716-717)
    tau_p = torch.nn.Parameter(torch.tensor(TAU, dtype=torch.float32)) if cfg["learn_tau"] else TAU
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step(Iv, Tv):
        Iv.grad = None
        Tv.grad = None
        if cfg["learn_tau"]:
            tau_p.grad = None
        loss = scb.weighted_loss(Iv, Tv, tau_p, W, group=group)
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

    # The whole step (about 50 launches, most of them tiny, plus the NCCL collectives at N > 1) is captured once into a
    # CUDA graph and replayed: the library is capturable by contract (no allocation, no host sync, caller's stream).
    # A learnable temperature living on the CPU (the reference's placement, sparsify_clip.py:716-717) needs a device-to-host
    # copy of its gradient every step, which a capture cannot contain: that configuration runs eager.
    graphed = None
    # measured (profiles/r02l_*): N = 8 1.27 ms replayed vs 1.33 ms eager (the shard step is ~50 launches for ~1.1 ms of
    # sweeps); N = 2 and N = 1 within noise of each other
    want_graph = args.launch == "graph" or (args.launch == "auto" and not cfg["learn_tau"])
    ok_flag = torch.ones(1, device=dev, dtype=torch.int32)
    if want_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step(I, T)
            torch.cuda.current_stream().wait_stream(side)
            barrier()
            gr = torch.cuda.CUDAGraph()
            I.grad = None
            T.grad = None
            # thread_local: the NCCL watchdog thread's own CUDA calls must not invalidate this thread's capture
            with torch.cuda.graph(gr, capture_error_mode="thread_local"):
                g_loss = scb.weighted_loss(I, T, tau_p, W, group=group)
                g_loss.backward()
            graphed = (gr, g_loss)
            gr.replay()
            torch.cuda.synchronize()
        except Exception as exc:      # fall back to eager timing, and say so
            sys.stderr.write(f"bench.py[rank {rank}]: CUDA-graph capture failed ({type(exc).__name__}: {exc}); timing eager launches\n")
            graphed = None
            ok_flag.zero_()
        if world > 1:                 # every rank replays, or none does
            dist.all_reduce(ok_flag, op=dist.ReduceOp.MIN)
            if ok_flag.item() == 0:
                graphed = None
        barrier()

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
    # Eager launches only: the host is kept at most ONE step ahead of the device.  Running further ahead changes which
    # blocks of the caching allocator are still pinned by the side streams (the gathers, the row split) when the next
    # step asks for its scratch, and the allocator then goes to cudaMalloc inside the timed region (seen at c4 on one
    # GPU, 50 ms steps: 15 device allocations and three steps of 86 / 190 / 284 ms).  The device never idles for it.
    def pace(done_events):
        if graphed is None and len(done_events) >= 2:
            done_events[-2].synchronize()

    gc.collect()
    gc.freeze()
    burst = []
    for _ in range(args.steps):
        flush.zero_()
        pace(burst)
        step(I, T)
        ev = torch.cuda.Event()
        ev.record()
        burst.append(ev)
    barrier()

    # ---- timed region: K steps, each bracketed by CUDA events, L2 flushed (untimed) in between
    first_row = len(sampler.rows)      # only samples taken from here on (the timed region) are reported
    gc0 = [g["collections"] for g in gc.get_stats()]
    mallocs0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
    evs = []
    launches0 = be.launches
    barrier()
    for _ in range(args.steps):
        flush.zero_()
        pace([e for _, e in evs])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = timed_step()
        e1.record()
        evs.append((e0, e1))
    barrier()
    launches = be.launches - launches0
    gc_in_region = [g["collections"] - a for g, a in zip(gc.get_stats(), gc0)]
    mallocs_in_region = torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - mallocs0
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
    flops_alg = 2.0 * cfg["contractions"] * B * B * D   # anchor 3 contractions, each lunif 2 (2 B^2 D each)
    peak, peak_src = peaks()
    peak_sust = sustained_peak()
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
        with torch.cuda.stream(copy_stream), torch.no_grad():
            copy_stream.wait_event(ev_free[bsel])          # the step that last read this buffer has finished
            dbuf[bsel][0].copy_(hI, non_blocking=True)
            dbuf[bsel][1].copy_(hT, non_blocking=True)
            ev_in[bsel].record(copy_stream)

    # with graph launches: one captured step per device buffer pair (the graph reads fixed addresses)
    e2e_graphs = None
    if graphed is not None:
        try:
            e2e_graphs = []
            for bsel in range(2):
                Iv = dbuf[bsel][0].requires_grad_(True)
                Tv = dbuf[bsel][1].requires_grad_(True)
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    step(Iv, Tv)
                torch.cuda.current_stream().wait_stream(side)
                barrier()
                Iv.grad = None
                Tv.grad = None
                gre = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gre, capture_error_mode="thread_local"):
                    le = scb.weighted_loss(Iv, Tv, tau_p, W, group=group)
                    le.backward()
                e2e_graphs.append((gre, le))
        except Exception as exc:
            sys.stderr.write(f"bench.py[rank {rank}]: e2e graph capture failed ({type(exc).__name__}: {exc}); e2e runs eager\n")
            e2e_graphs = None
            ok_flag.zero_()
        if world > 1:
            dist.all_reduce(ok_flag, op=dist.ReduceOp.MIN)
            if ok_flag.item() == 0:
                e2e_graphs = None
        barrier()

    def e2e_run(ksteps):
        cur = torch.cuda.current_stream()
        for e in ev_free:
            e.record(cur)
        prefetch(0)
        for k in range(ksteps):
            if k + 1 < ksteps:
                prefetch(k + 1)
            cur.wait_event(ev_in[k & 1])
            if e2e_graphs is not None:
                e2e_graphs[k & 1][0].replay()
                l = e2e_graphs[k & 1][1]
            else:
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

    parity = None
    if world > 1 and not args.no_parity:
        try:
            parity = sharded_parity(scb, dist, torch, dev, rank, world)
        except Exception as exc:           # the parity block must never cost the bench line
            parity = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    peak_mem = torch.cuda.max_memory_allocated(dev)

    if rank == 0:
        whole = flops_alg / world / (ms_per_step * 1e-3) / 1e12        # per-GPU algorithmic TFLOP/s over the WHOLE step
        traffic, traffic_src = measured_traffic(args.config, world)
        dom = max(per, key=lambda k: sum(per[k])) if per else None
        dom_calls = {"lse": 1.0, "anchor_grad": 2.0, "lunif": 2.0}      # algorithmic contractions one launch of each covers
        dom_ms = (sum(per[dom]) / len(per[dom])) if dom else None
        dom_flops = (2.0 * dom_calls.get(dom, 2.0) * B * B * D / world) if dom else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{cfg['name']} fwd+bwd, global B={B}, D={D}, tau={TAU}"
                                   f"{' (learnable, CPU 0-dim parameter)' if cfg['learn_tau'] else ''}, bf16 unit-norm rows, "
                                   f"rows sharded over {world} GPU(s)",
                       "inputs": "the same global batch for every N (8 row chunks, seeds 42..49): `loss` is comparable across N",
                       "l2": "256 MiB buffer written between timed iterations (untimed); per-step scratch also exceeds the 126 MB L2",
                       "timing": "sum of per-step CUDA-event intervals, max over ranks",
                       "launch": "one CUDA-graph replay per step" if graphed is not None else "eager launches"},
            "loss": loss_val,
            "step_ms_rank0": [round(x, 3) for x in step_ms],
            "python_gc_collections_in_timed_region": gc_in_region,      # per generation, rank 0
            "cuda_mallocs_in_timed_region": mallocs_in_region,          # caching-allocator misses (each one syncs the device)
            "peak_device_memory_bytes_rank0": int(peak_mem),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * n * D * 2, "d2h_bytes_per_step": 4,
                    "ms_per_step": et.item(),
                    "launch": "one CUDA-graph replay per step" if e2e_graphs is not None else "eager launches",
                    "how": "public API on device buffers filled from pinned host memory; the copy of step k+1 overlaps "
                           "the compute of step k (double buffer, copy stream); one CUDA-event region over all steps. "
                           "Only the 4-byte loss is read back: the gradients dI, dT stay on the device, where the training "
                           "loop's encoder backward consumes them (sparsify_clip.py:960-966)"},
            "gpu_launches": launches,
            "clocks": clocks,
            # `frac` is the WHOLE-STEP fraction: algorithmic FLOPs of the step / (ms_per_step x peak), i.e. it follows from
            # `ms_per_step` above.  The sweep-only and dominant-kernel figures explain it and sit under their own keys.
            "roofline": {"bound": "tensor", "achieved": whole, "peak": peak, "unit": "TFLOP/s", "frac": whole / peak,
                         "frac_vs_sustained_peak": (whole / peak_sust) if peak_sust else None,      # back-to-back cuBLAS, 4 s
                         "frac_vs_datasheet_2250": whole / 2250.0,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_flops_per_step": flops_alg,
                         "sweeps_only": {"ms_per_step": pass_ms, "achieved": achieved, "frac": achieved / peak,
                                         "note": "summed CUDA-event time of the B x B sweep launches, measured in 3 separate "
                                                 "instrumented steps after the timed region",
                                         "per_pass_ms": {k: sum(v) / prof_steps for k, v in per.items()}},
                         "dominant_kernel": {"pass": dom, "algorithmic_flops_per_launch": dom_flops, "ms_per_launch": dom_ms,
                                             "achieved": (dom_flops / (dom_ms * 1e-3) / 1e12) if dom else None,
                                             "frac": (dom_flops / (dom_ms * 1e-3) / 1e12 / peak) if dom else None,
                                             "note": "`traffic` = DRAM bytes of one launch of this kernel"}},
        }
        if parity is not None:
            line["parity"] = parity
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_record(args)
            line["eager_gpu_baseline"] = eager_gpu_baseline_record(args, scb, torch, dev)
        print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        # Leave without tearing the communicator down: destroying a process group whose collectives were captured into
        # CUDA graphs was measured to hang (the line above is already out), and nothing is left to clean up.
        graphed = e2e_graphs = None
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
