#!/usr/bin/env python
"""bench.py -- RegT-GCN hot path: fwd + loss + bwd samples/s on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path (oracle port)

One "step" = one pass of the hot path (forward + MSE loss + backward, all parameter gradients)
over one batch of B synthetic snapshots.  Workload at N=1: BASELINE.json configs[1]
(A3TGCN/TemporalGCN, METR-LA shape: 207 nodes, 12 periods, batch 64, hidden 64).  For N>1:
  * workloads without regions (config 2): every rank runs the same per-GPU batch on its own
    synthetic snapshots (weak scaling over the batch dimension -- snapshots are independent units);
  * regional workloads (--workload 3|4|5): each GPU owns a set of regions (LPT bin packing), reads
    its rows + 1-hop halo rows of x, and the whole job processes the config's B snapshots (strong
    scaling);
in both cases the shared-weight gradients (+ the loss) are all-reduced over NCCL inside the step.

Timing: CUDA events on the launching stream around every step, L2 flushed (256 MiB memset)
between timed steps, barrier + synchronize on both sides, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "regt-gcn_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

_REAL_STDOUT = None
METRIC = "regt_gcn_fwd_bwd_samples_per_s"
UNIT = "samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="2", help="SURVEY 8(d) config id (default 2 = BASELINE configs[1])")
    ap.add_argument("--batch", type=int, default=None, help="override the per-GPU batch B")
    ap.add_argument("--precision", default="bf16", choices=["fp32", "tf32x3", "bf16"],
                    help="bf16 = tcgen05 path (bf16 operands, fp32 accumulate; stated tolerance, tests/test_gpu_tc.py); "
                         "fp32 = FFMA path with 1e-5 parity; the default run reports both (fp32 under 'fp32_parity_mode')")
    ap.add_argument("--micro-batch", type=int, default=None)
    ap.add_argument("--shard", default="auto", choices=["auto", "region", "batch"],
                    help="N>1: 'region' = each GPU owns a set of regions (strong scaling, SURVEY 8(e)); "
                         "'batch' = each GPU runs its own per-GPU batch (weak scaling); auto = region when the workload has regions")
    ap.add_argument("--no-graph", action="store_true", help="do not capture the step in a CUDA graph")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: the gradient all-reduce as one kernel over NVLink peer memory (csrc/peer.cu) or through NCCL; "
                         "auto = the peer kernels (NCCL only if CUDA IPC is unavailable)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_oracle_rate(w, seconds: float, min_samples: int = 4, warm: int = 1):
    """times the reference's CPU path (the pure-torch fp32 oracle port: per-call re-normalisation,
    per-sample and per-period Python loops) on this box's host cores."""
    from oracle import regt_oracle as O
    from regt_b200 import workloads as W
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    if w.model == "TemporalGCN":
        m = O.TemporalGCN(8, w.T, w.O, hidden=w.H)
    else:
        m = O.RegionalTemporalGCN(8, w.N, w.T, w.O, hidden=w.H, n_regions=w.R)
    W.init_params_synthetic(m, 1234)
    x, y = w.inputs(max(min_samples, 8))
    g = w.graph_args()

    def one(b):
        out, _ = m(x[b % x.shape[0]], *g)
        loss = torch.mean((out - y[b % x.shape[0]]) ** 2)
        loss.backward()

    for b in range(warm):
        one(b)
    n, t0 = 0, time.perf_counter()
    while True:
        one(n); n += 1
        el = time.perf_counter() - t0
        if (el >= seconds and n >= min_samples) or n >= 4096:
            break
    return n / el, n, el, torch.get_num_threads()


def run_reference(args):
    """--impl reference: the reference's own implementation cannot be imported (torch_geometric is
    absent and not installable offline), so this arm times the oracle port of it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from regt_b200 import workloads as W
    w = W.make_workload(args.workload, args.batch)
    per_step = 4  # snapshots per "step": a bounded sample of the workload's batch
    from oracle import regt_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    m = (O.TemporalGCN(8, w.T, w.O, hidden=w.H) if w.model == "TemporalGCN"
         else O.RegionalTemporalGCN(8, w.N, w.T, w.O, hidden=w.H, n_regions=w.R))
    W.init_params_synthetic(m, 1234)
    x, y = w.inputs(per_step)
    g = w.graph_args()

    def step():
        for b in range(per_step):
            out, _ = m(x[b], *g)
            torch.mean((out - y[b]) ** 2).backward()

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    ts = []
    budget_t0 = time.perf_counter()
    for _ in range(args.steps):
        t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
        if time.perf_counter() - budget_t0 > 150:
            break
    ms = 1e3 * sum(ts) / len(ts)
    val = per_step / (ms / 1e3)
    sample = f"{per_step} of {w.B} snapshots per step, {len(ts)} steps, oracle fp32 port on host cores"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": len(ts),
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": dict(w.describe(), samples_per_step=per_step),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------
def kernel_alg_bytes(w, B: int) -> dict:
    """algorithmic bytes per launch of the two dominant kernels under SURVEY 8(d)'s save-4-planes
    accounting: per (b,n,t) row the forward cell reads X_t,S_t (2*F*4 B) and
    writes 4 planes (4*H*4 B); the backward cell reads the 4 planes; weights once per launch."""
    rows = B * w.N * w.T
    F, H = 8, w.H
    wts = (F + H) * 3 * H * 4
    fwd = rows * (2 * F * 4 + 4 * H * 4) + B * w.N * H * 4 + wts
    bwd = rows * (4 * H * 4) + B * w.N * H * 4 + wts
    return {"k_cell_fwd": fwd, "k_cell_bwd": bwd, "k_cell_fwd_tc": fwd, "k_cell_bwd_tc": bwd}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    # stdout carries exactly ONE JSON line: native libraries (NCCL prints its version banner to stdout) are sent
    # to stderr by pointing fd 1 there; the line is written to the saved descriptor.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    from models import RegionalTemporalGCN, TemporalGCN
    from regt_b200 import _lib, workloads as W
    lib = _lib.load()

    w = W.make_workload(args.workload, args.batch)
    B = w.B
    if w.model == "TemporalGCN":
        model = TemporalGCN(8, w.T, w.O, hidden=w.H, precision=args.precision)
    else:
        model = RegionalTemporalGCN(8, w.N, w.T, w.O, hidden=w.H, n_regions=w.R, precision=args.precision)
    W.init_params_synthetic(model, 1234)
    model = model.to(dev)
    graph_args = tuple(None if a is None else a.to(dev) for a in w.graph_args())

    from regt_b200 import shard as S
    os.environ["REGT_EXCHANGE"] = args.exchange
    sharded = world > 1 and w.R > 0 and args.shard != "batch"
    if args.shard == "region" and world > 1 and w.R == 0:
        raise SystemExit("--shard region needs a regional workload (3, 4 or 5)")

    # host (pinned) and device-resident inputs
    sm = None
    if sharded:
        # region shards: every rank sees every snapshot, owns the nodes of its regions (+ halo rows)
        sm = S.RegionShardedModel(model, graph_args[0], list(graph_args[1:1 + w.R]), list(graph_args[1 + w.R:]), rank, world)
        ex = sm.exchange
        xh, yh = w.inputs(B)
        xh = xh.index_select(1, torch.from_numpy(sm.shard.perm)).contiguous()   # the rank's loader reads its rows only
        yh = yh.index_select(1, torch.from_numpy(sm.shard.own)).contiguous()
    else:
        # one flat gradient buffer so that a single NCCL all-reduce covers every shared weight
        ex = S.GradExchange([p for n, p in model.named_parameters() if p.requires_grad], world)
        xh, yh = w.inputs(B, seed_offset=rank)     # batch shards: every rank gets its own snapshots
    xh, yh = xh.pin_memory(), yh.pin_memory()
    # two device input buffers: the end-to-end loop copies step k+1's inputs while step k computes
    bufs = [(xh.to(dev), yh.to(dev)), (xh.to(dev), yh.to(dev))]
    loss_h = [torch.zeros(1).pin_memory(), torch.zeros(1).pin_memory()]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def raw_step(b=0):
        xd, yd = bufs[b]
        if world > 1:
            ex.zero()     # per-step exchange: the buffer holds this step's contribution only
        if sharded:
            return sm.fused_step(None, None, micro_batch=args.micro_batch, local_inputs=(xd, yd), sync=False)[0]
        return model.fused_step(xd, yd, *graph_args, micro_batch=args.micro_batch)[0]

    # warm-up (eager) -- also builds and caches the static-graph plan (K1), excluded from timing
    lib.regt_launch_count(1)
    loss_d = None
    for _ in range(max(3, args.warmup)):
        loss_d = raw_step()
    torch.cuda.synchronize()
    launches_per_step = lib.regt_launch_count(1) // max(3, args.warmup)

    def exchange(loss):
        """the exchange step: shared-weight gradients + loss, ONE all-reduce of the flat buffer (peer-memory kernel or NCCL)"""
        if not sharded:
            ex.add_loss(loss)
        return ex.sync()

    def full_step(b=0):
        loss = raw_step(b)
        return exchange(loss) if dist is not None else loss

    if dist is not None:          # NCCL communicator warm-up before any capture
        for _ in range(2):
            loss_d = full_step()
        torch.cuda.synchronize()

    graphs, graph_loss, graph_has_exchange = None, [None, None], False
    if not args.no_graph:
        def capture(fn):
            gs, ls = [torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()], [None, None]   # one per input buffer
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                fn()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            for b in (0, 1):
                with torch.cuda.graph(gs[b]):
                    ls[b] = fn(b)
                torch.cuda.synchronize()
            return gs, ls
        if dist is not None:
            # the all-reduce is captured INSIDE the step's graph (one replay = compute + exchange); if this NCCL
            # build refuses capture on any rank, every rank falls back to replay + eager all-reduce
            ok = torch.ones(1, device=dev)
            try:
                graphs, graph_loss = capture(full_step)
            except Exception as e:  # noqa: BLE001
                ok.zero_()
                print(f"[rank {rank}] NCCL graph capture failed ({type(e).__name__}): eager exchange", file=sys.stderr)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            graph_has_exchange = bool(ok.item())
            if not graph_has_exchange:
                graphs, graph_loss = capture(raw_step)
        else:
            graphs, graph_loss = capture(raw_step)
    graph = graphs[0] if graphs else None

    def step(b=0):
        nonlocal loss_d
        if graphs is not None:
            graphs[b].replay()
            loss_d = graph_loss[b]
            if dist is not None and not graph_has_exchange:
                loss_d = exchange(loss_d)
        else:
            loss_d = full_step(b)

    for _ in range(3):
        step()
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput ("value") ----------------
    K = args.steps
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    align = torch.zeros(1, device=dev)
    for a, b in evs:
        flush.zero_()                      # evict L2 between timed iterations (outside the event pair)
        if dist is not None:
            dist.all_reduce(align)         # ... and re-align the ranks: the flush's skew is not part of the step
        a.record(); step(); b.record()
    barrier()
    per = [a.elapsed_time(b) for a, b in evs]
    tot_ms = torch.tensor([sum(per)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tot_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(tot_ms) / K
    clocks = sampler.stop() if sampler else None

    # ---------------- end to end: host buffers in, loss out, every step ----------------
    # Every step's x,y are copied from pinned host memory and every step's loss is read on the host, all
    # inside the timed region.  The copies run on a second stream into the buffer the running step does
    # not use, and a step's loss is read while the next step computes (the reference reads it
    # synchronously, run.py:180; the values are the same, only one step late).
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main, cs = torch.cuda.current_stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event(), torch.cuda.Event()]
    ev_done = [torch.cuda.Event(), torch.cuda.Event()]
    ev_loss = [torch.cuda.Event(), torch.cuda.Event()]

    def h2d(b, wait_free):
        with torch.cuda.stream(cs):
            if wait_free:
                cs.wait_event(ev_done[b])       # the step that last used this buffer has finished
            bufs[b][0].copy_(xh, non_blocking=True)
            bufs[b][1].copy_(yh, non_blocking=True)
            ev_in[b].record(cs)

    barrier()
    cs.wait_stream(main)
    e0.record()
    h2d(0, False)
    losses = []
    for k in range(K):
        b = k & 1
        if k + 1 < K:
            h2d(1 - b, k >= 1)
        main.wait_event(ev_in[b])
        step(b)
        ev_done[b].record(main)
        loss_h[b].copy_(loss_d, non_blocking=True)
        ev_loss[b].record(main)
        if k >= 1:
            ev_loss[1 - b].synchronize()
            losses.append(float(loss_h[1 - b]))
    ev_loss[(K - 1) & 1].synchronize()
    losses.append(float(loss_h[(K - 1) & 1]))
    e1.record()
    barrier()
    assert len(losses) == K and all(v == v for v in losses), "end-to-end loop lost a loss value"
    e2e_ms = torch.tensor([e0.elapsed_time(e1) / K], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms)

    # ---------------- per-kernel breakdown for the roofline (rank 0, eager, events per launch) -----
    roofline, breakdown = None, None
    if rank == 0:
        st_ptr = torch.cuda.current_stream().cuda_stream
        nprof = 5
        agg = {}
        for i in range(nprof):
            flush.zero_()
            torch.cuda.synchronize()
            lib.regt_profile(1, st_ptr)
            raw_step()
            torch.cuda.synchronize()
            for name, ms in _lib.profile_read():
                agg.setdefault(name, []).append(ms)
            lib.regt_profile(0, None)
        # launches of one kernel name within a step are summed per step
        per_kernel = {k: sum(v) / nprof for k, v in agg.items()}
        counts = {k: len(v) // nprof for k, v in agg.items()}
        step_sum = sum(per_kernel.values())
        breakdown = {k: {"ms_per_step": round(v, 5), "launches": counts[k], "share": round(v / step_sum, 4)}
                     for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1])}
        pk = peaks()
        ab = kernel_alg_bytes(w, min(B, args.micro_batch or getattr(sm.model if sharded else model, '_mb', None) or B))
        dom = max((k for k in per_kernel if k in ab), key=lambda k: per_kernel[k], default=None)
        mb_used = args.micro_batch or getattr(sm.model if sharded else model, "_mb", None) or B
        mb_used = min(B, mb_used)
        if args.precision == "tf32x3" and "k_gemm_nt_tma_ts" in per_kernel or "k_gemm_nt_tma" in per_kernel:
            # generic 3xTF32 path: the H x H gate contractions (TMA-fed tcgen05 GEMMs, csrc/gemm_tma.cu) dominate.
            # Algorithmic work of the six NT GEMMs per step (z, r, c forward; dHR, dhg (K = 2H) backward):
            #   flops = 2*rows*H*H * (2 + 1 + 1 + 2);   bytes = rows * 4 * (2H + 2H + 2H + 2H + 3H)   (A read once, C written once)
            # The hardware executes 3 tf32 products per contraction at half the bf16 rate -> fp32-equivalent peak = bf16 / 6.
            # The binding roofline is whichever of the two times is longer (H = 128: HBM; H = 256: tensor pipe).
            gk = "k_gemm_nt_tma_ts" if "k_gemm_nt_tma_ts" in per_kernel else "k_gemm_nt_tma"   # A through TMEM (default) / shared memory
            nmb = (B + mb_used - 1) // mb_used
            rows_ = float(B) * w.N * w.T
            flops = 2.0 * rows_ * w.H * w.H * 6
            nbytes = rows_ * 4.0 * 11 * w.H
            t_all = per_kernel[gk] * 1e-3
            peak_tc = pk["bf16_tflops"] / 6.0
            t_tc, t_hbm = flops / (peak_tc * 1e12), nbytes / (pk["hbm_gbs"] * 1e9)
            if t_hbm >= t_tc:
                ach, peak, unit, bound = nbytes / t_all / 1e9, pk["hbm_gbs"], "GB/s", "hbm"
            else:
                ach, peak, unit, bound = flops / t_all / 1e12, peak_tc, "TFLOP/s", "tensor"
            roofline = {"kernel": gk, "bound": bound, "achieved": ach, "peak": peak, "unit": unit,
                        "frac": ach / peak, "traffic": None,
                        "peak_source": pk["source"] + ("" if bound == "hbm" else ": bf16 dense / 6 (tf32 = half the bf16 rate, 3 products per contraction)"),
                        "alg_flops_per_step": flops, "alg_bytes_per_step": nbytes, "t_tensor_ms": t_tc * 1e3, "t_hbm_ms": t_hbm * 1e3,
                        "launches_per_step": counts[gk], "micro_batches": nmb,
                        "launch_ms": per_kernel[gk] / counts[gk],
                        "share_of_step": per_kernel[gk] / step_sum}
        elif dom is not None:
            t_launch = per_kernel[dom] / counts[dom] * 1e-3
            ach = ab[dom] / t_launch / 1e9
            # DRAM bytes of that kernel per launch from the committed `ncu --set full` capture of this same
            # command (profiles/ncu_traffic.json, written by tools/ncu_summary.py); null for other workloads
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
            if os.path.exists(tpath) and str(args.workload) == "2" and args.precision == "bf16" and not args.batch and world == 1:
                traffic = json.load(open(tpath)).get(dom, {}).get("bytes_per_launch")
            roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                        "frac": ach / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk["source"],
                        "alg_bytes_per_launch": ab[dom], "launch_ms": t_launch * 1e3,
                        "share_of_step": per_kernel[dom] / step_sum}
        step_bytes = W.alg_bytes_per_step(w, B)
        step_roof = {"alg_bytes_per_step": step_bytes, "achieved_gbs": step_bytes / (ms_per_step * 1e-3) / 1e9,
                     "frac_of_hbm_peak": step_bytes / (ms_per_step * 1e-3) / 1e9 / pk["hbm_gbs"],
                     "fwd_bwd_flops_per_step": 3 * W.fwd_flops_per_step(w, B)}

    # ---------------- the fp32 (1e-5 parity) mode of the same step, same inputs, same run ----------------
    fp32_mode = None
    if rank == 0 and world == 1 and args.precision == "bf16" and str(args.workload) == "2":   # small config only: a second workspace
        p32 = "tf32x3" if w.H % 32 == 0 else "fp32"
        m32 = (TemporalGCN(8, w.T, w.O, hidden=w.H, precision=p32) if w.model == "TemporalGCN"
               else RegionalTemporalGCN(8, w.N, w.T, w.O, hidden=w.H, n_regions=w.R, precision=p32))
        W.init_params_synthetic(m32, 1234)
        m32 = m32.to(dev)
        xd, yd = bufs[0]
        for _ in range(3):
            m32.fused_step(xd, yd, *graph_args, micro_batch=args.micro_batch)
        ts = []
        for _ in range(5):
            flush.zero_()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record(); m32.fused_step(xd, yd, *graph_args, micro_batch=args.micro_batch); b_.record()
            torch.cuda.synchronize()
            ts.append(a_.elapsed_time(b_))
        ms32 = sum(ts) / len(ts)
        fp32_mode = {"value": B / (ms32 * 1e-3), "unit": UNIT, "ms_per_step": ms32, "steps": 5,
                     "precision": p32,
                     "note": "fp32-equivalent arithmetic (tf32x3 = 3xTF32 split on the tensor cores with TMA-fed GEMMs, fp32 = FFMA kernels): "
                             "1e-5 normwise parity vs the fp64 oracle on every output and gradient; eager launches"}
        del m32

    # ---------------- CPU baseline (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, n, el, thr = cpu_oracle_rate(w, args.cpu_seconds)
        cpu = {"value": rate, "unit": UNIT, "cores": thr, "kind": "port",
               "sample": f"{n} snapshots of the same workload in {el:.1f}s (oracle fp32 port, one snapshot at a time)"}

    if rank == 0:
        job_B = B if sharded else world * B     # snapshots the whole job processes per step
        h2d_bytes = xh.numel() * 4 + yh.numel() * 4
        xport = {"peer": "one-kernel all-reduce over NVLink peer memory (CUDA IPC, csrc/peer.cu)", "nccl": "NCCL all-reduce",
                 "gloo": "gloo all-reduce"}[ex.transport] if world > 1 else ""
        out = {
            "metric": METRIC, "value": job_B / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak",
            "vs_baseline": None, "dtype": {"fp32": "f32", "tf32x3": "f32 (3xTF32 tensor cores)", "bf16": "bf16 operands, f32 accumulate (f32 inputs, outputs, loss, gradients)"}[args.precision],
            "data": "synthetic",
            "config": dict(w.describe(), per_gpu_batch=B, precision=args.precision, micro_batch=mb_used, l2="flushed between timed steps (256 MiB memset)",
                           cuda_graph=graph is not None, exchange_in_graph=graph_has_exchange, optimizer="none: metric is fwd+bwd; the reference steps once per epoch (run.py:194)",
                           parallelism=("single GPU" if world == 1 else
                                        f"region-sharded x{world} (LPT regions->ranks, halo rows of x read locally), {xport} of the flat gradient buffer"
                                        if sharded else
                                        f"batch-sharded x{world}, {xport} of the flat gradient buffer")),
            "clocks": clocks,
            "e2e": {"value": job_B / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms, "pipeline": "double-buffered H2D on a copy stream; each step's loss read on the host one step late"},
            "gpu_launches": int(launches_per_step * K),
            "launches_per_step": int(launches_per_step),
            "roofline": roofline, "roofline_step": step_roof, "kernels": breakdown, "cpu_baseline": cpu,
            "fp32_parity_mode": fp32_mode,
        }
        _REAL_STDOUT.write(json.dumps(out) + "\n")
        _REAL_STDOUT.flush()
    if dist is not None:
        # CUDA graphs that captured NCCL kernels are still alive: a regular communicator teardown can wait on them
        # forever.  Everything is measured and printed; leave without the teardown.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
