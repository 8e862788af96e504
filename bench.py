#!/usr/bin/env python
"""bench.py -- epoch seconds of the embedding-propagation hot path on N B200s (driver contract).

A "step" is one pass of the hot path over one batch: full-graph propagation through the encoder,
fused gather + BPR + L2 loss, backward through the same kernels, Adam.  ``value`` = epoch seconds =
ceil(E / batch) x mean step time with the batch's triples already in HBM; ``e2e`` = the same with the
triples arriving in pinned HOST memory every step and the two loss scalars read back (the reference's
sampler yields CPU LongTensors and its loop calls ``.item()``, model/graph/LightGCN.py:49-59).

Workloads (``--workload``; SURVEY.md section 8 shapes, synthetic power-law graphs, seed 1234):
  c5w (default)  BASELINE configs[4] weak-scaled: 1.25 M users x 0.25 M items x 125 M interactions PER GPU
                 (N = 8 is exactly the 10 M x 2 M x 1 B graph); inputs are larger than L2, no flush needed
  c4 / c3 / c2   the named Amazon-Book / Gowalla / ml-1m shapes on one GPU; L2 is flushed between steps

``--impl reference`` times the reference's own CPU torch path (oracle/torch_path.py, the same torch
calls the reference makes) on the host cores, on a bounded sample of the same workload, and prints
the same JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (users, items, train interactions, batch, scales_with_gpus)
    "c5w": (1_250_000, 250_000, 125_000_000, 1_048_576, True),
    "c4": (52_000, 92_000, 3_000_000, 4096, False),
    "c3": (30_000, 41_000, 1_000_000, 4096, False),
    "c2": (6_040, 3_706, 750_000, 2048, False),
}
WORKLOAD_NOTE = {
    "c5w": "BASELINE configs[4] (10 M x 2 M x 1 B, hypergraph-diffusion encoder, row-sharded over 8 GPUs) weak-scaled: every GPU holds 1/8 of it, "
           "N = 8 is the configuration itself; the per-GPU slice is the largest shape of the list whose tables exceed L2",
    "c4": "BASELINE configs[3] (Amazon-Book shape, hypergraph-diffusion / EquivSetConv)",
    "c3": "BASELINE configs[2] (Gowalla shape; --model hccf is the HCCF + hypergraph SSL configuration)",
    "c2": "BASELINE configs[1] (ml-1m shape, LightGCN 3 layers, emb 64, 1 GPU)",
}
MODELS = {"hgnn_hd3": "HGNN_HD3 local encoder (EquivSetConv + HGCNConv), 2 layers", "lightgcn": "LightGCN, 3 layers",
          "hccf": "HCCF, 2 layers, 128 learned hyperedges, edge keep 0.8, contrastLoss on the batch's unique users/items (temp 0.2)"}
HCCF_CONF = {"lrate": 0.001, "lr_decay": 1.0, "max_epoch": 1, "batch_size": 4096, "reg": 0.0, "embedding_size": 64, "hyper_dim": 128,
             "drop_rate": 0.2, "p": 0.5, "n_layers": 2}
HCCF_TEMP, HCCF_SS_RATE, HCCF_KEEP = 0.2, 0.1, 0.8
D = 64
REG = 0.01
LR = 0.001
CPU_SAMPLE_DIV = {"c5w": 64, "c4": 1, "c3": 1, "c2": 1}
# dram__bytes_read.sum + dram__bytes_write.sum of one spmm_rows_async_kernel launch (ncu --set full, profiles/spmm_r1.md)
SPMM_DRAM_TRAFFIC = {("c5w", 1): 19.96e9, ("c4", 1): 99.8e6}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5w", choices=sorted(WORKLOADS))
    ap.add_argument("--model", default="hgnn_hd3", choices=sorted(MODELS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--no-clocks", action="store_true", help="diagnosis: do not poll nvidia-smi during the timed region")
    ap.add_argument("--cuda-graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the training step as one CUDA graph (auto: single-GPU L2-resident workloads, not hccf)")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def spmm_algorithmic_bytes(n_rows, n_cols, nnz, d):
    """indptr + col + val + gather + write, int64 row offsets as stored.
    Gather = the whole X once when it is L2-resident (<= 64 MB), else one row per nonzero."""
    x_bytes = 4 * d * n_cols
    gather = x_bytes if x_bytes <= 64e6 else 4 * d * nnz
    return 8 * (n_rows + 1) + 4 * nnz + 4 * nnz + gather + 4 * d * n_rows


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        note = None
        if not self.lines:  # timed region shorter than one polling period: one query right after it
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=20).stdout
                self.lines = [ln.strip() for ln in out.splitlines() if ln.strip()]
                note = "timed region shorter than the 200 ms polling period: sampled right after it"
            except (OSError, subprocess.SubprocessError):
                pass
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
               "samples": len(sm)}
        if note:
            out["note"] = note
        return out


def workload_dims(name, n_gpus):
    u, i, e, b, scales = WORKLOADS[name]
    k = n_gpus if scales else 1
    return u * k, i * k, e * k, b * k


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's CPU torch path on a bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(workload, model_name, steps, warmup, n_gpus):
    import numpy as np
    import torch

    from hypergraph_diffusion_for_recommendation_b200.synth import powerlaw_interactions
    from oracle import hgr_oracle as O
    from oracle import torch_path as T

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    U, I, E, B = workload_dims(workload, n_gpus)
    div = CPU_SAMPLE_DIV[workload] * (n_gpus if WORKLOADS[workload][4] else 1)
    su, si, se, sb = max(U // div, 64), max(I // div, 64), max(E // div, 1024), max(B // div, 256)
    g = powerlaw_interactions(su, si, se, seed=1234)
    csr = O.build_norm_adj(g.train_u, g.train_i, su, si)
    adj = T.coo_from_csr(*csr, (su + si, su + si))
    torch.manual_seed(1234)
    if model_name == "hccf":
        model = T.HCCF(adj, su, si, D, 128, 2)
        model.train()
    else:
        model = T.HGNNModel(adj, su, si, D, 2) if model_name == "hgnn_hd3" else T.LGCN(adj, su, si, D, 3)
        model.eval()  # dropout off, like the GPU arm (the reference's loop leaves it off after the first batch)
    opt = torch.optim.Adam(model.parameters(), lr=LR)
    rng = np.random.default_rng(7)
    times = []
    for s in range(warmup + steps):
        pick = rng.integers(0, g.train_u.size, sb)
        u = torch.from_numpy(g.train_u[pick])
        p = torch.from_numpy(g.train_i[pick])
        n = torch.from_numpy(rng.integers(0, si, sb))
        t0 = time.perf_counter()
        if model_name == "hccf":
            T.train_step_hccf(model, opt, u, p, n, HCCF_TEMP, HCCF_SS_RATE, HCCF_KEEP)
        else:
            T.train_step(model, opt, u, p, n, REG, sb)
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    step_s = sum(times) / len(times)
    steps_per_epoch = math.ceil(E / B)
    # work per step is proportional to nnz: scale the sampled step back to the full graph
    epoch_s = step_s * div * steps_per_epoch
    sample = "%d x %d x %d interactions (1/%d of the workload), batch %d, %d threads; step time x %d x %d steps/epoch" % (
        su, si, se, div, sb, cores, div, steps_per_epoch)
    return {"epoch_s": epoch_s, "step_ms_sample": step_s * 1e3, "cores": cores, "sample": sample, "div": div,
            "nnz_per_s": 2 * se * {"hgnn_hd3": 12, "lightgcn": 6, "hccf": 4}[model_name] / step_s}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.workload, args.model, max(1, min(args.steps, 3)), min(args.warmup, 1), args.gpus)
    U, I, E, B = workload_dims(args.workload, args.gpus)
    line = {
        "impl": "reference", "metric": "epoch_s", "value": r["epoch_s"], "unit": "s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["step_ms_sample"] * r["div"], "higher_is_better": False,
        "scaling": "weak" if WORKLOADS[args.workload][4] else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: %d users x %d items x %d interactions, emb %d, %s, batch %d" % (args.workload, U, I, E, D, MODELS[args.model], B)},
        "cpu_baseline": {"value": r["epoch_s"], "unit": "s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["epoch_s"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import types

    import torch
    import torch.distributed as dist

    from hypergraph_diffusion_for_recommendation_b200 import _lib, encoders, ops, trainer
    from hypergraph_diffusion_for_recommendation_b200.synth import powerlaw_interactions_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run --nproc-per-node N)" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libhgr.so has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()

    U, I, E, B = workload_dims(args.workload, world)
    small = not WORKLOADS[args.workload][4]
    u, i = powerlaw_interactions_device(U, I, E, dev, seed=1234)
    part, build_s = None, None
    if world > 1:
        from hypergraph_diffusion_for_recommendation_b200 import dist as hdist

        ctx = hdist.build_partitioned(u, i, U, I, rank, world, dev)
        data, adj, part = ctx.data, ctx.adj, ctx.part
    else:
        # the product builder (csrc/graph_build.cu: radix sort + dedup + LUT normalisation), timed: this is what replaces
        # Interaction.__create_sparse_bipartite_adjacency + normalize_graph_mat (data/ui_graph.py:70-84, data/graph.py:11-25)
        from hypergraph_diffusion_for_recommendation_b200 import graph as hgraph

        torch.cuda.synchronize()
        t_build = time.perf_counter()
        adj = hgraph.build_norm_adj(u, i, U, I, device=dev)
        torch.cuda.synchronize()
        build_s = time.perf_counter() - t_build
        data = types.SimpleNamespace(n_users=U, n_items=I, norm_adj=None, norm_adj_device=adj)
    nnz = int(adj._nnz())
    torch.manual_seed(1234)
    if args.model == "hgnn_hd3":
        model = encoders.HGNNModel(data, {"hyper_dim": D, "n_layers": 2, "p": 0.3, "drop_rate": 0.2, "batch_size": B}).to(dev)
    elif args.model == "hccf":
        if world > 1:
            raise SystemExit("--model hccf is a single-GPU workload (BASELINE configs[2])")
        model = encoders.HCCFEncoder(HCCF_CONF, data).to(dev)
    else:
        model = encoders.LGCN_Encoder(data, D, 3).to(dev)
    if args.model == "hccf":
        model.train()  # HCCF.train calls model.train() every batch (HCCF.py:81): dropout on the learned incidence stays on
    else:
        model.eval()  # dropout off: the reference's HGNN_HD3 loop calls .eval() after its first batch (HGNN_HD3.py:186-204)
    use_graph = world == 1 and (args.cuda_graph == "on" or (args.cuda_graph == "auto" and small))
    optimizer = torch.optim.Adam(model.parameters(), lr=LR, fused=True, capturable=use_graph)

    # triples: the device sampler (csrc/sampler.cu: shuffled positives + rejection-sampled negatives, the body of
    # util/sampler.py:237-264) runs INSIDE every device-timed step; the e2e leg replays host-resident triples, the
    # interface of the reference's sampler (CPU LongTensors).  Every rank draws the same batch (same seed).
    from hypergraph_diffusion_for_recommendation_b200.sampler import PairwiseSampler

    sampler = PairwiseSampler(u, i, U, I, device=dev, seed=99)
    sgen = torch.Generator(device=dev)
    sgen.manual_seed(1234)
    perm = torch.randperm(int(u.numel()), device=dev, generator=sgen)
    n_steps = args.warmup + args.steps
    gen = torch.Generator(device=dev)
    gen.manual_seed(99)  # the same batch on every rank: the sharded step evaluates the loss on the whole batch (dist.py)
    b_local = B
    host_triples, dev_triples = [], []
    for s in range(n_steps):
        pick = torch.randint(0, int(u.numel()), (b_local,), device=dev, generator=gen)
        tu, tp = u[pick].to(torch.int64), i[pick].to(torch.int64)
        tn = torch.randint(0, I, (b_local,), device=dev, generator=gen)
        dev_triples.append((tu, tp, tn))
        host_triples.append(tuple(t.cpu().pin_memory() for t in (tu, tp, tn)))
    eval_inputs = build_eval_inputs(u, i, U, I, part, rank, dev)
    del u, i
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if small else None

    spmm_events = []
    ops.PROFILE_EVENTS = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graphed = None
    if use_graph:
        # warm-up (three Adam steps on an all-zero batch, off the default stream) and capture; the timed steps then replay the
        # graph on the real batches.  The warm-up steps nudge the parameters, which timing does not depend on
        if args.model == "hccf":
            # torch.unique's variable-length result is replaced by the fixed-size sorted-with-gaps form (loss_torch.unique_padded)
            graphed = trainer.GraphedStep(lambda tu, tp, tn: trainer.train_step_hccf(model, optimizer, tu, tp, tn, HCCF_TEMP, HCCF_SS_RATE,
                                                                                     HCCF_KEEP, static_shapes=True), b_local, dev)
        else:
            graphed = trainer.GraphedTrainStep(model, optimizer, REG, B, b_local)

    def step(tri):
        if graphed is not None:
            return graphed(tri[0], tri[1], tri[2])
        if world > 1:
            return hdist.train_step(model, optimizer, adj, tri[0], tri[1], tri[2], REG, B)
        if args.model == "hccf":
            return trainer.train_step_hccf(model, optimizer, tri[0], tri[1], tri[2], HCCF_TEMP, HCCF_SS_RATE, HCCF_KEEP)
        return trainer.train_step(model, optimizer, tri[0], tri[1], tri[2], REG, B)

    def timed(kind):
        """K steps; returns (total ms over the timed steps as max over ranks, last losses)."""
        losses = None
        for s in range(args.warmup):
            losses = step(dev_triples[s])
            if kind == "e2e":
                losses.tolist()
        barrier()
        per_step = []
        launches0 = _lib.launch_count()
        ops.PROFILE_EVENTS = spmm_events if kind == "device" else None
        t_start = torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        if not small:
            t_start.record()
        for s in range(args.warmup, n_steps):
            if small:
                flush.fill_(s & 0xff)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            if kind == "e2e":
                tri = tuple(t.to(dev, non_blocking=True) for t in host_triples[s])
                losses = step(tri)
                host_losses = losses.tolist()  # D2H of the step's result, like the reference's .item() calls
            else:
                off = (s * B) % max(int(perm.numel()) - B, 1)
                losses = step(sampler.batch(perm, off, min(B, int(perm.numel()))))
            if small:
                e1.record()
                per_step.append((e0, e1))
        if not small:
            t_end.record()
        barrier()
        ops.PROFILE_EVENTS = None
        launches = _lib.launch_count() - launches0
        if graphed is not None:  # kernels replayed from the captured graph do not pass through the library's host counter
            launches += graphed.kernels_per_replay * args.steps
        total = sum(a.elapsed_time(b) for a, b in per_step) if small else t_start.elapsed_time(t_end)
        if world > 1:
            t = torch.tensor([total], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        return total, losses, launches

    clocks = ClockSampler(local_rank)
    if not args.no_clocks:
        clocks.start()
    if world > 1:
        adj.n_fused = adj.n_collective = adj.n_published = 0
        adj.phase_events = []
    total_ms, losses, launches = timed("device")
    clock_info = clocks.stop()
    exchange = None
    if world > 1:
        # per step: propagations whose all-gather rode on the kernel epilogue (peer stores) vs NCCL collectives
        exchange = {"fused_gathers_per_step": adj.n_fused / (args.steps + args.warmup), "nccl_gathers_per_step": adj.n_collective / (args.steps + args.warmup),
                    "copy_kernel_gathers_per_step": adj.n_published / (args.steps + args.warmup), "fused": bool(adj.fused)}
        # where the sharded step spends its time on this rank (CUDA events between the phases, timed steps only)
        torch.cuda.synchronize()
        phases = {}
        for marks in adj.phase_events[args.warmup:]:
            for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
                phases[name] = phases.get(name, 0.0) + e0.elapsed_time(e1) / args.steps
        exchange["phases_ms_rank0"] = {k: round(v, 3) for k, v in phases.items()}
        adj.phase_events = None
    e2e_ms, _, _ = timed("e2e")
    eval_info = run_eval(model, eval_inputs, part, adj, world, rank, dev, barrier)

    ms_per_step = total_ms / args.steps
    steps_per_epoch = math.ceil(E / B)
    epoch_s = ms_per_step * steps_per_epoch / 1e3
    e2e_epoch_s = e2e_ms / args.steps * steps_per_epoch / 1e3

    # roofline of the dominant kernel (spmm_rows_async_kernel + its partial-row reduce), timed live
    torch.cuda.synchronize()
    if graphed is not None:
        # a replayed graph has no Python between its kernels to record events from: time the propagation launches of three
        # eager forward passes under the same conditions (L2 flushed); the step holds the same launches forward and backward
        ops.PROFILE_EVENTS = spmm_events
        with torch.no_grad():
            for it in range(3):
                flush.fill_(it)
                model()
        ops.PROFILE_EVENTS = None
        torch.cuda.synchronize()
    spmm_total_ms = sum(a.elapsed_time(b) for a, b, _ in spmm_events)
    spmm_ms = [spmm_total_ms / max(sum(c for _, _, c in spmm_events), 1)] * sum(c for _, _, c in spmm_events)
    peak, peak_src = measured_peaks()
    n_rows, n_cols = (adj.block if world > 1 else adj).shape  # the rank's block of rows when sharded
    alg = spmm_algorithmic_bytes(n_rows, n_cols, nnz, D)
    avg_ms = sum(spmm_ms) / max(len(spmm_ms), 1)
    achieved = alg / (avg_ms * 1e-3) / 1e9 if spmm_ms else None
    traffic = SPMM_DRAM_TRAFFIC.get((args.workload, world))
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                "traffic": traffic,
                # the same launch time against the bytes that really crossed the HBM interface (ncu capture of this kernel and shape)
                "dram_gbs": traffic / (avg_ms * 1e-3) / 1e9 if traffic and spmm_ms else None,
                "dram_frac": traffic / (avg_ms * 1e-3) / 1e9 / peak if traffic and spmm_ms else None, "kernel": "spmm_rows_async_kernel<16,4,5> (+ spmm_heavy_reduce_kernel)", "launch_ms": avg_ms,
                "launches_timed": len(spmm_ms), "algorithmic_bytes": alg, "peak_source": peak_src,
                "spmm_share_of_step": ((2 * len(spmm_ms) / 3 * avg_ms) / ms_per_step if graphed is not None else sum(spmm_ms) / total_ms) if spmm_ms else None,
                "note": "algorithmic bytes charge one 256-B row per nonzero to HBM (SURVEY.md 8d); the power-law graph lets L2 absorb "
                        "about two thirds of that (traffic = dram bytes per launch from ncu, profiles/spmm_r1.md), so achieved can exceed the "
                        "copy peak; ncu: 66.9 GB from L2 to the SMs in 6.0 ms = 5 940 B/clk, ~94 % of the ~6 300 B/clk L2-slice cap; DRAM at 51 % of the copy peak"}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            r = cpu_reference_run(args.workload, args.model, args.cpu_steps, 1, world)
            cpu = {"value": r["epoch_s"], "unit": "s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        line = {
            "metric": "epoch_s", "value": epoch_s, "unit": "s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "weak" if WORKLOADS[args.workload][4] else "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: %d users x %d items x %d interactions, emb %d, %s, batch %d" % (args.workload, U, I, E, D, MODELS[args.model], B),
                       "baseline_config": WORKLOAD_NOTE[args.workload],
                       "steps_per_epoch": steps_per_epoch, "nnz": nnz,
                       "l2": "flushed between steps (256 MiB write)" if small else "inputs larger than L2 (CSR %.1f GB + tables %.2f GB)" % (nnz * 8 / 1e9, (U + I) * D * 4 / 1e9),
                       "cuda_graph": bool(use_graph),
                       "parallelism": "1 GPU" if world == 1 else "row-partitioned x%d, NCCL all-gather per propagation" % world},
            "e2e": {"value": e2e_epoch_s, "unit": "s", "h2d_bytes_per_step": 3 * 8 * b_local, "d2h_bytes_per_step": 8,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches), "clocks": clock_info, "roofline": roofline, "cpu_baseline": cpu,
            "eval": eval_info, "exchange": exchange,
            "graph_build": None if build_s is None else {"seconds": build_s, "interactions": E, "nnz": nnz,
                                                         "what": "device COO -> normalised CSR + split plan (graph.build_norm_adj), one synchronisation"},
            "loss": [float(x) for x in losses.tolist()],
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


EVAL_K = 20
EVAL_MAX_USERS = 262_144  # test users ranked per GPU in the bench (all of them when the shape has fewer)


def build_eval_inputs(u, i, U, I, part, rank, dev):
    """Training matrix rows (the mask) of the users this rank evaluates, as CSR over local user ids."""
    import torch

    u0, u1 = (0, U) if part is None else part.users_of(rank)
    n_own = min(u1 - u0, EVAL_MAX_USERS)
    m = (u >= u0) & (u < u0 + n_own)
    key = torch.sort((u[m].to(torch.int64) - u0) * I + i[m].to(torch.int64)).values
    indptr = torch.zeros(n_own + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.bincount(torch.div(key, I, rounding_mode="floor"), minlength=n_own), 0, out=indptr[1:])
    # held-out interactions of the same users: fresh draws from the generator's distribution that are not training pairs
    # (the reference's 75 / 25 split, dataset_util.py:20-37), as CSR for the metric code
    from hypergraph_diffusion_for_recommendation_b200.synth import powerlaw_interactions_device

    n_draw = max(int(u.numel()) // 3, 1024)
    tu, ti = powerlaw_interactions_device(U, I, n_draw, dev, seed=4321, cover=False, perm_seed=1234)
    tm = (tu >= u0) & (tu < u0 + n_own)
    tkey = torch.unique((tu[tm].to(torch.int64) - u0) * I + ti[tm].to(torch.int64))
    tkey = tkey[~torch.isin(tkey, key)]
    tptr = torch.zeros(n_own + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.bincount(torch.div(tkey, I, rounding_mode="floor"), minlength=n_own), 0, out=tptr[1:])
    return {"n_own": n_own, "indptr": indptr, "indices": (key % I).to(torch.int32), "n_items": I,
            "truth_indptr": tptr.cpu().numpy(), "truth_items": (tkey % I).cpu().numpy()}


def run_eval(model, ev, part, adj, world, rank, dev, barrier, iters=3):
    """Full-ranking evaluation sharded by user: every rank ranks its own users against the whole item table
    (gathered once), top-EVAL_K with the training items masked.  users/s = users of all ranks / max time."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from hypergraph_diffusion_for_recommendation_b200 import evaluation as E

    with torch.no_grad():
        out_u, out_i = model()[:2]
        users = torch.arange(ev["n_own"], device=dev, dtype=torch.int32)
        if world > 1:
            from hypergraph_diffusion_for_recommendation_b200 import dist as hdist

            def rank_all():
                return hdist.fullrank_topk_sharded(adj, out_u, out_i, users, ev["indptr"], ev["indices"], ev["n_items"], EVAL_K,
                                                   return_stats=True)
        else:
            user_tab, item_tab = out_u[:ev["n_own"]].contiguous(), out_i.contiguous()

            def rank_all():
                return E.fullrank_topk(user_tab, item_tab, users, ev["indptr"], ev["indices"], EVAL_K, mode="exact", engine="auto",
                                       return_stats=True)
        times, stats = [], None
        for it in range(iters + 1):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ids, sc, stats = rank_all()
            host_ids = ids.cpu()  # the [n_test, K] id matrix is what the reference's metric code consumes
            e1.record()
            torch.cuda.synchronize()
            if it > 0:
                times.append(e0.elapsed_time(e1))
        ms = sum(times) / len(times)
        n_users_total = ev["n_own"]
        if world > 1:
            t = torch.tensor([ms, float(ev["n_own"])], device=dev, dtype=torch.float64)
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ms, n_users_total = float(tmax[0]), int(t[1])
        # Recall@20 / NDCG@20 of this rank's users through the reference's metric arithmetic (util/evaluation.py), users
        # with held-out items only, exactly like GraphRecommender.test iterates data.test_set; outside the timed region
        has = np.diff(ev["truth_indptr"]) > 0
        sel = np.nonzero(has)[0]
        ptr = np.zeros(sel.size + 1, dtype=np.int64)
        np.cumsum(np.diff(ev["truth_indptr"])[sel], out=ptr[1:])
        t_m = time.perf_counter()
        measures = E.ranking_evaluation_device(ptr, ev["truth_items"], ids[torch.from_numpy(sel).to(dev)], [EVAL_K]) if sel.size else []
        measure_s = time.perf_counter() - t_m
        s = stats.tolist()
        flops = 2.0 * D * ev["n_items"] * n_users_total
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        tf = flops / (ms * 1e-3) / 1e12
        return {"users_per_s": n_users_total / ms * 1e3, "ms": ms, "users": n_users_total, "items": ev["n_items"], "k": EVAL_K,
                "mode": "exact", "d2h_bytes": int(host_ids.numel() * 4), "score_tflops": tf,
                "tensor_frac_of_measured_bf16": tf / (peaks.get("bf16_tflops", 1590.0) * world),
                "candidates_per_user": s[0] / max(ev["n_own"], 1), "rescored_per_user": s[1] / max(ev["n_own"], 1),
                "fallback_users": s[2], "tensor_error_ppm_of_bound": s[3], "metrics_rank0": [m.strip() for m in measures], "metrics_users_rank0": int(sel.size),
                "metrics_s": measure_s}


_JSON_FD = None


def _reserve_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner with printf when the
    first communicator comes up), so everything else is sent to stderr and the line goes out through a saved descriptor."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    args = parse_args()
    _reserve_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
