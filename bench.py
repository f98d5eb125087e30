#!/usr/bin/env python
"""bench.py -- cells/s of the MinHash kNN-graph build (BASELINE.json metric).

Workload (config.workload = "s10"): 10 000 synthetic cells at 1 Mb (nbins 2744, ~14 k non-zero
bin pairs per cell), 800 hash functions, k = 100, exact euclidean rerank (fast=False) -- BASELINE
configs[2], the configuration the metric is quoted on. One "step" = one whole kNN-graph build:
K1 signatures -> (signature / CSR all-gather when N > 1) -> K3 collision counts -> top-k select ->
K4 rerank, neighbour lists left in HBM.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (N>1: launch with torchrun)
  python bench.py --impl reference [...]                        CPU arm: the oracle port on host cores

`value`  : inputs resident in HBM, device-timed with CUDA events, max over ranks.
`e2e`    : same build through the C-ABI host entry points (pinned host CSR in, neighbour lists out),
           H2D and D2H inside the timed region.
`roofline`: the dominant kernel of the step -- K4, the exact rerank -- against the measured HBM copy peak
           (MEASURED_PEAKS.json), algorithmic bytes of SURVEY 8(d); K1 (signatures) is reported beside it with
           the INT32 issue rates measured live by schic_int32_peak.
The total work is fixed as N grows ("scaling": "strong"): the 10 k cells are sharded over the ranks.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "cells/s kNN-graph build, 10k cells@1Mb, 800 hashes"
UNIT = "cells/s"
INT_OPS_PER_EVAL = 15      # algorithmic int32 ops per hash evaluation (reference formulation)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="s10", choices=["s10", "nagano", "s50", "sweep"])
    ap.add_argument("--cells", type=int, default=None, help="override the number of cells (debug)")
    ap.add_argument("--hashes", type=int, default=800)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--hash-width", type=int, default=32, choices=[32, 64],
                    help="HashSpec of SURVEY 8(c) rule 2: evaluate the mix in uint32 (default) or uint64 before the modulo")
    ap.add_argument("--fast", type=int, default=None, help="1: collision distances, 0: exact rerank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: device-resident leg alone")
    return ap.parse_args()


def workload(args):
    from schicexplorer_b200 import synth
    cfg = synth.CONFIGS[args.workload]
    if args.cells:
        cfg = synth.SynthConfig(**{**cfg.__dict__, "n_cells": args.cells})
    fast = args.fast if args.fast is not None else (0 if args.workload == "s10" else 1)
    kw = dict(n_neighbors=args.k, number_of_hash_functions=args.hashes, fast=bool(fast),
              minimal_blocks_in_common=100, excess_factor=1, max_bin_size=100000, shingle_size=5,
              hash_width=args.hash_width)
    return cfg, kw


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (also the cpu_baseline of our arm)
# ---------------------------------------------------------------------------------------------
def oracle_step(ora, kw, k, ip, ix, dt, m):
    """One bounded-sample build on the CPU: the first m cells, fit + kneighbors."""
    h = ora.create(**kw)
    a = int(ip[m])
    t0 = time.perf_counter()
    ora.fit_csr(h, ip[:m + 1], ix[:a], None if kw["fast"] else dt[:a], False)
    ora.kneighbors(h, min(k, m), -1)
    t = time.perf_counter() - t0
    ora.destroy(h)
    return t


def pick_sample(ora, kw, k, ip, ix, dt, budget_s):
    """Largest sample (whole workload if it fits) whose estimated CPU time stays within budget_s."""
    n = len(ip) - 1
    m0 = min(n, 256)
    t0 = oracle_step(ora, kw, k, ip, ix, dt, m0)
    for m in (n, n // 2, n // 4, n // 8, n // 16):
        if m <= m0:
            break
        est = t0 * (m / m0) ** 1.5      # between linear (hashing) and quadratic (collision counting)
        if est <= budget_s:
            return m
    return m0


def run_reference(args):
    """--impl reference: the reference's CPU implementation of this path is the third-party package
    sparse-neighbors-search, absent from /root/reference and not installable, so this
    arm times the oracle port (oracle/oracle.cpp, OpenMP, all host threads) -- kind "port"."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_api
    from schicexplorer_b200 import synth
    cfg, kw = workload(args)
    ora = oracle_api()
    cores = os.cpu_count() or 1
    ip, ix, dt, _ = synth.generate_csr(cfg)
    # the whole arm (warm-up + K steps) stays within a few minutes: ~150 s of CPU work in total
    m = pick_sample(ora, kw, args.k, ip, ix, dt, budget_s=min(25.0, 150.0 / (args.steps + min(args.warmup, 1))))
    for _ in range(max(0, min(args.warmup, 1))):
        oracle_step(ora, kw, args.k, ip, ix, dt, m)
    times = [oracle_step(ora, kw, args.k, ip, ix, dt, m) for _ in range(args.steps)]
    t = float(np.mean(times))
    v = m / t
    sample = f"first {m} of {cfg.n_cells} cells per step (fit + kneighbors k={min(args.k, m)}, fast={int(kw['fast'])})"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": args.workload, "n_cells": cfg.n_cells, "n_hash": args.hashes, "k": args.k,
                   "fast": int(kw["fast"]), "sample_cells": m},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from schicexplorer_b200 import MinHash, _capi, synth
    from schicexplorer_b200.dist import ShardedMinHash, partition_rows

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    # synthetic shard first (forks worker processes), CUDA / NCCL only afterwards
    cfg, kw = workload(args)
    n_total, H, k = cfg.n_cells, args.hashes, args.k
    bounds = partition_rows(n_total, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    workers = max(1, min(32, (os.cpu_count() or 8) // world))
    ip, ix, dt, _ = synth.generate_csr(cfg, lo, hi, workers=workers)
    nnz_local = int(ip[-1])
    fast = kw["fast"]
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    api = _capi.cuda_api()

    # ---- device-resident inputs ("value") ------------------------------------------------
    d_ip = torch.from_numpy(ip).to(dev)
    d_ix = torch.from_numpy(ix.view(np.int32)).to(dev)
    d_dt = None if fast else torch.from_numpy(dt).to(dev)
    # what a caller knows about its input before the first build: the matrix has nbins^2 columns, and how many
    # non-zeros every rank's shard holds (metadata of the row partition, exchanged once)
    n_features = cfg.chroms.nbins ** 2
    if world > 1:
        t_nnz = torch.zeros(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(t_nnz, torch.tensor([nnz_local], dtype=torch.int64, device=dev))
        shard_nnz = [int(x) for x in t_nnz.tolist()]
    else:
        shard_nnz = [nnz_local]
    sh = ShardedMinHash(api, n_total, bounds, dev, n_features=n_features, shard_nnz=shard_nnz, **kw)
    out = (torch.empty((hi - lo, k), dtype=torch.int32, device=dev),
           torch.empty((hi - lo, k), dtype=torch.float32, device=dev),
           torch.empty(hi - lo, dtype=torch.int32, device=dev))

    def step():
        sh.fit_local(d_ip, d_ix, d_dt)
        sh.gather()
        sh.kneighbors_local(k, out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = api.launch_count(sh.h)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_acc = {}
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = api.launch_count(sh.h) - launches0
    stage = api.stage_ms(sh.h)          # stages of the last timed step (CUDA events on the launch stream)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    n_launch = torch.tensor([launches], dtype=torch.int64, device=dev)
    nnz_all = torch.tensor([nnz_local], dtype=torch.int64, device=dev)
    k1_ms = torch.tensor([stage["signatures"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_launch, op=dist.ReduceOp.SUM)
        dist.all_reduce(nnz_all, op=dist.ReduceOp.SUM)
    ms_per_step = float(t_ms.item()) / args.steps
    value = n_total / (ms_per_step * 1e-3)

    # ---- end to end through the C-ABI host entry points ("e2e") -----------------------------
    if args.skip_e2e:
        if rank == 0:
            print(json.dumps({"profiling_only": True, "value": value, "ms_per_step": ms_per_step, "stage_ms": stage,
                              "gpu_launches": int(n_launch.item())}))
        sh.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    p_ip = api.pinned_empty(ip.shape, np.int64); p_ip[:] = ip
    p_ix = api.pinned_empty(ix.shape, np.uint32); p_ix[:] = ix
    p_dt = None
    if not fast:
        # the values are integer contact counts: they cross the C ABI as the host class sends them
        # (MinHash._wire_values: uint8 / uint16 when the data allows, else float32)
        wire = MinHash._wire_values(dt)
        p_dt = api.pinned_empty(wire.shape, wire.dtype); p_dt[:] = wire
    h_idx = torch.empty((hi - lo, k), dtype=torch.int32).pin_memory()
    h_dst = torch.empty((hi - lo, k), dtype=torch.float32).pin_memory()
    h_nv = torch.empty(hi - lo, dtype=torch.int32).pin_memory()
    sh2 = ShardedMinHash(api, n_total, bounds, dev, n_features=n_features, shard_nnz=shard_nnz, **kw)

    host_ms = {"fit_csr": 0.0, "gather": 0.0, "kneighbors": 0.0, "d2h_sync": 0.0}

    def step_e2e():
        t = [time.perf_counter()]
        sh2.fit_local_host(p_ip, p_ix, p_dt, async_upload=True); t.append(time.perf_counter())
        sh2.gather(); t.append(time.perf_counter())
        sh2.kneighbors_local(k, out); t.append(time.perf_counter())
        h_idx.copy_(out[0], non_blocking=True)
        h_dst.copy_(out[1], non_blocking=True)
        h_nv.copy_(out[2], non_blocking=True)
        torch.cuda.synchronize(); t.append(time.perf_counter())
        for name, a, b in zip(host_ms, t[:-1], t[1:]):
            host_ms[name] += (b - a) * 1e3

    for _ in range(2):
        step_e2e()
    barrier()
    for name in host_ms:
        host_ms[name] = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    host_ms = {name: v / args.steps for name, v in host_ms.items()}
    e2e_s = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = n_total / float(e2e_s.item())
    h2d = ip.nbytes + ix.nbytes + (0 if fast else p_dt.nbytes)
    d2h = h_idx.numel() * 4 + h_dst.numel() * 4 + h_nv.numel() * 4
    stage_e2e = api.stage_ms(sh2.h)

    if rank == 0:
        peak = api.int32_peak(local_rank, 2048)
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            hbm_src = "measured (MEASURED_PEAKS.json)"
        except Exception:
            hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        # ---- roofline of the dominant kernel: K4 (exact rerank), HBM bound ---------------------------
        # algorithmic bytes: 8 B (id + value) per non-zero of every candidate row; the units of one
        # launch are this rank's n_local * k candidate pairs of the mean row length
        mean_row = nnz_all.item() / max(1, n_total)
        k4_s = stage["rerank"] * 1e-3
        k4_bytes = 8.0 * mean_row * (hi - lo) * min(k, n_total)
        roofline = None
        if not fast and k4_s > 0:
            gbs = k4_bytes / k4_s / 1e9
            roofline = {
                "kernel": "k4_rerank_window_kernel", "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": gbs / hbm_peak, "traffic": 58.78e9 * (hi - lo) / 10000.0 if args.workload == "s10" else None,
                "peak_source": hbm_src, "kernel_ms": stage["rerank"], "share_of_step": stage["rerank"] / ms_per_step,
                "algorithmic_bytes": k4_bytes,
                "note": "achieved = 8 B x mean row length x (n_local*k candidate pairs) / K4 CUDA-event time of the last timed "
                        "step. frac can exceed 1: 31 % of the sectors hit in L2 and windows the query row does not touch are "
                        "never read, so DRAM traffic (ncu dram__bytes of one S10 launch, profiles/, scaled by rows) is about "
                        "half the algorithmic bytes. ncu: L1/shared data pipe 67 % and issue 68 % busy, DRAM 46 %.",
            }
        # ---- K1 (signatures): reference-formulation rate vs. what is executed ------------------------
        evals = nnz_local * H
        k1_s = float(k1_ms.item()) * 1e-3
        k1 = {
            "kernels": "k1t_mark/build/lookup/fallback (shared table) or k1_filter_kernel", "bound": "int32-issue",
            "k1_ms": float(k1_ms.item()), "share_of_step": float(k1_ms.item()) / ms_per_step,
            "reference_formulation_gevals_per_s": evals / k1_s / 1e9 if k1_s > 0 else 0.0,
            "reference_formulation_tintops": evals * INT_OPS_PER_EVAL / k1_s / 1e12 if k1_s > 0 else 0.0,
            "int32_issue_peak_tintops": peak["mix_lop3_imad"],
            "note": "SURVEY 8(d) counts nnz*H evaluations x 15 int32 ops. The shared-table path evaluates each (distinct id, "
                    "hash function) once per batch (S10: 3.77e6 ids x 800 = 3.0e9 evaluations, 1.26 ms, ALU pipe 79 % busy) "
                    "and looks the rest up, so the reference-formulation rate exceeds the issue peak many times over without "
                    "any evaluation being approximated (bit-exact vs the oracle). Per-cell filtered kernel for comparison: "
                    "35.6 ms, 9.3 issued ops/evaluation, ALU pipe 91 % busy.",
            "int32_peaks_tops": peak,
        }
        if roofline is None:      # fast path (no rerank): K3 is the dominant kernel, INT32 issue bound
            k3_s = stage["collisions"] * 1e-3
            compares = (hi - lo) * n_total * H * (0.5 if world == 1 else 1.0)
            ach = compares / k3_s / 1e12 if k3_s > 0 else 0.0
            roofline = {"kernel": "k3_collision_kernel", "bound": "int32-issue", "achieved": ach, "peak": peak["lop3"],
                        "unit": "Tcompare/s", "frac": ach / peak["lop3"], "traffic": None,
                        "note": "one ISETP (ALU pipe) per signature compare; peak = measured single-pipe issue rate"}
        roofline["k1_signatures"] = k1
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": args.workload, "n_cells": n_total, "n_hash": H, "k": k, "fast": int(fast),
                       "hash_width": args.hash_width,
                       "nnz_total": int(nnz_all.item()), "sharding": f"cells/{world}",
                       "l2": "inputs (CSR %.2f GB/rank) larger than L2" % (h2d / 1e9)},
            "stage_ms": stage, "gpu_launches": int(n_launch.item()), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": float(e2e_s.item()) * 1e3, "stage_ms": stage_e2e, "host_call_ms": host_ms},
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle_api
            ora = oracle_api()
            m = pick_sample(ora, kw, k, ip, ix, dt, budget_s=20.0)
            t = oracle_step(ora, kw, k, ip, ix, dt, m)
            line["cpu_baseline"] = {"value": m / t, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"first {m} of {n_total} cells, fit + kneighbors k={min(k, m)}, "
                                              f"fast={int(fast)}, {t:.1f} s on the host cores (oracle/oracle.cpp, OpenMP)"}
        print(json.dumps(line))
    sh.close()
    sh2.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
