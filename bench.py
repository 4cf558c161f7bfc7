#!/usr/bin/env python3
"""bench.py — GCUPS of the B200-native Smith-Waterman path on the BASELINE config-3 shape.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                      the reference's own CPU path (oracle/_ref)

A "step" is one pass of the alignment hot path (score pass + arg-max + traceback, i.e. what
SWAligner::calculateScore does per read, smithwaterman.cpp:80-108) over one batch of synthetic reads.
Workload: `--reads` 150 bp reads per GPU per step against a seeded 1 Mbp synthetic
reference, SAT_U8 arithmetic (Similarity_Matrix_Skewed semantics), default scoring +3/-3, gap 2.
GCUPS counts cells the way the reference drivers do: sum len(read) * len(ref) (sw_solve_small.cpp:89).

One JSON line is printed by rank 0 (see README / DESIGN.md §Measurement for every key).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GCUPS (device-timed, whole box) at 1/2/4/8 B200 vs host OpenMP; % ALU roofline"
UNIT = "GCUPS"
OPS_PER_CELL = {"SAT_U8": 9, "EXACT": 8}   # SURVEY.md §8(d) contract figure
READ_LEN, REF_LEN = 150, 1_000_000


def load_peaks():
    """Measured peaks: integer-ALU lane-ops/s from our microbenchmark (profiles/alu_peak_r*.json, VIADDMNMX /
    VIMNMX / IADD3 all issue at the same rate) and HBM GB/s from the driver-written MEASURED_PEAKS.json."""
    p_int, src = 148 * 64 * 1.965e9 / 1e12, "nominal 148 SM x 64 lanes x 1.965 GHz (fallback)"
    prof = os.path.join(ROOT, "profiles")
    if os.path.isdir(prof):
        for fn in sorted(os.listdir(prof), reverse=True):
            if fn.startswith("alu_peak_r") and fn.endswith(".json"):
                with open(os.path.join(prof, fn)) as f:
                    rows = {r["inst"]: r["tera_lane_ops_per_s"] for r in json.load(f)["rows"]}
                p_int, src = rows["VIADDMNMX.S16x2"], f"measured, profiles/{fn} (VIADDMNMX.S16x2 lane-ops/s)"
                break
    hbm, hsrc = 6650.0, "fallback"
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(mp):
        with open(mp) as f:
            hbm, hsrc = json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return p_int, src, hbm, hsrc


def load_traffic(n_reads):
    """dram__bytes_read + dram__bytes_write of one score-kernel launch from the committed ncu capture, when it was
    taken at this batch size (profiles/ncu_traffic_r*.json); None otherwise."""
    prof = os.path.join(ROOT, "profiles")
    if os.path.isdir(prof):
        for fn in sorted(os.listdir(prof), reverse=True):
            if fn.startswith("ncu_traffic_r") and fn.endswith(".json"):
                with open(os.path.join(prof, fn)) as f:
                    k = json.load(f)["score_kernel"]
                if k["reads_per_gpu"] == n_reads:
                    return k["dram_bytes_read"] + k["dram_bytes_write"], k.get("alu_pipe_busy")
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_batch(pkg, n_reads, seed):
    ref = pkg.synth.c3_reference(REF_LEN)
    ref_u8 = np.frombuffer(ref.encode("ascii"), dtype=np.uint8)
    reads = pkg.synth.mutated_reads_fast(ref_u8, n_reads, READ_LEN, seed=seed, sub=0.01)
    return ref, reads


def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path: SWAligner<Similarity_Matrix_Skewed> per read
    (= sw_solve_small.cpp:84-88) compiled from /root/reference into oracle/_ref, with a harness-level
    OpenMP loop over reads on all host cores.  Each step is a bounded sample."""
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as o
    pkg = importlib.import_module("parallel-genomeseq_b200")
    cores = os.cpu_count() or 1
    kind = "reference"
    n_sample = args.cpu_reads or 2 * cores
    ref, reads = make_batch(pkg, n_sample, seed=1000)
    read_list = [reads[i].tobytes().decode("ascii") for i in range(n_sample)]
    if o.ref() is None:
        kind = "port"
    times, it_us = [], []
    for s in range(args.warmup + args.steps):
        if kind == "reference":
            wall, us, _, _ = o.ref_bench_reads(read_list, ref, smt=0, nthreads=cores)
        else:
            t0 = time.perf_counter()
            for x in read_list[:max(1, n_sample // cores)]:
                o.align(x, ref)
            wall, us = (time.perf_counter() - t0), 0.0
        if s >= args.warmup:
            times.append(wall); it_us.append(us)
    n_eff = n_sample if kind == "reference" else max(1, n_sample // cores)
    cells = n_eff * READ_LEN * REF_LEN
    sec = float(np.mean(times))
    gcups = cells / sec / 1e9
    gcups_iter = cells * (cores if kind == "reference" else 1) / (float(np.mean(it_us)) * 1e-6) / 1e9 if kind == "reference" and np.mean(it_us) > 0 else None
    sample = f"{n_eff} reads x {READ_LEN} bp vs {REF_LEN} bp per step, whole calculateScore() (alloc + iterate + maxCoeff + traceback) by wall clock"
    if gcups_iter:
        sample += f"; iterate()-only by the reference's own convention (sw_solve_small.cpp:88-89): {gcups_iter:.2f} GCUPS over {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": gcups, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"C3 batched read mapping: {READ_LEN} bp reads vs {REF_LEN} bp synthetic reference, SAT_U8 (+3/-3, gap 2), bounded CPU sample"},
            "cpu_baseline": {"value": gcups, "unit": UNIT, "cores": cores if kind == "reference" else 1, "kind": kind, "sample": sample},
            "e2e": {"value": gcups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def cuda_array(ptr, n, typestr):
    class _A:
        pass
    a = _A()
    a.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}
    return a


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=151552, help="reads per GPU per step")
    ap.add_argument("--cpu-reads", type=int, default=0, help="reads in the CPU-baseline sample (default 2 per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, rank, world)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("parallel-genomeseq_b200")
    sharding = importlib.import_module("parallel-genomeseq_b200.sharding")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    use_dist = world > 1
    if use_dist:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    p_int, p_src, hbm_peak, hbm_src = load_peaks()

    # ---- inputs: rank-private shard of reads (weak scaling), replicated reference -------------------------
    ref, reads = make_batch(pkg, args.reads, seed=2300 + rank)
    n_reads = reads.shape[0]
    blob = torch.from_numpy(reads.reshape(-1).copy()).pin_memory()
    offs = torch.arange(n_reads + 1, dtype=torch.int64) * READ_LEN
    offs_np = offs.numpy().astype(np.uint64)
    blob_np = blob.numpy()
    cells_step = n_reads * READ_LEN * REF_LEN
    eng = pkg.Engine(local_rank)
    eng.set_scoring_match(pkg.MODE_SAT_U8, 3, -3, 2)
    eng.set_reference(ref)
    cons_stride = 2 * READ_LEN + 64
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_results():
        """the one collective of the path: (score, pos) of every read to all ranks over NCCL (8 B per read)"""
        if not use_dist:
            return 0.0
        ds, dp = eng.device_results()
        s = torch.as_tensor(cuda_array(ds, n_reads, "<i4"), device="cuda")
        p = torch.as_tensor(cuda_array(dp, n_reads, "<u4"), device="cuda").view(torch.int32)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sharding.gather_score_pos(s, p)
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) * 1e3

    # ---- device-resident arm: inputs staged in HBM once, K timed passes -----------------------------------
    eng.stage((blob_np, offs_np), consensus=True, cons_stride=cons_stride)
    for _ in range(args.warmup):
        eng.run(); gather_results()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    dev_us, p1_us, p2_us, launches = [], [], [], 0
    for _ in range(args.steps):
        flush.fill_(1)                    # L2 flush between timed iterations (not inside the event-timed region)
        torch.cuda.synchronize()
        us = eng.run()
        us += gather_results()
        st = eng.stats()
        dev_us.append(us); p1_us.append(st["pass1_us"]); p2_us.append(st["pass2_us"]); launches += st["kernel_launches"]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    st = eng.stats()
    res = eng.fetch()

    # ---- end-to-end arm: host buffers in, host buffers out, through the one-call C ABI ---------------------
    h2d = blob_np.nbytes + offs_np.nbytes
    d2h = n_reads * (4 + 4 + 8 + 4 + 4) + 2 * n_reads * cons_stride
    for _ in range(2):
        eng.align((blob_np, offs_np), consensus=True, cons_stride=cons_stride, decode=False)
    barrier()
    e2e_t = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        out = eng.align((blob_np, offs_np), consensus=True, cons_stride=cons_stride, decode=False)
        if use_dist:
            gather_results()
        torch.cuda.synchronize()
        e2e_t.append(time.perf_counter() - t0)
    barrier()
    assert (out["score"] == res["score"]).all() and (out["pos"] == res["pos"]).all()

    # ---- reduce over ranks: max time, sum of work ------------------------------------------------------------
    step_us = float(np.mean(dev_us))
    e2e_s = float(np.mean(e2e_t))
    vec = torch.tensor([step_us, e2e_s, float(np.mean(p1_us))], dtype=torch.float64, device="cuda")
    if use_dist:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
    step_us, e2e_s, pass1_us = [float(v) for v in vec.tolist()]
    total_cells = cells_step * world
    gcups = total_cells / (step_us * 1e-6) / 1e9
    e2e_gcups = total_cells / e2e_s / 1e9

    if rank == 0:
        ops = OPS_PER_CELL["SAT_U8"]
        traffic, alu_busy = load_traffic(n_reads)
        k_gcups = cells_step / (pass1_us * 1e-6) / 1e9          # dominant kernel (score pass), this GPU
        achieved = k_gcups * 1e9 * ops / 1e12                    # algorithmic Tops/s
        hbm_bytes = h2d + st["cells_executed"] / (2 * st["rows_per_lane"] * st["block_steps"]) * (st["rows_per_lane"] + 2) * 4  # checkpoints + block maxima
        line = {
            "metric": METRIC, "value": gcups, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_us / 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "s16x2 (u8-saturating semantics)", "data": "synthetic",
            "config": {"workload": f"C3 batched read mapping: {READ_LEN} bp reads vs {REF_LEN} bp synthetic reference, SAT_U8 (+3/-3, gap 2), score+argmax+traceback",
                       "reads_per_gpu_per_step": n_reads, "cells_per_step_per_gpu": cells_step, "parallelism": f"reads sharded over {world} GPU(s), reference replicated, one NCCL all-gather of (score,pos)",
                       "l2": "flushed (256 MiB write) between timed iterations; per-step work buffers (checkpoints) exceed L2",
                       "geometry": {"lanes_per_pair": st["lanes_per_pair"], "rows_per_lane": st["rows_per_lane"], "block_steps": st["block_steps"]}},
            "e2e": {"value": e2e_gcups, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "alu", "achieved": achieved, "peak": p_int, "unit": "Tops/s (int lane-ops)", "frac": achieved / p_int,
                         "traffic": traffic, "traffic_unit": "bytes per score_kernel launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                         "alu_pipe_busy_ncu": alu_busy, "ops_per_cell": ops, "kernel": "score_kernel (pass 1)", "kernel_gcups": k_gcups, "kernel_ms": pass1_us / 1e3,
                         "kernel_share_of_step": pass1_us / step_us, "peak_source": p_src,
                         "executed_cell_fraction": cells_step / max(1, st["cells_executed"]),
                         "hbm": {"achieved_gbs": hbm_bytes / (pass1_us * 1e-6) / 1e9, "peak_gbs": hbm_peak, "peak_source": hbm_src}},
            "pass2_ms": float(np.mean(p2_us)) / 1e3, "wall_ms_per_step": t_wall / args.steps * 1e3,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(pkg, args)
        print(json.dumps(line), flush=True)
    eng.close()
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def cpu_baseline(pkg, args):
    """oracle/_ref (the reference compiled from /root/reference) timed on this box's host cores on a bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as o
    cores = os.cpu_count() or 1
    n_sample = args.cpu_reads or 2 * cores
    ref, reads = make_batch(pkg, n_sample, seed=1000)
    read_list = [reads[i].tobytes().decode("ascii") for i in range(n_sample)]
    if o.ref() is not None:
        o.ref_bench_reads(read_list[:cores], ref, smt=0, nthreads=cores)
        wall, us, _, _ = o.ref_bench_reads(read_list, ref, smt=0, nthreads=cores)
        cells = n_sample * READ_LEN * REF_LEN
        return {"value": cells / wall / 1e9, "unit": UNIT, "cores": cores, "kind": "reference",
                "iterate_only_gcups": cells * cores / (us * 1e-6) / 1e9 if us > 0 else None,
                "sample": f"{n_sample} reads x {READ_LEN} bp vs {REF_LEN} bp, SWAligner<Similarity_Matrix_Skewed> per read (oracle/_ref), OpenMP over reads on {cores} threads, whole calculateScore() by wall clock"}
    t0 = time.perf_counter()
    k = max(1, n_sample // cores)
    for x in read_list[:k]:
        o.align(x, ref)
    wall = time.perf_counter() - t0
    return {"value": k * READ_LEN * REF_LEN / wall / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{k} reads x {READ_LEN} bp vs {REF_LEN} bp, scalar C restatement (oracle/sw_oracle.c), 1 thread"}


if __name__ == "__main__":
    sys.exit(main())
