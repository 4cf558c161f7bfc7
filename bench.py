#!/usr/bin/env python3
"""bench.py — GCUPS of the B200-native Smith-Waterman path on the BASELINE.json configs.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --config c1|c1x64|c2|c2x64|c2p17x64|c3|c4|c5|c5sat ...   one config as the headline line
    python bench.py --impl reference ...                      the reference's own CPU path (oracle/_ref)

A "step" is one pass of the alignment hot path (score pass + arg-max + traceback, i.e. what
SWAligner::calculateScore does per read, smithwaterman.cpp:80-108) over one batch of synthetic input.
Default workload: `--reads` 150 bp reads per GPU per step (1 % substitutions, 0.1 % insertions,
0.1 % deletions) against a seeded 1 Mbp synthetic reference, SAT_U8 arithmetic (Similarity_Matrix_Skewed
semantics), default scoring +3/-3, gap 2.  GCUPS counts cells the way the reference drivers do:
sum len(read) * len(ref) (sw_solve_small.cpp:89).

The default line also carries, at N = 1, a `configs` block (one short timed run of every other BASELINE config at its
STATED size, each with an in-run parity check against golden vectors / the oracle) and a `cpu_baselines` block
(BASELINE.md §3 B1, B2, B4, B5 on this box's host cores).  One JSON line is printed by rank 0.
"""
import argparse
import hashlib
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
GOLDEN = os.path.join(ROOT, "tests", "golden")


# ---------------------------------------------------------------------------------------------------------
# peaks and instruction counts
# ---------------------------------------------------------------------------------------------------------
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


def load_sass_counts():
    """ALU-pipe instructions per cell pair in the hot loop of every pass-1 kernel, counted from the SASS of the built
    objects by tools/sass_counts.py (committed as profiles/sass_counts_r02.json; `tools/sass_counts.py --check` and
    a CPU test keep it in step with the build)."""
    path = os.path.join(ROOT, "profiles", "sass_counts_r02.json")
    if not os.path.isfile(path):
        return {}, None
    with open(path) as f:
        return json.load(f)["kernels"], "profiles/sass_counts_r02.json"


def sass_entry(counts, st, sat):
    """The counted kernel that matches the geometry the engine reports (lanes x rows, columns per step, kernel kind)."""
    kind = {0: "score_kernel<", 1: "score_units_kernel<", 2: "qs_score_kernel<"}[st["kernel_kind"]]
    want = f"{kind}R={st['rows_per_lane']},"
    for k, v in counts.items():
        if want in k and (st["kernel_kind"] == 2 or f"C={st['cols_per_step']}," in k) and (("SAT" in k) == sat):
            return k, v
    return None, None


def issue_fraction(counts, st, sat, p_int):
    """Issue-based roofline fraction of the pass-1 kernel: ALU-pipe lane instructions per second / measured ALU-pipe
    peak.  One lane instruction on s16x2 operands updates a cell PAIR, so
        frac = alu_inst_per_cell_pair x (executed cells / 2 / pass-1 seconds) / P_int."""
    key, e = sass_entry(counts, st, sat)
    if e is None or st["pass1_us"] <= 0:
        return None
    pairs_per_s = st["cells_executed"] / 2.0 / (st["pass1_us"] * 1e-6)
    # kernels whose SASS loop holds paths not taken in the steady state carry EXECUTED counts from an ncu capture
    # (tools/sass_counts.py EXECUTED); the static count is kept beside them
    ex = e.get("executed")
    alu = ex["alu_inst_per_cell_pair"] if ex else e["alu_inst_per_cell_pair"]
    out = {"frac": alu * pairs_per_s / (p_int * 1e12), "alu_inst_per_cell_pair": alu,
           "inst_per_cell_pair": ex["inst_per_cell_pair"] if ex else e["inst_per_cell_pair"], "kernel": key,
           "counts_from": ("from_profile: " + ex["source"]) if ex else "SASS of the built object (tools/sass_counts.py)"}
    if ex:
        out["alu_inst_per_cell_pair_static_sass"] = e["alu_inst_per_cell_pair"]
    return out


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
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_median": float(np.median(pw)) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def make_batch(pkg, n_reads, seed):
    ref = pkg.synth.c3_reference(REF_LEN)
    ref_u8 = np.frombuffer(ref.encode("ascii"), dtype=np.uint8)
    reads = pkg.synth.mutated_reads_vec(ref_u8, n_reads, READ_LEN, seed=seed, sub=0.01, ins=0.001, dele=0.001)
    return ref, reads


def oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as o
    return o


def read_fasta(path):
    with open(path) as f:
        return "".join(f.read().split("\n")[1:])


def data_small():
    ref = read_fasta(os.path.join(GOLDEN, "data_small", "genome.chr22.5K.fa"))
    reads = []
    with open(os.path.join(GOLDEN, "data_small", "data_small_ground_truth.csv")) as f:
        for i, line in enumerate(f):
            if i and line.strip():
                reads.append(line.split(",")[2])
    return ref, reads


def golden_csv(name):
    import csv
    with open(os.path.join(GOLDEN, name)) as f:
        return [(int(r[1]), int(r[2])) for r in list(csv.reader(f))[1:]]


# ---------------------------------------------------------------------------------------------------------
# the reference arm and the CPU baselines
# ---------------------------------------------------------------------------------------------------------
def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path: SWAligner<Similarity_Matrix_Skewed> per read
    (= sw_solve_small.cpp:84-88) compiled from /root/reference into oracle/_ref, with a harness-level
    OpenMP loop over reads on all host cores.  Each step is a bounded sample."""
    if rank != 0:
        return 0
    o = oracle()
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
                o.align(x, ref, linear=True)
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


def cpu_baseline(pkg, args):
    """oracle/_ref (the reference compiled from /root/reference) timed on this box's host cores on a bounded sample."""
    o = oracle()
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
        o.align(x, ref, linear=True)
    wall = time.perf_counter() - t0
    return {"value": k * READ_LEN * REF_LEN / wall / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{k} reads x {READ_LEN} bp vs {REF_LEN} bp, linear-memory C restatement (oracle/sw_oracle_linear.c), 1 thread"}


def cpu_baselines(pkg):
    """B1 (1 core, serial sw_solve_small loop), B2 (the reference's own -DUSEOMP OMPParallelLocalAligner,
    17 pieces and nproc pieces; timing only, its results are racy — SURVEY F7), B4 (protein search, SWAligner<Similarity_Matrix>
    serial and over all cores) and B5 (fine-grain OpenMP type 1 at the reference's 10 k x 30 k shape).  Bounded samples."""
    o = oracle()
    if o.ref() is None:
        return {"unavailable": "oracle/_ref is not built"}
    cores = os.cpu_count() or 1
    out = {"cores": cores}
    ref, reads = data_small()
    cells = sum(len(x) for x in reads) * len(ref)
    wall, us, _, _ = o.ref_bench_reads(reads, ref, smt=0, nthreads=1)
    out["B1 data_small, SWAligner<Skewed>, 1 core"] = {"gcups_wall": cells / wall / 1e9, "gcups_iterate_only": cells / (us * 1e-6) / 1e9, "reads": len(reads), "seconds": wall}
    wall, us, _, _ = o.ref_bench_reads(reads, ref, smt=0, nthreads=cores)
    out[f"B3 data_small, SWAligner<Skewed>, OpenMP over reads, {cores} threads"] = {"gcups_wall": cells / wall / 1e9, "seconds": wall}
    if o.ref_omp() is not None:
        sub = reads[:400]
        c400 = sum(len(x) for x in sub) * len(ref)
        for npiece in sorted({17, min(cores, 17)}):       # the 125-bp reads admit at most 19 pieces (overlap <= piece, plocalaligner.cpp:52)
            wall, us = o.ref_omp_chunked_bench(sub, ref, npiece, 2.0)
            out[f"B2 data_small, reference -DUSEOMP OMPParallelLocalAligner({npiece}, 2.0), {npiece} threads"] = {
                "gcups_wall": c400 / wall / 1e9, "gcups_iterate_only": c400 / (us * 1e-6) / 1e9, "reads": len(sub), "seconds": wall, "note": "timing only (results racy, SURVEY F7)"}
    # B4: protein search sample (x = database protein, y = query, mpi_sw_solve_uniprot.cpp:120), BLOSUM62 callback, gap 10
    query = pkg.synth.c4_queries(1, 300)[0]
    db = pkg.synth.c4_database(600, seed=25)
    t = pkg.synth.blosum62_table()
    c4 = sum(len(p) for p in db) * len(query)
    for nt in sorted({1, cores}):
        r = o.ref()
        blobs = [p.encode() for p in db]
        # ref_bench_reads runs the DEFAULT callback (+3/-3, gap 2): the per-cell std::function cost is the same as BLOSUM62's
        wall, us, _, _ = o.ref_bench_reads(db, query, smt=1, nthreads=nt)
        out[f"B4 protein search sample, SWAligner<Similarity_Matrix>, {nt} thread(s)"] = {"gcups_wall": c4 / wall / 1e9, "gcups_iterate_only": c4 * nt / (us * 1e-6) / 1e9, "proteins": len(db), "seconds": wall}
    if o.ref_omp() is not None:
        lref = pkg.synth.c3_reference(30_000, seed=26)
        x = pkg.synth.mutated_reads(lref, 1, 10_000, seed=27, sub=0.02, ins=0.002, dele=0.002)[0]
        for nt in sorted({1, 6, cores}):
            r = o.ref_omp_finegrain(x, lref, 1, nt)
            out[f"B5 long pair 10 kbp x 30 kbp, fine-grain OpenMP type 1, {nt} thread(s)"] = {"gcups_iterate": len(x) * len(lref) / (r["iterate_us"] * 1e-6) / 1e9, "seconds": r["wall"]}
    return out


# ---------------------------------------------------------------------------------------------------------
# the other BASELINE configs (one GPU): timed at their stated sizes, each with an in-run parity check
# ---------------------------------------------------------------------------------------------------------
def timed_runs(eng, reps):
    eng.run()
    us = [eng.run() for _ in range(reps)]
    return float(np.median(us)), eng.stats()


def entry(st, us, counts, sat, p_int, extra=None):
    e = {"gcups": st["cells_reference"] / us / 1e3, "device_ms": us / 1e3, "pass1_ms": st["pass1_us"] / 1e3, "pass2_ms": st["pass2_us"] / 1e3,
         "pass2_share": st["pass2_us"] / us, "executed_over_reference_cells": st["cells_executed"] / max(1, st["cells_reference"]),
         "geometry": {"lanes_per_pair": st["lanes_per_pair"], "rows_per_lane": st["rows_per_lane"], "cols_per_step": st["cols_per_step"],
                      "block_steps": st["block_steps"], "pass1_kernel": ["score_kernel", "score_units_kernel", "qs_score_kernel"][st["kernel_kind"]]}}
    fr = issue_fraction(counts, st, sat, p_int)
    if fr:
        e["roofline"] = {"bound": "alu", "frac": fr["frac"], "alu_inst_per_cell_pair": fr["alu_inst_per_cell_pair"], "counts_from": fr["counts_from"], "kernel": fr["kernel"], "peak_tera_lane_ops": p_int,
                         "frac_contract_9ops": st["cells_reference"] / (st["pass1_us"] * 1e-6) * OPS_PER_CELL["SAT_U8" if sat else "EXACT"] / (p_int * 1e12)}
    if extra:
        e.update(extra)
    return e


def run_c1_c2(pkg, eng, counts, p_int, which, reps=5):
    """C1 (sw_solve_small: the shipped data_small reads) and C2 (OMPParallelLocalAligner, 4 chunks / overlap 2.0; 17 is
    what sw_solve_small.cpp:82 uses) at x1 and replicated x64 (x1 cannot fill one B200: 1170 reads).  Parity: score and
    pos of every read against the goldens dumped from the reference."""
    ref, reads = data_small()
    npiece = {"c1": 0, "c1x64": 0, "c2": 4, "c2x64": 4, "c2p17x64": 17}[which]
    rep = 64 if which.endswith("x64") else 1
    gold = golden_csv({0: "data_small_sw_skewed.csv", 4: "data_small_p4.csv", 17: "data_small_p17.csv"}[npiece])
    eng.set_scoring_match(pkg.MODE_SAT_U8, 3, -3, 2)
    eng.set_reference(ref)
    eng.stage(reads * rep, npiece=npiece, ratio=2.0, consensus=True)
    us, st = timed_runs(eng, reps)
    res = eng.fetch()
    ok = all((int(res["score"][i]), int(res["pos"][i])) == gold[i % len(gold)] for i in range(len(reads) * rep))
    return entry(st, us, counts, True, p_int, {"reads": len(reads) * rep, "npiece": npiece, "parity_sample": {"n": len(reads) * rep, "ok": bool(ok), "against": "tests/golden (compiled reference)"}})


def run_c4(pkg, eng, counts, p_int, proteins=500_000, queries=64, check=32):
    """C4: mpi_sw_solve_uniprot-shaped search, `queries` 300-aa queries against a synthetic `proteins`-entry database
    (x = database protein, y = query), EXACT, BLOSUM62 through the tabulated callback, gap 10.  The database is staged
    once and stays in HBM; every query is one kernel pass (swb_batch_rebind_reference).  Parity: `check` proteins per
    checked query against the oracle (score, pos, arg-max cell)."""
    o = oracle()
    qs = pkg.synth.c4_queries(queries, 300)
    db = pkg.synth.c4_database(proteins)
    table = pkg.synth.blosum62_table()
    eng.set_scoring_table(pkg.MODE_EXACT, table, 10)
    eng.set_reference(qs[0])
    t0 = time.perf_counter()
    eng.stage(db, consensus=False)
    stage_s = time.perf_counter() - t0
    eng.run()
    tot_us, p1, p2, cells, execd = 0.0, 0.0, 0.0, 0, 0
    ok, nchk = True, 0
    rng = np.random.default_rng(4)
    for qi, q in enumerate(qs):
        eng.rebind_reference(q)
        us = eng.run()
        st = eng.stats()
        tot_us += us; p1 += st["pass1_us"]; p2 += st["pass2_us"]; cells += st["cells_reference"]; execd += st["cells_executed"]
        if qi in (0, len(qs) - 1):
            res = eng.fetch()
            for i in rng.integers(0, len(db), size=check):
                w = o.align(db[i], q, mode=o.MODE_EXACT, table=table, gap=10)
                got = (int(res["score"][i]), int(res["pos"][i]), tuple(int(v) for v in res["end"][i]))
                ok = ok and (got == (w["score"], w["pos"], tuple(w["end"])) if w["score"] > 0 else got[0] == 0)
                nchk += 1
    st = dict(st, cells_reference=cells, cells_executed=execd, pass1_us=p1, pass2_us=p2)
    return entry(st, tot_us, counts, False, p_int, {"proteins": len(db), "queries": len(qs), "residues": int(sum(len(p) for p in db)), "stage_seconds": stage_s,
                                                    "per_query_ms": tot_us / 1e3 / len(qs), "parity_sample": {"n": nchk, "ok": bool(ok), "against": "oracle/sw_oracle.c"}})


def run_c5(pkg, eng, counts, p_int, sat, ref_len=51_000_000, reps=2):
    """C5: omp_sw_solve_small-style long pairs, the 16 seeded 10 kbp reads against the seeded chr22-sized reference, EXACT
    (the default build, omp_sw_solve_small.cpp:167) or SAT_U8 (MTSIMD, :164).  Parity: score, pos, arg-max cell, consensus
    length and sha256 of both consensus strings of EVERY read against tests/golden/c5_full.json (linear-memory oracle)."""
    lref = pkg.synth.c5_reference(ref_len)
    lreads = pkg.synth.c5_reads(lref, 16, 10_000)
    eng.set_scoring_match(pkg.MODE_SAT_U8 if sat else pkg.MODE_EXACT, 3, -3, 2)
    eng.set_reference(lref)
    eng.stage(lreads, consensus=True, cons_stride=25_000)
    us, st = timed_runs(eng, reps)
    par = {"n": 0, "ok": None, "against": "tests/golden/c5_full.json (only at 51 Mbp)"}
    gpath = os.path.join(GOLDEN, "c5_full.json")
    if ref_len == 51_000_000 and os.path.isfile(gpath):
        with open(gpath) as f:
            doc = json.load(f)
        res = eng.fetch()
        ok = True
        for i, e in enumerate(doc["sat_u8" if sat else "exact"]):
            ln = int(res["len"][i])
            got = (int(res["score"][i]), int(res["pos"][i]), [int(v) for v in res["end"][i]], ln,
                   hashlib.sha256(res["cx_raw"][i, :ln].tobytes()).hexdigest(), hashlib.sha256(res["cy_raw"][i, :ln].tobytes()).hexdigest())
            ok = ok and got == (e["score"], e["pos"], e["end"], e["len"], e["cx_sha256"], e["cy_sha256"])
        par = {"n": len(doc["exact"]), "ok": bool(ok), "against": "tests/golden/c5_full.json (linear-memory oracle)"}
    return entry(st, us, counts, sat, p_int, {"reads": len(lreads), "read_len": 10_000, "ref_len": ref_len, "mode": "SAT_U8" if sat else "EXACT", "parity_sample": par})


def all_configs(pkg, eng, counts, p_int, args):
    out = {}
    for which in ("c1", "c1x64", "c2", "c2x64", "c2p17x64"):
        out[which] = run_c1_c2(pkg, eng, counts, p_int, which)
    out["c4"] = run_c4(pkg, eng, counts, p_int, args.c4_proteins, args.c4_queries)
    out["c5"] = run_c5(pkg, eng, counts, p_int, False, args.c5_ref)
    out["c5sat"] = run_c5(pkg, eng, counts, p_int, True, args.c5_ref)
    return out


def cuda_array(ptr, n, typestr):
    class _A:
        pass
    a = _A()
    a.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}
    return a


# ---------------------------------------------------------------------------------------------------------
# multi-GPU shapes that are NOT embarrassingly parallel: C4 database partition, C5 reference split
# ---------------------------------------------------------------------------------------------------------
def run_c4_sharded(pkg, sharding, eng, args, rank, world, dist, torch):
    """C4 over N GPUs (mpi_sw_solve_uniprot.cpp:65-72 gives every rank a block of database files): the database is
    partitioned by residues (length-sorted greedy balance), every rank searches its partition with every query, one
    all-gather per query puts (score, pos) back into database order.  Strong scaling: the job is fixed."""
    o = oracle() if rank == 0 else None
    qs = pkg.synth.c4_queries(args.c4_queries, 300)
    db = pkg.synth.c4_database(args.c4_proteins)
    lens = np.array([len(p) for p in db], dtype=np.int64)
    parts = sharding.balanced_partition(lens, world)
    mine = parts[rank].tolist()
    table = pkg.synth.blosum62_table()
    eng.set_scoring_table(pkg.MODE_EXACT, table, 10)
    eng.set_reference(qs[0])
    eng.stage([db[i] for i in mine], consensus=False)
    n_mine = len(mine)
    gather = sharding.make_index_gather(parts, torch.device("cuda", torch.cuda.current_device())) if world > 1 else None

    def one_query(q):
        eng.rebind_reference(q)
        us = eng.run()
        ds, dp = eng.device_results()
        s = torch.as_tensor(cuda_array(ds, n_mine, "<i4"), device="cuda")
        p = torch.as_tensor(cuda_array(dp, n_mine, "<u4"), device="cuda").view(torch.int32)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if world > 1:
            s_all, p_all = gather(s, p)
        else:
            s_all, p_all = s, p
        e1.record(); e1.synchronize()
        return us, e0.elapsed_time(e1) * 1e3, s_all, p_all

    one_query(qs[0])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    tot, coll = 0.0, 0.0
    for q in qs:
        us, cus, s_all, p_all = one_query(q)
        tot += us + cus; coll += cus
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    vec = torch.tensor([tot, coll], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
    tot, coll = [float(v) for v in vec.tolist()]
    residues = [int(lens[p.numpy()].sum()) for p in parts]
    ok, nchk = None, 0
    if rank == 0:
        ok = True
        sa, pa = s_all.cpu().numpy(), p_all.cpu().numpy()
        order = np.arange(len(db)) if world > 1 else np.array(mine)
        rng = np.random.default_rng(5)
        for k in rng.integers(0, len(order), size=48):
            i = int(order[k]) if world == 1 else int(k)
            w = o.align(db[i], qs[-1], mode=o.MODE_EXACT, table=table, gap=10)
            got = (int(sa[k if world == 1 else i]), int(pa[k if world == 1 else i]))
            ok = ok and (got == (w["score"], w["pos"]) if w["score"] > 0 else got[0] == 0)
            nchk += 1
    cells = int(lens.sum()) * 300 * len(qs)
    return {"gcups": cells / tot / 1e3, "ms": tot / 1e3, "collective_ms": coll / 1e3, "collective": "one all_gather of the packed (score, pos) words per query, database order restored by one index_select",
            "proteins": len(db), "queries": len(qs), "residues_per_rank": residues, "load_imbalance_max_over_mean": max(residues) / (sum(residues) / world),
            "parity_sample": {"n": nchk, "ok": ok, "against": "oracle/sw_oracle.c"}}, cells, tot


def run_c5_sharded(pkg, sharding, eng, args, rank, world, dist, torch, sat=False):
    """C5 over N GPUs (plocalaligner.cpp:44-67,106-143): the REFERENCE is cut into N overlapping ranges with the
    reference's own _make_string_range rule (halo = floor(ratio * m) columns), every rank aligns all reads against its
    range, one all-reduce(max) on score * N + (N - 1 - rank) picks the lowest-index range with the strictly greatest
    score, the winner contributes the global position with one all-reduce(sum) — the result of the serial
    OMPParallelLocalAligner(x, y, npiece = N, ratio).  Strong scaling.  The in-run check compares a 10 kbp x 240 kbp
    instance against the oracle's chunked aligner."""
    ratio = 2.0
    lref = pkg.synth.c5_reference(args.c5_ref)
    lreads = pkg.synth.c5_reads(lref, 16, 10_000)
    mode = pkg.MODE_SAT_U8 if sat else pkg.MODE_EXACT

    def staged(ref_full, reads):
        rng_ = pkg.make_string_range(world, len(reads[0]), len(ref_full), ratio)
        left, right = rng_[rank]
        eng.set_scoring_match(mode, 3, -3, 2)
        eng.set_reference(ref_full[left:right])
        eng.stage(reads, consensus=True, cons_stride=25_000)
        return left

    def step(left, n_reads):
        us = eng.run()
        ds, dp = eng.device_results()
        s = torch.as_tensor(cuda_array(ds, n_reads, "<i4"), device="cuda").to(torch.int64)
        p = torch.as_tensor(cuda_array(dp, n_reads, "<u4"), device="cuda").view(torch.int32).to(torch.int64)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        packed = s * world + (world - 1 - rank)
        if world > 1:
            dist.all_reduce(packed, op=dist.ReduceOp.MAX)
        winner = (world - 1) - (packed % world)
        mine = winner == rank
        out = torch.stack([torch.where(mine, s, torch.zeros_like(s)), torch.where(mine, p + left, torch.zeros_like(p))])
        if world > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM)
        e1.record(); e1.synchronize()
        return us, e0.elapsed_time(e1) * 1e3, out, winner

    # in-run oracle check on a small instance of the same shape
    ok = None
    sref = pkg.synth.c5_reference(240_000, seed=31)
    sreads = pkg.synth.mutated_reads(sref, 4, 10_000, seed=32, sub=0.02, ins=0.002, dele=0.002)
    left = staged(sref, sreads)
    _, _, out, winner = step(left, len(sreads))
    if rank == 0:
        o = oracle()
        ok = True
        for i, x in enumerate(sreads):
            w = o.align_chunked(x, sref, world, ratio, mode=o.MODE_SAT_U8 if sat else o.MODE_EXACT) if world > 1 else o.align(x, sref, mode=o.MODE_SAT_U8 if sat else o.MODE_EXACT, linear=True)
            ok = ok and (int(out[0][i]), int(out[1][i])) == (w["score"], w["pos"])
    left = staged(lref, lreads)
    step(left, len(lreads))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    tot, coll = 0.0, 0.0
    reps = 2
    for _ in range(reps):
        us, cus, out, winner = step(left, len(lreads))
        tot += us + cus; coll += cus
    vec = torch.tensor([tot / reps, coll / reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
    tot, coll = [float(v) for v in vec.tolist()]
    cells = sum(len(x) for x in lreads) * len(lref)
    return {"gcups": cells / tot / 1e3, "ms": tot / 1e3, "collective_ms": coll / 1e3, "collective": "all_reduce(max) on score*N+(N-1-rank), then all_reduce(sum) of the winner's (score, pos)",
            "reads": len(lreads), "ref_len": len(lref), "halo_columns": int(10_000 * ratio), "mode": "SAT_U8" if sat else "EXACT",
            "winner_ranks": [int(v) for v in winner.tolist()], "parity_sample": {"n": len(sreads), "ok": ok, "against": "oracle align_chunked(x, y, N, 2.0) on 10 kbp x 240 kbp"}}, cells, tot


# ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=["c3", "c1", "c1x64", "c2", "c2x64", "c2p17x64", "c4", "c5", "c5sat"])
    ap.add_argument("--reads", type=int, default=151552, help="C3: reads per GPU per step")
    ap.add_argument("--cpu-reads", type=int, default=0, help="reads in the CPU-baseline sample (default 2 per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block of the default line")
    ap.add_argument("--c4-proteins", type=int, default=500_000)
    ap.add_argument("--c4-queries", type=int, default=64)
    ap.add_argument("--c5-ref", type=int, default=51_000_000)
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
    counts, counts_src = load_sass_counts()
    eng = pkg.Engine(local_rank)

    if args.config != "c3":
        rc = other_config_line(args, pkg, sharding, eng, rank, world, local_rank, dist, torch, counts, p_int, p_src)
        eng.close()
        if use_dist:
            dist.barrier()
            dist.destroy_process_group()
        return rc

    # ---- inputs: rank-private shard of reads (weak scaling), replicated reference -------------------------
    ref, reads = make_batch(pkg, args.reads, seed=2300 + rank)
    n_reads = reads.shape[0]
    blob = torch.from_numpy(reads.reshape(-1).copy()).pin_memory()
    offs = torch.arange(n_reads + 1, dtype=torch.int64) * READ_LEN
    offs_np = offs.numpy().astype(np.uint64)
    blob_np = blob.numpy()
    cells_step = n_reads * READ_LEN * REF_LEN
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

    # ---- parity of the TIMED batch: a sample of its reads against the oracle, outside the timed region --------
    parity = None
    if rank == 0:
        o = oracle()
        idx = np.linspace(0, n_reads - 1, 16).astype(int)
        ok = True
        for i in idx:
            w = o.align(reads[i].tobytes().decode("ascii"), ref, mode=o.MODE_SAT_U8, linear=True)
            ln = int(res["len"][i])
            got = (int(res["score"][i]), int(res["pos"][i]), tuple(int(v) for v in res["end"][i]), res["cx_raw"][i, :ln].tobytes().decode("latin-1"), res["cy_raw"][i, :ln].tobytes().decode("latin-1"))
            ok = ok and got == (w["score"], w["pos"], tuple(w["end"]), w["cx"], w["cy"])
        parity = {"n": len(idx), "ok": bool(ok), "against": "oracle/sw_oracle_linear.c (score, pos, arg-max cell, both consensus strings) on reads of the timed batch"}
        if not ok:
            print(json.dumps({"error": "parity_sample failed on the timed batch", "parity_sample": parity}), flush=True)
            return 1

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
    # where the end-to-end time goes: the same call in its three parts (three extra iterations, not part of E; medians)
    parts = []
    for _ in range(3):
        t0 = time.perf_counter()
        eng.stage((blob_np, offs_np), consensus=True, cons_stride=cons_stride)
        t1 = time.perf_counter()
        eng.run()
        t2 = time.perf_counter()
        eng.fetch()
        t3 = time.perf_counter()
        parts.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
    med = [float(np.median([p_[k] for p_ in parts])) for k in range(3)]
    e2e_parts = {"stage_ms": med[0], "run_ms": med[1], "fetch_ms": med[2], "stage_ms_all": [round(p_[0], 2) for p_ in parts],
                 "step_ms_all": [round(t * 1e3, 1) for t in e2e_t],
                 "note": "swb_batch_stage (host preparation + H2D) / swb_batch_run (kernels) / swb_batch_fetch (D2H into host arrays); medians of 3"}

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
        k_gcups = cells_step / (pass1_us * 1e-6) / 1e9          # dominant kernel (score pass), this GPU
        contract = k_gcups * 1e9 * ops / 1e12                    # algorithmic Tops/s (9-op contract figure)
        st1 = dict(st, pass1_us=pass1_us)
        fr = issue_fraction(counts, st1, True, p_int)
        hbm_bytes = h2d + st["cells_executed"] / (2 * st["rows_per_lane"] * st["block_steps"]) * (st["rows_per_lane"] + 2) * 4  # checkpoints + block maxima
        roof = {"bound": "alu", "unit": "T lane-ops/s (ALU pipe)", "peak": p_int, "peak_source": p_src,
                "achieved": fr["frac"] * p_int if fr else None, "frac": fr["frac"] if fr else None,
                "frac_definition": "ALU-pipe lane instructions per second of the score kernel / measured ALU-pipe peak; instructions per cell pair counted from the built SASS",
                "alu_inst_per_cell_pair": fr["alu_inst_per_cell_pair"] if fr else None, "inst_per_cell_pair": fr["inst_per_cell_pair"] if fr else None,
                "sass_counts": counts_src, "sass_kernel": fr["kernel"] if fr else None,
                "frac_contract_9ops": contract / p_int, "achieved_contract_9ops": contract, "ops_per_cell": ops,
                "traffic": None, "traffic_note": "not measured in this run; an ncu capture of this kernel is summarised in profiles/score_kernel_r02_ncu.txt",
                "kernel": "score_kernel (pass 1)", "kernel_gcups": k_gcups, "kernel_ms": pass1_us / 1e3, "kernel_share_of_step": pass1_us / step_us,
                "executed_cell_fraction": cells_step / max(1, st["cells_executed"]),
                "hbm": {"achieved_gbs": hbm_bytes / (pass1_us * 1e-6) / 1e9, "peak_gbs": hbm_peak, "peak_source": hbm_src, "note": "algorithmic bytes (inputs + checkpoints) / kernel time: far from binding"}}
        line = {
            "metric": METRIC, "value": gcups, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_us / 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "s16x2 (u8-saturating semantics)", "data": "synthetic",
            "config": {"workload": f"C3 batched read mapping: {READ_LEN} bp reads (1% sub, 0.1% ins, 0.1% del) vs {REF_LEN} bp synthetic reference, SAT_U8 (+3/-3, gap 2), score+argmax+traceback",
                       "reads_per_gpu_per_step": n_reads, "cells_per_step_per_gpu": cells_step, "parallelism": f"reads sharded over {world} GPU(s), reference replicated, one NCCL all-gather of (score,pos)",
                       "l2": "flushed (256 MiB write) between timed iterations; per-step work buffers (checkpoints) exceed L2",
                       "geometry": {"lanes_per_pair": st["lanes_per_pair"], "rows_per_lane": st["rows_per_lane"], "block_steps": st["block_steps"]}},
            "e2e": {"value": e2e_gcups, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3, "parts": e2e_parts},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "parity_sample": parity,
            "pass2_ms": float(np.mean(p2_us)) / 1e3, "pass2_share": float(np.mean(p2_us)) / step_us, "wall_ms_per_step": t_wall / args.steps * 1e3,
        }
        if world == 1 and not args.no_configs:
            del flush
            torch.cuda.empty_cache()
            line["configs"] = all_configs(pkg, eng, counts, p_int, args)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(pkg, args)
            line["cpu_baselines"] = cpu_baselines(pkg)
        print(json.dumps(line), flush=True)
    eng.close()
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def other_config_line(args, pkg, sharding, eng, rank, world, local_rank, dist, torch, counts, p_int, p_src):
    """--config X: the same JSON line with config X as the headline.  c1*/c2*: every rank runs the same batch (weak
    scaling, reads sharded like C3); c4 / c5: the job is fixed and sharded over the ranks (strong scaling)."""
    sampler = ClockSampler(local_rank)
    sampler.start()
    scaling = "weak"
    if args.config in ("c4",) :
        scaling = "strong"
        e, cells, us = run_c4_sharded(pkg, sharding, eng, args, rank, world, dist, torch)
        value = cells / us / 1e3
    elif args.config in ("c5", "c5sat"):
        scaling = "strong"
        e, cells, us = run_c5_sharded(pkg, sharding, eng, args, rank, world, dist, torch, sat=args.config == "c5sat")
        value = cells / us / 1e3
    else:
        e = run_c1_c2(pkg, eng, counts, p_int, args.config, reps=max(3, args.steps))
        t = torch.tensor([e["device_ms"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        value = e["gcups"] * e["device_ms"] / float(t[0]) * world
        us = float(t[0]) * 1e3
    clocks = sampler.stop()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": us / 1e3,
                "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "s16x2", "data": "synthetic" if args.config[:2] in ("c4", "c5") else "data/data_small (shipped)",
                "config": {"workload": args.config, "detail": e}, "clocks": clocks, "gpu_launches": None, "parity_sample": e.get("parity_sample")}
        print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
