#!/usr/bin/env python3
"""Instruction mix of the interior loop of every pass-1 kernel the BASELINE configs run, counted from the SASS of the
built objects (cuobjdump) — the basis of the ISSUE-BASED roofline fraction bench.py reports:

    frac = (ALU-pipe instructions per cell pair in the hot loop) x (executed cell pairs / s) / P_int

A lane instruction on s16x2 operands updates a cell PAIR; P_int is the measured ALU-pipe issue rate in lane-ops/s
(profiles/alu_peak_r01.json: 64 lanes/clk/SM for IADD3 / LOP3 / VIMNMX / VIADDMNMX alike).  IMAD goes to the FMA pipe and
LDS / LDG / STG / SHFL / BRA to other units; they are listed but not counted.

    python tools/sass_counts.py            -> writes profiles/sass_counts_r02.json
    python tools/sass_counts.py --check    -> exit 1 if the committed file differs from the current build
"""
import json
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "parallel-genomeseq_b200", "build")
OUT = os.path.join(ROOT, "profiles", "sass_counts_r02.json")

# pipes by mnemonic (B300_MICROARCH.md / the alu_peak microbenchmark): everything here issues on the ALU pipe
ALU = {"VIADDMNMX", "VIMNMX3", "VIMNMX", "VIADD", "IADD3", "IADD", "LOP3", "ISETP", "SEL", "PRMT", "SHF", "LEA", "IMNMX", "HSET2", "HMNMX2",
       "MOV", "PLOP3", "P2R", "R2P", "VABSDIFF", "BMSK", "SGXT", "FLO", "POPC", "IABS", "CS2R"}
FMA = {"IMAD", "HFMA2", "FFMA", "FMUL", "FADD", "IDP"}

# key -> (object, mangled kernel, DP instructions that identify the loop, cell pairs per lane per loop trip)
#   loop trip = two wavefront steps of R rows x C columns
KERNELS = {
    "c3  score_kernel<R=19,C=1,SAT,profile>  (8 lanes x 19 rows)": ("sw_inst_r19.o", "_ZN3swb12score_kernelILi19ELi1ELi1ELb1EEEvNS_10PassParamsE", 2 * 19 * 1),
    "c1x64/c2x64  score_kernel<R=16,C=1,SAT,profile>  (8 lanes x 16 rows)": ("sw_inst_r16.o", "_ZN3swb12score_kernelILi16ELi1ELi1ELb1EEEvNS_10PassParamsE", 2 * 16 * 1),
    "c1/c2 x1  score_kernel<R=4,C=2,SAT,profile>  (32 lanes x 4 rows, 2 columns)": ("sw_inst_r4.o", "_ZN3swb12score_kernelILi4ELi2ELi1ELb1EEEvNS_10PassParamsE", 2 * 4 * 2),
    "c2 x1  score_kernel<R=4,C=1,SAT,profile>  (32 lanes x 4 rows)": ("sw_inst_r4.o", "_ZN3swb12score_kernelILi4ELi1ELi1ELb1EEEvNS_10PassParamsE", 2 * 4 * 1),
    "c4  qs_score_kernel<R=19,EXACT,paired 16-bit profile>  (16 lanes x 19 rows, query-stationary)": ("sw_inst_r19.o", "_ZN3swb15qs_score_kernelILi19ELb0ELb1EEEvNS_8QsParamsE", 2 * 19 * 1),
    "c5  score_units_kernel<R=5,C=8,EXACT,profile>  (pipelined strips, SWB_STRIP_R=5)": ("sw_inst_r5.o", "_ZN3swb18score_units_kernelILi5ELi8ELi0ELb1EEEvNS_10PassParamsE", 2 * 5 * 8),
    "c5  score_units_kernel<R=5,C=8,SAT,profile>  (pipelined strips, SWB_STRIP_R=5)": ("sw_inst_r5.o", "_ZN3swb18score_units_kernelILi5ELi8ELi1ELb1EEEvNS_10PassParamsE", 2 * 5 * 8),
    "c5  score_units_kernel<R=4,C=8,EXACT,profile>  (pipelined strips, 32 lanes x 4 rows, 8 columns)": ("sw_inst_r4.o", "_ZN3swb18score_units_kernelILi4ELi8ELi0ELb1EEEvNS_10PassParamsE", 2 * 4 * 8),
    "c5  score_units_kernel<R=4,C=8,SAT,profile>  (pipelined strips, 32 lanes x 4 rows, 8 columns)": ("sw_inst_r4.o", "_ZN3swb18score_units_kernelILi4ELi8ELi1ELb1EEEvNS_10PassParamsE", 2 * 4 * 8),
    "c5  score_units_kernel<R=8,C=8,EXACT,profile>  (pipelined strips, when the boundary rows of 4-row strips do not fit HBM)": ("sw_inst_r8.o", "_ZN3swb18score_units_kernelILi8ELi8ELi0ELb1EEEvNS_10PassParamsE", 2 * 8 * 8),
    "c5  score_units_kernel<R=8,C=8,SAT,profile>  (pipelined strips, when the boundary rows of 4-row strips do not fit HBM)": ("sw_inst_r8.o", "_ZN3swb18score_units_kernelILi8ELi8ELi1ELb1EEEvNS_10PassParamsE", 2 * 8 * 8),
}


# EXECUTED instructions per cell pair, from ncu counters, for kernels whose loop body holds paths that are not taken in the
# steady state (score_units_kernel: progress polling, NANOSLEEP, writer-lane and chunk hand-off code) — there the static
# count overstates the work: 8.9 instructions per cell pair in the SASS loop, 7.3 executed.
#   profiles/score_units_kernel_r02_c5_ncu.txt (2), 16 x 10 kbp x 51 Mbp EXACT, one launch: smsp__inst_executed.sum = 941 937 707 768,
#   ALU pipe 36.397 % of 0.5 warp-instructions/clk x 3 963 930 802 active cycles x 592 sub-partitions = 4.2706e11 instructions,
#   cell pairs per lane = 79 strips x 128 rows x 51 003 392 columns x 8 pairs / 32 lanes = 1.2894e11.
# The SAT_U8 variant has the same static ALU count (334 per trip) and was not captured: the EXACT figures are used for it.
EXECUTED = {
    "score_units_kernel<R=4,C=8,": dict(alu_inst_per_cell_pair=3.3121, inst_per_cell_pair=7.3054,
                                        source="ncu, profiles/score_units_kernel_r02_c5_ncu.txt (2): 10 kbp x 51 Mbp, EXACT"),
}


def registers(obj, kernel):
    out = subprocess.run(["cuobjdump", "--dump-resource-usage", os.path.join(OBJ, obj)], capture_output=True, text=True, check=True).stdout
    m = re.search(re.escape(kernel) + r":\s*\n\s*REG:(\d+)", out)
    return int(m.group(1)) if m else None


def loop_mix(obj, kernel, cell_pairs):
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", kernel, os.path.join(OBJ, obj)], capture_output=True, text=True, check=True).stdout
    ins = []
    lines = sass.split("\n")
    for n, line in enumerate(lines):
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)(.*?);", line)
        if m:
            # the second 64-bit word of the encoding carries the scheduler's stall count (bits 41..44 of that word)
            hi = re.search(r"/\* (0x[0-9a-f]{16}) \*/", lines[n + 1]) if n + 1 < len(lines) else None
            ins.append((int(m.group(1), 16), m.group(2), m.group(3), (int(hi.group(1), 16) >> 41) & 0xF if hi else 0))
    want = 2 * cell_pairs                       # at least two DPX instructions per cell pair (three in SAT_U8)
    best = None
    for a, op, rest, _ in ins:
        if op == "BRA":
            m = re.search(r"0x([0-9a-f]+)", rest)
            if m and int(m.group(1), 16) < a:
                body = [x for x in ins if int(m.group(1), 16) <= x[0] <= a]
                dp = sum(x[1] == "VIADDMNMX" for x in body)
                if dp >= want and (best is None or len(body) < len(best)):
                    best = body
    if best is None:
        raise SystemExit(f"no hot loop found in {kernel}")
    mix = Counter(x[1] for x in best)
    alu = sum(v for k, v in mix.items() if k in ALU)
    fma = sum(v for k, v in mix.items() if k in FMA)
    unknown = sorted(k for k in mix if k not in ALU and k not in FMA and k not in {"LDS", "LDG", "STG", "STS", "SHFL", "BRA", "LDC", "LDCU", "NANOSLEEP", "WARPSYNC", "BSSY", "BSYNC", "NOP",
                                                                                "UMOV", "UIADD3", "ULEA", "UISETP", "ULOP3", "USEL", "UIMAD", "USHF", "R2UR", "S2R", "MEMBAR", "ATOMG", "LD", "ST", "CCTL", "ERRBAR", "VOTE", "UFLO", "UPOPC", "REDUX", "S2UR", "UPRMT", "UMNMX", "BAR", "YIELD", "CALL", "RET", "EXIT", "BMOV", "DEPBAR"})
    return dict(object=obj, kernel=kernel, registers=registers(obj, kernel), loop_instructions=len(best), scheduled_stall_cycles=sum(x[3] for x in best),
                cell_pairs_per_lane_per_trip=cell_pairs,
                alu_pipe_instructions=alu, fma_pipe_instructions=fma, alu_inst_per_cell_pair=round(alu / cell_pairs, 4),
                inst_per_cell_pair=round(len(best) / cell_pairs, 4), mix=dict(mix.most_common()), unclassified=unknown)


def main():
    doc = {"how": "tools/sass_counts.py: smallest backward-branch loop of each kernel holding >= 2 VIADDMNMX per cell pair (cuobjdump -sass of parallel-genomeseq_b200/build/*.o)",
           "kernels": {k: loop_mix(*v) for k, v in KERNELS.items()}}
    for k, v in doc["kernels"].items():
        for pat, ex in EXECUTED.items():
            if pat in k:
                v["executed"] = ex
    if "--check" in sys.argv:
        with open(OUT) as f:
            old = json.load(f)
        bad = [k for k in doc["kernels"] if old["kernels"].get(k, {}).get("alu_pipe_instructions") != doc["kernels"][k]["alu_pipe_instructions"]
               or old["kernels"].get(k, {}).get("loop_instructions") != doc["kernels"][k]["loop_instructions"]
               or old["kernels"].get(k, {}).get("registers") != doc["kernels"][k]["registers"]]
        # occupancy guard: the batched score kernels must keep 4 blocks of 128 threads per SM
        bad += [k + " (more than 128 registers)" for k, v in doc["kernels"].items() if " score_kernel<" in k and (v["registers"] or 999) > 128]
        if bad:
            print("stale:", bad)
            sys.exit(1)
        print("profiles/sass_counts_r02.json matches the current build")
        return
    with open(OUT, "w") as f:
        json.dump(doc, f, indent=1)
    for k, v in doc["kernels"].items():
        print(f"{k}\n    {v['registers']} registers, loop {v['loop_instructions']} instr ({v['scheduled_stall_cycles']} scheduled cycles), ALU pipe {v['alu_pipe_instructions']} = {v['alu_inst_per_cell_pair']} / cell pair, FMA pipe {v['fma_pipe_instructions']}, unclassified {v['unclassified']}")


if __name__ == "__main__":
    main()
