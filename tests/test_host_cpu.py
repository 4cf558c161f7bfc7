"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol include/swb200.h declares,
host-only entry points behave like the reference, the product refuses to run without a CUDA device,
and the N>1 sharding + gather logic works over gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import pyoracle as o
from conftest import ROOT, load_pkg


def test_library_exports_every_declared_symbol():
    pkg = load_pkg()
    lib = pkg.load_library()
    with open(os.path.join(ROOT, "include", "swb200.h")) as f:
        hdr = f.read()
    declared = sorted(set(re.findall(r"\b(swb_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(pkg.EXPORTS) == declared
    assert b"sm_100a" in lib.swb_version()


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    pkg = load_pkg()
    with pytest.raises(pkg.SwbError) as ei:
        pkg.Engine(0)
    assert ei.value.code == -1


def test_product_never_touches_the_oracle():
    """The product package must not import, load or name anything under oracle/."""
    pdir = os.path.join(ROOT, "parallel-genomeseq_b200")
    for root, _, files in os.walk(pdir):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                with open(os.path.join(root, fn), errors="ignore") as f:
                    txt = f.read()
                assert "pyoracle" not in txt and "sw_oracle" not in txt and "liboracle" not in txt and "libref_aligner" not in txt, fn


def test_make_string_range_matches_oracle_and_reference_values():
    """swb_make_string_range == _make_string_range (plocalaligner.cpp:44-67), incl. the assert preconditions."""
    pkg = load_pkg()
    assert pkg.make_string_range(4, 125, 4980, 2.0) == [(0, 1432), (1182, 2614), (2364, 3796), (3546, 4980)]
    rng = np.random.default_rng(0)
    for _ in range(300):
        npiece = int(rng.integers(1, 20)); m = int(rng.integers(1, 400)); n = int(rng.integers(1, 20000)); ratio = float(rng.choice([0.5, 1.0, 1.5, 2.0, 2.5]))
        want = o.make_string_range(npiece, m, n, ratio)
        if isinstance(want, int):
            with pytest.raises(pkg.SwbError):
                pkg.make_string_range(npiece, m, n, ratio)
        else:
            assert pkg.make_string_range(npiece, m, n, ratio) == want


def test_pack_sequences():
    pkg = load_pkg()
    blob, offs = pkg.pack_sequences(["ACG", "T", "GGCA"])
    assert blob.tobytes() == b"ACGTGGCA" and offs.tolist() == [0, 3, 4, 8]
    arr = np.frombuffer(b"ACGTACGT", dtype=np.uint8).reshape(2, 4)
    blob, offs = pkg.pack_sequences(arr)
    assert offs.tolist() == [0, 4, 8]


def test_block_partition_like_mpi_driver():
    pkg = load_pkg()
    import importlib
    sh = importlib.import_module("parallel-genomeseq_b200.sharding")
    assert sh.block_partition(1170, 8) == [(i * 146, (i + 1) * 146) for i in range(7)] + [(1022, 1170)]
    assert sh.block_partition(10, 1) == [(0, 10)]
    parts = sh.block_partition(1_000_000, 8)
    assert parts[0] == (0, 125000) and parts[-1] == (875000, 1000000)


def test_synth_workloads_are_seeded():
    pkg = load_pkg()
    a, b = pkg.synth.c3_reference(5000), pkg.synth.c3_reference(5000)
    assert a == b and set(a) <= set("ACGT")
    r = pkg.synth.c3_reads(a, 5, read_len=150)
    assert all(len(x) == 150 for x in r)
    t = pkg.synth.blosum62_table()
    assert t[ord("W"), ord("W")] == 11 and t[ord("A"), ord("R")] == -1 and (t == t.T).all()


_WORKER = r'''
import importlib, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=rank, world_size=world)
sh = importlib.import_module("parallel-genomeseq_b200.sharding")
n = 1171
parts = sh.block_partition(n, world)
lo, hi = parts[rank]
score = torch.arange(lo, hi, dtype=torch.int32) * 3 + 1       # stand-in for this rank's kernel results
pos = torch.arange(lo, hi, dtype=torch.int32) + 7
s_all, p_all = sh.gather_score_pos(score, pos, counts=[b - a for a, b in parts])
assert s_all.tolist() == [3 * i + 1 for i in range(n)], rank
assert p_all.tolist() == [i + 7 for i in range(n)], rank
e_s, e_p = sh.gather_score_pos(score[:500].contiguous(), pos[:500].contiguous())
assert e_s.numel() == 500 * world and e_s[500].item() == 3 * parts[1][0] + 1
# database search: entries balanced by residues, results gathered back into entry order
lens = np.random.default_rng(5).integers(10, 5000, size=333)
parts = sh.balanced_partition(lens, world)
assert sorted(torch.cat(parts).tolist()) == list(range(333))
loads = [int(lens[p.numpy()].sum()) for p in parts]
assert max(loads) - min(loads) <= int(lens.max())
mine = parts[rank]
vals = sh.gather_by_index((mine * 5 + 2).to(torch.int32), parts)
assert vals.tolist() == [5 * i + 2 for i in range(333)], rank
gather = sh.make_index_gather(parts, torch.device("cpu"))       # the per-query gather of a database search: built once
for rnd in range(3):
    s_g, p_g = gather((mine * 5 + rnd).to(torch.int32), (mine * 7 + 3).to(torch.int32))
    assert s_g.tolist() == [5 * i + rnd for i in range(333)] and p_g.tolist() == [7 * i + 3 for i in range(333)], rank
# long pairs: the reference split over the ranks, one all-reduce(max) picks the piece like the serial
# OMPParallelLocalAligner(x, y, npiece = world, ratio) does (oracle as the stand-in aligner)
sys.path.insert(0, os.path.join(sys.argv[1], "oracle"))
import pyoracle as o
pkg = importlib.import_module("parallel-genomeseq_b200")
rng = np.random.default_rng(11)
y = "".join(rng.choice(list("ACGT"), size=3000))
reads = []
for k in range(12):
    m = 150 if k < 8 else int(rng.integers(60, 200))
    s0 = int(rng.integers(0, len(y) - m))
    x = list(y[s0:s0 + m])
    for q in range(m):
        if rng.random() < 0.05:
            x[q] = str(rng.choice(list("ACGT")))
    reads.append("".join(x))
reads.append("".join(rng.choice(list("ACGT"), size=90)))
for (ma, mi, g) in ((3, -3, 2), (2, -4, 3)):
    def piece_aligner(match, mismatch, gap):
        def f(rs, yp):
            w = [o.align(r, yp, mode=o.MODE_EXACT, match=match, mismatch=mismatch, gap=gap) for r in rs]
            return [v["score"] for v in w], [v["pos"] for v in w]
        return f
    sc, ps, win = sh.reference_sharded_align(piece_aligner(ma, mi, g), reads, y, 2.0, pkg.make_string_range,
                                             realign=None if (ma, mi, g) == (3, -3, 2) else piece_aligner(3, -3, 2))
    for i, x in enumerate(reads):
        w = o.align_chunked(x, y, world, 2.0, mode=o.MODE_EXACT, match=ma, mismatch=mi, gap=g)
        assert (int(sc[i]), int(ps[i]), int(win[i])) == (w["score"], w["pos"], w["piece"]), (rank, i, (ma, mi, g), int(sc[i]), int(ps[i]), int(win[i]), w["score"], w["pos"], w["piece"])
t = torch.tensor([float(rank + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)                       # bench.py: max-over-ranks time
assert t.item() == world
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_sharding_and_gather_world_size_2_gloo(tmp_path):
    """The N>1 paths on CPU, world_size 2, gloo: block partition + ragged all-gather + max-over-ranks (bench.py),
    residue-balanced database partition + gather by index (config 4), and the reference split over ranks with
    an all-reduce(max) piece selection equal to the serial OMPParallelLocalAligner (config 5, oracle as aligner)."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + (os.getpid() % 2000))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT, port], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, out
        assert f"rank {r} ok" in out


def test_bench_reference_arm_runs_on_cpu():
    """bench.py --impl reference times the reference's own CPU implementation (oracle/_ref, else the oracle port)
    and prints one JSON line with the contract keys — no GPU needed."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-reads", "2"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().split("\n")[-1])
    assert line["impl"] == "reference" and line["unit"] == "GCUPS" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True


def test_bench_peaks_and_metric_contract():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    p_int, src, hbm, hsrc = b.load_peaks()
    assert 10.0 < p_int < 40.0 and hbm > 1000
    import json
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        assert b.METRIC == json.load(f)["metric"]
    assert b.OPS_PER_CELL["SAT_U8"] == 9 and b.OPS_PER_CELL["EXACT"] == 8


def test_issue_fraction_uses_executed_counts_for_the_pipelined_strip_kernel():
    """The SASS loop of score_units_kernel holds polling and slow paths that are not taken in the steady state, so its
    roofline fraction must come from the EXECUTED instruction counts (ncu capture, tagged from_profile) — and equal the
    ALU-pipe utilisation that capture measured; every other kernel is counted from the SASS of the build."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    counts, src = b.load_sass_counts()
    assert src and counts
    p_int = 18.5238
    # the geometry and timing of the ncu capture at 10 kbp x 51 Mbp (profiles/score_units_kernel_r02_c5_ncu.txt (2))
    cells = 79 * 128 * 51_003_392 * 2 * 8
    st = {"kernel_kind": 1, "rows_per_lane": 4, "cols_per_step": 8, "cells_executed": cells, "pass1_us": 2.019096e6}
    fr = b.issue_fraction(counts, st, False, p_int)
    assert fr["counts_from"].startswith("from_profile") and fr["alu_inst_per_cell_pair"] < fr["alu_inst_per_cell_pair_static_sass"]
    assert abs(fr["frac"] - 0.364) < 0.01, fr            # ncu: ALU pipe 36.4 % busy
    st3 = {"kernel_kind": 0, "rows_per_lane": 19, "cols_per_step": 1, "cells_executed": 10 ** 13, "pass1_us": 1.0e6}
    fr3 = b.issue_fraction(counts, st3, True, p_int)
    assert fr3["counts_from"].startswith("SASS") and 3.5 < fr3["alu_inst_per_cell_pair"] < 4.0


def test_sass_counts_match_the_build():
    """profiles/sass_counts_r02.json (the instruction counts bench.py's issue-based roofline fraction uses) is what
    tools/sass_counts.py computes from the objects of the current build."""
    import shutil
    if not shutil.which("cuobjdump") or not os.path.isfile(os.path.join(ROOT, "parallel-genomeseq_b200", "build", "sw_inst_r19.o")):
        pytest.skip("needs cuobjdump and the build objects (build container)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_counts.py"), "--check"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr


def test_reference_headers_mode_compiles():
    """cuda_aligner.h with -DSWB_WITH_REFERENCE_HEADERS against the reference's own headers, through the reference's
    unit test file (tests/cpp/ref_test_swap.cpp); running it needs a GPU (tests/test_gpu_parity.py)."""
    if not os.path.isdir("/root/reference"):
        pytest.skip("needs /root/reference (build container)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "cpp", "build_ref_tests.py"), "--force"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and os.path.isfile(os.path.join(ROOT, "tests", "cpp", "ref_test_localaligner")), r.stdout + r.stderr


def test_dataprep_tools(tmp_path):
    """the dependency-free data preparation — SAM -> ground-truth CSV, FASTQ -> one read per line, custom
    reference / read generation, multi-FASTA splitting and the packed 5-bit database blob.  In the build container the
    SAM the reference ships must reproduce the reference's own data_small_ground_truth.csv byte for byte."""
    import importlib
    dp = importlib.import_module("parallel-genomeseq_b200.dataprep")
    sam = tmp_path / "t.sam"
    sam.write_text("@HD\tVN:1.0\n@SQ\tSN:22_5K\tLN:4980\nr1\t99\t22_5K\t17\t60\t5M\t=\t40\t28\tACGTA\tIIIII\tNM:i:0\nr2\t147\t22_5K\t40\t60\t4M\t=\t17\t-28\tGGCC\tIIII\n")
    assert dp.sam_to_ground_truth(str(sam), str(tmp_path / "gt.csv")) == 2
    assert (tmp_path / "gt.csv").read_text() == "index,QNAME,SEQ,POS\n0,r1,ACGTA,17\n1,r2,GGCC,40\n"
    ref_sam = "/root/reference/data/data_small/output_tiny_30xCov.mod.sam"
    if os.path.isfile(ref_sam):
        dp.sam_to_ground_truth(ref_sam, str(tmp_path / "gt_ref.csv"))
        with open("/root/reference/data/data_small_ground_truth.csv") as f:
            assert (tmp_path / "gt_ref.csv").read_text() == f.read()
        with open(os.path.join(ROOT, "tests", "golden", "data_small", "data_small_ground_truth.csv")) as f:
            assert (tmp_path / "gt_ref.csv").read_text() == f.read()
    fq = tmp_path / "t.fq"
    fq.write_text("@a\nACGT\n+\nIIII\n@b\nTTGA\n+\nIIII\n")
    assert dp.fastq_sequences(str(fq), str(tmp_path / "lines.txt")) == ["ACGT", "TTGA"] and (tmp_path / "lines.txt").read_text() == "ACGT\nTTGA\n"
    fa = tmp_path / "long.fa"
    fa.write_text(">chr\n" + "acgtn" * 4 + "\n" + "GGCCN" * 4 + "\n" + "TTTTT" * 4 + "\n")
    assert dp.read_fa(str(fa)) == "acgtn" * 4 + "GGCCN" * 4 + "TTTTT" * 4
    assert dp.gen_ref_custom(str(fa), str(tmp_path / "ref.fa"), start_pos=20, ref_len=20) == "GGCC" * 4      # the line that starts in the window, N removed
    rows = dp.gen_reads_custom(str(tmp_path / "ref.fa"), str(tmp_path / "reads.csv"), read_len=5, n_reads=7, seed=3)
    lines = (tmp_path / "reads.csv").read_text().strip().split("\n")
    assert lines[0] == "index,QNAME,SEQ,POS" and len(lines) == 8 and all(("GGCC" * 4)[p:p + 5] == s for _, s, p in rows)
    assert (tmp_path / "reads_readsonly.txt").read_text().count("\n") == 7
    mf = tmp_path / "db.fasta"
    prots = ["MKV", "ARNDCQEGHILKMFPSTWYVBZX", "MKTAYIAKQRQISFVKSHFSRQ", "uoj", "A"]
    mf.write_text("".join(f">sp|P{i}|x\n{p[:10]}\n{p[10:]}\n" for i, p in enumerate(prots)))
    assert dp.split_multifasta(str(mf), str(tmp_path / "database.fasta"), str(tmp_path / "stats.txt")) == 5
    assert (tmp_path / "stats.txt").read_text() == "5" and (tmp_path / "database.fasta").read_text().split("\n")[1] == prots[1]
    assert dp.pack_database(str(mf), str(tmp_path / "db.swbdb")) == (5, sum(len(p) for p in prots))
    blob, offs, orig = dp.load_database(str(tmp_path / "db.swbdb"))
    got = [blob[int(offs[i]):int(offs[i + 1])].tobytes().decode() for i in range(5)]
    assert [len(g) for g in got] == sorted((len(p) for p in prots), reverse=True)
    assert {int(o_): g for o_, g in zip(orig, got)} == {i: p.upper() for i, p in enumerate(prots)}
    assert os.path.getsize(tmp_path / "db.swbdb") < 8 + 20 + 32 + 4 * 5 + 8 * 6 + sum(len(p) for p in prots)      # 5 bits per residue


def test_packed_database_cpp_and_python_agree(tmp_path):
    """The packed 5-bit database written by the C++ driver (sw_search_uniprot --pack, cpp/packed_db.h) is byte-identical to
    the one parallel-genomeseq_b200/dataprep.py writes (no GPU involved)."""
    import importlib
    dp = importlib.import_module("parallel-genomeseq_b200.dataprep")
    drv = os.path.join(ROOT, "parallel-genomeseq_b200", "drivers", "sw_search_uniprot")
    if not os.path.isfile(drv):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(drv)])
    synth = importlib.import_module("parallel-genomeseq_b200.synth")
    prots = synth.c4_database(300, seed=3) + ["MKV", "x*u", "A"]
    mf = tmp_path / "db.fasta"
    mf.write_text("".join(f">sp|P{i}|x\n" + "\n".join(p[k:k + 60] for k in range(0, len(p), 60)) + "\n" for i, p in enumerate(prots)))
    r = subprocess.run([drv, "--pack", str(mf), str(tmp_path / "cpp.swbdb")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    dp.pack_database(str(mf), str(tmp_path / "py.swbdb"))
    assert (tmp_path / "cpp.swbdb").read_bytes() == (tmp_path / "py.swbdb").read_bytes()
    blob, offs, orig = dp.load_database(str(tmp_path / "cpp.swbdb"))
    assert blob[int(offs[0]):int(offs[1])].tobytes().decode() == max(prots, key=len).upper()
