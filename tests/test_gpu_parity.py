"""GPU parity: the CUDA path, called through the C ABI (include/swb200.h), against
  (a) the golden vectors dumped from the compiled reference (tests/golden/), and
  (b) the CPU oracle (oracle/sw_oracle.c) on fresh seeded inputs.
Bit-exact on score, pos, arg-max cell and both consensus strings — integer work, no tolerance."""
import hashlib

import numpy as np
import pytest

import pyoracle as o
import synth
from conftest import read_golden_csv

pytestmark = pytest.mark.gpu


def _check(res, i, want, tag=""):
    got = (int(res["score"][i]), int(res["pos"][i]), res["cx"][i], res["cy"][i])
    exp = (want["score"], want["pos"], want["cx"], want["cy"])
    assert got == exp, (tag, i, got[:2], exp[:2])
    assert res["flags"][i] == 0


def test_known_answer(engine, pkg):
    """test/test_localaligner.cpp:8-27,53-59 through the CUDA path, both arithmetic modes."""
    for mode in (pkg.MODE_SAT_U8, pkg.MODE_EXACT):
        engine.set_scoring_match(mode, 3, -3, 2)
        engine.set_reference("TGTTACGG")
        r = engine.align(["GGTTGACTA"])
        assert (int(r["score"][0]), int(r["pos"][0]), r["cx"][0], r["cy"][0]) == (13, 2, "CAGTTG", "CA-TTG")
        assert tuple(r["end"][0]) == (7, 6)


def test_random_pairs_golden(engine, pkg, random_pairs):
    """414 seeded pairs x {Skewed, plain} x 5 scorings x {whole reference, chunked} from the reference itself."""
    for k, c in enumerate(random_pairs):
        sc = c["scoring"]
        mode = pkg.MODE_SAT_U8 if c["smt"] == 0 else pkg.MODE_EXACT
        engine.set_scoring_match(mode, sc["match"], sc["mismatch"], sc["gap"])
        engine.set_reference(c["y"])
        r = engine.align([c["x"]], npiece=c["npiece"], ratio=c["ratio"], cons_stride=len(c["x"]) + len(c["y"]) + 2)
        _check(r, 0, c, tag=(k, len(c["x"]), len(c["y"]), c["smt"], sc, c["npiece"]))


@pytest.mark.parametrize("name,mode,npiece", [("data_small_sw_skewed.csv", 0, 0), ("data_small_p4.csv", 0, 4),
                                              ("data_small_p17.csv", 0, 17), ("data_small_sw_float.csv", 1, 0)])
def test_data_small_golden(engine, pkg, data_small, name, mode, npiece):
    """BASELINE configs 1 and 2: all 1170 data_small reads in one batch, identical to the reference."""
    ref, truth = data_small
    gold = read_golden_csv(name)
    engine.set_scoring_match(mode, 3, -3, 2)
    engine.set_reference(ref)
    r = engine.align([t[2] for t in truth], npiece=npiece, ratio=2.0)
    for i, g in enumerate(gold):
        _check(r, i, g, tag=name)


def test_c3_sample_golden(engine, pkg, c3_sample):
    """BASELINE config 3 shape: 150 bp reads vs the seeded 1 Mbp reference, golden from the reference."""
    ref = synth.c3_reference(c3_sample["ref_len"])
    assert hashlib.sha256(ref.encode()).hexdigest() == c3_sample["ref_sha256"]
    engine.set_scoring_match(pkg.MODE_SAT_U8, 3, -3, 2)
    engine.set_reference(ref)
    r = engine.align([e["x"] for e in c3_sample["reads"]])
    for i, e in enumerate(c3_sample["reads"]):
        _check(r, i, e, tag="c3")


def test_c4_sample_golden(engine, pkg, c4_sample):
    """BASELINE config 4 shape: DB proteins (x) vs a 300-aa query (y), BLOSUM62 callback, gap 10."""
    engine.set_scoring_table(pkg.MODE_EXACT, synth.blosum62_table(), c4_sample["gap"])
    engine.set_reference(c4_sample["query"])
    ents = c4_sample["entries"]
    assert max(len(e["x"]) for e in ents) > 1024       # exercises row strips (sequences longer than one warp holds)
    r = engine.align([e["x"] for e in ents], cons_stride=6000)
    for i, e in enumerate(ents):
        _check(r, i, e, tag="c4")


def test_oracle_fresh_random(engine, pkg):
    """Fresh seeded batches with ragged lengths against the CPU oracle (both modes, custom scoring)."""
    rng = np.random.default_rng(99)
    for rep, (mode, omode) in enumerate(((pkg.MODE_SAT_U8, o.MODE_SAT_U8), (pkg.MODE_EXACT, o.MODE_EXACT))):
        n = 700 + 13 * rep
        y = "".join(rng.choice(list("ACGT"), size=n))
        xs = []
        for k in range(97):
            m = int(rng.integers(1, 260))
            if m == n:
                m += 1
            if k % 2:
                s = int(rng.integers(0, n - m)) if m < n else 0
                x = list(y[s:s + m])
                for q in range(len(x)):
                    if rng.random() < 0.1:
                        x[q] = str(rng.choice(list("ACGT")))
                xs.append("".join(x))
            else:
                xs.append("".join(rng.choice(list("ACGT"), size=m)))
        for (ma, mi, g) in ((3, -3, 2), (2, -5, 1), (9, -1, 4)):
            engine.set_scoring_match(mode, ma, mi, g)
            engine.set_reference(y)
            r = engine.align(xs, cons_stride=n + 300)
            for i, x in enumerate(xs):
                w = o.align(x, y, mode=omode, match=ma, mismatch=mi, gap=g)
                if w["score"] == 0:
                    assert int(r["score"][i]) == 0 and int(r["len"][i]) == 0  # documented: reference UB (F10)
                    continue
                _check(r, i, w, tag=("fresh", mode, ma, mi, g, len(x)))
                assert tuple(r["end"][i]) == w["end"]


def test_oracle_read_longer_than_reference(engine, pkg):
    """len(x) > len(y): the skewed layout's other orientation (similaritymatrix.cpp:341-344,481-517)."""
    rng = np.random.default_rng(5)
    for mode, omode in ((pkg.MODE_SAT_U8, o.MODE_SAT_U8), (pkg.MODE_EXACT, o.MODE_EXACT)):
        y = "".join(rng.choice(list("ACGT"), size=61))
        xs = ["".join(rng.choice(list("ACGT"), size=int(m))) for m in (62, 90, 128, 200, 333)]
        engine.set_scoring_match(mode, 3, -3, 2)
        engine.set_reference(y)
        r = engine.align(xs, cons_stride=500)
        for i, x in enumerate(xs):
            _check(r, i, o.align(x, y, mode=omode), tag=("long-x", mode, len(x)))


def test_saturation_ties(engine, pkg):
    """Many cells tie at 255: the winner must be the first in the reference's skewed raw order."""
    rng = np.random.default_rng(11)
    y = "".join(rng.choice(list("ACGT"), size=1500))
    xs = [y[s:s + 140] for s in (0, 1, 100, 700, 1359, 1360)] + ["A" * 120, y[-140:], y[:140][::-1]]
    y2 = y + "A" * 300
    for ref in (y, y2):
        engine.set_scoring_match(pkg.MODE_SAT_U8, 40, -7, 9)
        engine.set_reference(ref)
        r = engine.align(xs, cons_stride=2200)
        for i, x in enumerate(xs):
            w = o.align(x, ref, mode=o.MODE_SAT_U8, match=40, mismatch=-7, gap=9)
            _check(r, i, w, tag=("ties", i))
            assert w["score"] == 255 or i == 8


def test_chunk_range_error(engine, pkg):
    """plocalaligner.cpp:52: overlap > piece aborts the reference; the C ABI returns SWB_ERR_RANGE."""
    engine.set_scoring_match(pkg.MODE_SAT_U8, 3, -3, 2)
    engine.set_reference("ACGT" * 40)
    with pytest.raises(pkg.SwbError) as ei:
        engine.align(["ACGTACGTAC" * 10], npiece=4, ratio=2.0)
    assert ei.value.code == -3


def test_cpp_shims_and_driver(tmp_path, data_small):
    """The C++ drop-in layer: the reference's unit tests with the aligner type swapped (tests/cpp/test_shim.cpp)
    and the batched sw_solve_small driver, whose output CSV must carry the reference's pos/score columns."""
    import os
    import subprocess
    from conftest import ROOT, GOLDEN
    exe = os.path.join(ROOT, "tests", "cpp", "test_shim")
    if not os.path.isfile(exe):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_shim.cpp"),
                               "-L", os.path.join(ROOT, "parallel-genomeseq_b200"), "-lswb200", "-lpthread", "-Wl,-rpath," + os.path.join(ROOT, "parallel-genomeseq_b200")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "SHIM OK" in out.stdout and "LAZY n=600" in out.stdout and "MULTI-GPU OK" in out.stdout, out.stdout + out.stderr
    drv = os.path.join(ROOT, "parallel-genomeseq_b200", "drivers", "sw_solve_small")
    if not os.path.isfile(drv):
        subprocess.check_call(["make", "-C", os.path.dirname(drv)])
    for args, gold in (([], "data_small_sw_skewed.csv"), (["--npiece", "17", "--ratio", "2.0"], "data_small_p17.csv"), (["--float"], "data_small_sw_float.csv")):
        csv_out = str(tmp_path / "align_output.csv")
        r = subprocess.run([drv, os.path.join(GOLDEN, "data_small", "genome.chr22.5K.fa"), os.path.join(GOLDEN, "data_small", "data_small_ground_truth.csv"), csv_out] + args,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "GCUP:" in r.stdout, r.stdout + r.stderr
        want = read_golden_csv(gold)
        with open(csv_out) as f:
            rows = f.read().strip().split("\n")
        assert rows[0].endswith(",pos_pred,score") and len(rows) == 1171
        for line, g in zip(rows[1:], want):
            f_ = line.split(", ")
            assert int(f_[-2]) == g["pos"] and float(f_[-1]) == g["score"]


def test_long_sequences_row_strips(engine, pkg):
    """Sequences longer than 32 lanes x 32 rows are cut into row strips (boundary rows in HBM).  Oracle parity
    for both modes, incl. len(x) > len(y) and a traceback that crosses strip boundaries."""
    rng = np.random.default_rng(21)
    y = "".join(rng.choice(list("ACGT"), size=2600))
    xs = []
    for m, s0 in ((1025, 0), (1500, 700), (2047, 300), (2300, 100), (3100, 0)):
        base = list((y * 2)[s0:s0 + m])
        for q in range(m):
            if rng.random() < 0.06:
                base[q] = str(rng.choice(list("ACGT")))
        xs.append("".join(base))
    xs.append("".join(rng.choice(list("ACGT"), size=1300)))
    for mode, omode, sc in ((pkg.MODE_SAT_U8, o.MODE_SAT_U8, (3, -3, 2)), (pkg.MODE_EXACT, o.MODE_EXACT, (3, -3, 2)), (pkg.MODE_EXACT, o.MODE_EXACT, (2, -3, 4))):
        engine.set_scoring_match(mode, *sc)
        engine.set_reference(y)
        r = engine.align(xs, cons_stride=6000)
        for i, x in enumerate(xs):
            w = o.align(x, y, mode=omode, match=sc[0], mismatch=sc[1], gap=sc[2])
            _check(r, i, w, tag=("strips", mode, sc, len(x)))
            assert tuple(r["end"][i]) == w["end"]


def test_long_pair_c5_shape(engine, pkg):
    """BASELINE config 5 at the reference's own published shape (10 kbp read vs 30 kbp reference, py/eval.py:54),
    EXACT semantics (omp_sw_solve_small.cpp:167) and SAT_U8 (MTSIMD, :164), against the oracle."""
    ref = synth.c3_reference(30_000, seed=26)
    reads = synth.mutated_reads(ref, 2, 10_000, seed=27, sub=0.02, ins=0.002, dele=0.002)
    for mode, omode in ((pkg.MODE_EXACT, o.MODE_EXACT), (pkg.MODE_SAT_U8, o.MODE_SAT_U8)):
        engine.set_scoring_match(mode, 3, -3, 2)
        engine.set_reference(ref)
        r = engine.align(reads, cons_stride=25_000)
        for i, x in enumerate(reads):
            _check(r, i, o.align(x, ref, mode=omode), tag=("c5", mode))


def test_pipelined_strips_sparse_publishing(engine, pkg):
    """Few long pairs against a long reference: the strips of a pair run concurrently (score_units_kernel) and a
    strip publishes its progress only every few checkpoint blocks (interval > 1 needs many blocks per strip,
    which the 30 kbp case above does not have).  Oracle parity, both modes, incl. a read without a good match."""
    ref = synth.c3_reference(120_000, seed=41)
    reads = synth.mutated_reads(ref, 2, 1_400, seed=42, sub=0.03, ins=0.003, dele=0.003)
    rng = np.random.default_rng(43)
    reads.append("".join(rng.choice(list("ACGT"), size=1_150)))
    for mode, omode in ((pkg.MODE_EXACT, o.MODE_EXACT), (pkg.MODE_SAT_U8, o.MODE_SAT_U8)):
        engine.set_scoring_match(mode, 3, -3, 2)
        engine.set_reference(ref)
        r = engine.align(reads, cons_stride=4_000)
        st = engine.stats()
        assert st["lanes_per_pair"] == 32 and st["rows_per_lane"] * 32 < 1_150, st   # really cut into strips
        for i, x in enumerate(reads):
            w = o.align(x, ref, mode=omode)
            _check(r, i, w, tag=("units", mode, len(x)))
            assert tuple(r["end"][i]) == w["end"]


def test_dense_matrix_accessor(engine, pkg):
    """operator()(row, col) surface: every cell of H through the device path equals the oracle
    (test/test_skewedmatrix.cpp:39-66 compares the two SMTs cell by cell; test_localaligner.cpp:31-42 golden H)."""
    want = o.matrix("GGTTGACTA", "TGTTACGG", mode=o.MODE_SAT_U8)
    for mode in (pkg.MODE_SAT_U8, pkg.MODE_EXACT):
        engine.set_scoring_match(mode, 3, -3, 2)
        engine.set_reference("TGTTACGG")
        assert (engine.matrix("GGTTGACTA") == want).all()
    rng = np.random.default_rng(17)
    for (m, n) in ((9, 7), (7, 9), (130, 77), (64, 300), (700, 90), (1300, 200)):
        x = "".join(rng.choice(list("ACGT"), size=m))
        y = "".join(rng.choice(list("ACGT"), size=n))
        for mode, omode, sc in ((pkg.MODE_SAT_U8, o.MODE_SAT_U8, (40, -7, 9)), (pkg.MODE_EXACT, o.MODE_EXACT, (3, -3, 2))):
            engine.set_scoring_match(mode, *sc)
            engine.set_reference(y)
            got = engine.matrix(x)
            exp = o.matrix(x, y, mode=omode, match=sc[0], mismatch=sc[1], gap=sc[2])
            assert (got == exp).all(), (m, n, mode)
    t = synth.blosum62_table()
    q = synth.c4_queries(1, 120)[0]
    prot = synth.c4_database(3)[0][:200]
    engine.set_scoring_table(pkg.MODE_EXACT, t, 10)
    engine.set_reference(q)
    assert (engine.matrix(prot) == o.matrix(prot, q, mode=o.MODE_EXACT, table=t, gap=10)).all()


def test_consensus_truncation_flag_and_errors(engine, pkg):
    """A consensus longer than cons_stride is flagged (score / end cell stay exact); bad arguments return codes."""
    rng = np.random.default_rng(4)
    y = "".join(rng.choice(list("ACGT"), size=900))
    x = y[100:400]
    engine.set_scoring_match(pkg.MODE_EXACT, 3, -3, 2)
    engine.set_reference(y)
    full = engine.align([x], cons_stride=700)
    cut = engine.align([x], cons_stride=64)
    assert full["flags"][0] == 0 and cut["flags"][0] & 1
    assert int(cut["score"][0]) == int(full["score"][0]) == 900 and tuple(cut["end"][0]) == tuple(full["end"][0])
    with pytest.raises(pkg.SwbError) as ei:
        engine.align(["ACGT", ""])
    assert ei.value.code == -2
    with pytest.raises(pkg.SwbError) as ei:
        engine.set_scoring_match(pkg.MODE_EXACT, 2.5, -3, 2)
    assert ei.value.code == -4
    engine.set_scoring_match(pkg.MODE_EXACT, 3, -3, 2)


def test_edge_cases_against_oracle(engine, pkg):
    """Ragged and degenerate inputs: single read, odd batch (an unpaired last task), 1- and 2-symbol sequences,
    bytes outside ACGT ('N', lower case: scored by raw byte equality, similaritymatrix.cpp:415), duplicates,
    a read equal to the whole reference window, a reference shorter than every read."""
    rng = np.random.default_rng(31)
    y = "".join(rng.choice(list("ACGTN"), size=400, p=[0.24, 0.24, 0.24, 0.24, 0.04]))
    xs = ["A", "AC", y[10:11], y[5:130], y[5:130], y[100:240].lower(), y[100:240], "N" * 30 + y[50:90], y[:399], y[1:400],
          "".join(rng.choice(list("ACGT"), size=77)), y[200:333] + "acgtn"]
    for mode, omode in ((pkg.MODE_SAT_U8, o.MODE_SAT_U8), (pkg.MODE_EXACT, o.MODE_EXACT)):
        engine.set_scoring_match(mode, 3, -3, 2)
        engine.set_reference(y)
        for batch in (xs, xs[:1], xs[:3], xs[3:4] * 5):
            r = engine.align(batch, cons_stride=900)
            for i, x in enumerate(batch):
                w = o.align(x, y, mode=omode)
                if w["score"] == 0:
                    assert int(r["score"][i]) == 0 and int(r["len"][i]) == 0 and int(r["pos"][i]) == 0
                    continue
                _check(r, i, w, tag=("edge", mode, i, len(x)))
    # reference shorter than the reads, one-column reference
    for ref in ("ACGTTGCA", "G"):
        engine.set_scoring_match(pkg.MODE_SAT_U8, 3, -3, 2)
        engine.set_reference(ref)
        batch = ["TTACGTTGCATT", "GGGGGGGGGG", "ACGTTGCAA"]
        r = engine.align(batch, cons_stride=64)
        for i, x in enumerate(batch):
            if len(x) == len(ref):
                continue   # square input: reference defect
            _check(r, i, o.align(x, ref, mode=o.MODE_SAT_U8), tag=("short-ref", ref, i))


def test_exact_scores_beyond_16_bits_use_wide_lanes(engine, pkg, monkeypatch):
    """Similarity_Matrix is exact to 2^24 (f32, similaritymatrix.cpp:49-54).  EXACT scores that could leave the 16-bit
    lanes run with one s32 cell per register (AM_WIDE) — batched kernels, row strips, pipelined strips, chunking, the
    dense-matrix accessor — against the oracle; beyond 2^24 the batch is refused loudly, never computed wrongly."""
    rng = np.random.default_rng(123)
    y = "".join(rng.choice(list("ACGT"), size=2100))
    xs = []
    for m, s0 in ((600, 100), (601, 900), (40, 5), (1500, 300), (333, 1700), (1100, 0)):
        x = list((y * 2)[s0:s0 + m])
        for q in range(m):
            if rng.random() < 0.05:
                x[q] = str(rng.choice(list("ACGT")))
        xs.append("".join(x))
    for (ma, mi, g) in ((100, -90, 40), (1000, -3, 2)):
        engine.set_scoring_match(pkg.MODE_EXACT, ma, mi, g)
        engine.set_reference(y)
        r = engine.align(xs, cons_stride=5000)
        assert max(int(v) for v in r["score"]) > 40_000
        for i, x in enumerate(xs):
            w = o.align(x, y, mode=o.MODE_EXACT, match=ma, mismatch=mi, gap=g)
            _check(r, i, w, tag=("wide", ma, len(x)))
            assert tuple(r["end"][i]) == w["end"]
        rc = engine.align(xs[:3], npiece=3, ratio=1.5, cons_stride=5000)
        for i, x in enumerate(xs[:3]):
            _check(rc, i, o.align_chunked(x, y, 3, 1.5, mode=o.MODE_EXACT, match=ma, mismatch=mi, gap=g), tag=("wide-chunked", ma, i))
    engine.set_scoring_match(pkg.MODE_EXACT, 100, -90, 40)
    engine.set_reference(y[:300])
    assert (engine.matrix(xs[2]) == o.matrix(xs[2], y[:300], mode=o.MODE_EXACT, match=100, mismatch=-90, gap=40)).all()
    # the wide kernels on ordinary scores (forced): identical results to the 16-bit lanes, incl. pipelined strips
    monkeypatch.setenv("SWB_FORCE_WIDE", "1")
    ref = synth.c3_reference(60_000, seed=41)
    reads = synth.mutated_reads(ref, 3, 2_500, seed=42, sub=0.03, ins=0.003, dele=0.003)
    engine.set_scoring_match(pkg.MODE_EXACT, 3, -3, 2)
    engine.set_reference(ref)
    r = engine.align(reads, cons_stride=8_000)
    assert engine.stats()["kernel_kind"] == 1
    for i, x in enumerate(reads):
        w = o.align(x, ref, mode=o.MODE_EXACT, linear=True)
        _check(r, i, w, tag=("wide-units", len(x)))
    monkeypatch.delenv("SWB_FORCE_WIDE")
    engine.set_scoring_match(pkg.MODE_EXACT, 4000, -3, 2)
    engine.set_reference("ACGT" * 2000)
    with pytest.raises(pkg.SwbError) as ei:
        engine.align(["ACGT" * 1500])
    assert ei.value.code == -5
    engine.set_scoring_match(pkg.MODE_EXACT, 3, -3, 2)


def test_chunked_custom_scoring_uses_default_for_the_final_alignment(engine, pkg):
    """OMPParallelLocalAligner picks the piece with the constructor's scoring but re-aligns it with the
    DEFAULT scoring (plocalaligner.cpp:135).  Batch of ragged reads, custom scoring, against the oracle."""
    rng = np.random.default_rng(77)
    y = "".join(rng.choice(list("ACGT"), size=2400))
    xs = []
    for m in (40, 40, 61, 61, 100, 100, 100, 33):
        s0 = int(rng.integers(0, 2400 - m)); x = list(y[s0:s0 + m])
        for q in range(m):
            if rng.random() < 0.07:
                x[q] = str(rng.choice(list("ACGT")))
        xs.append("".join(x))
    for mode, omode in ((pkg.MODE_SAT_U8, o.MODE_SAT_U8), (pkg.MODE_EXACT, o.MODE_EXACT)):
        engine.set_scoring_match(mode, 5, -4, 3)
        engine.set_reference(y)
        r = engine.align(xs, npiece=5, ratio=2.0, cons_stride=600)
        for i, x in enumerate(xs):
            _check(r, i, o.align_chunked(x, y, 5, 2.0, mode=omode, match=5, mismatch=-4, gap=3), tag=("f8", mode, i))
    engine.set_scoring_match(pkg.MODE_SAT_U8, 3, -3, 2)


def test_drivers_sw_solve_big_and_uniprot_search(tmp_path, data_small, c4_sample):
    """The two other batched drivers: sw_solve_big (single-line reference, nrepeat min-of-N, [INFO]
    GCUPS lines; npiece=2 means 4 pieces like sw_solve_big.cpp:78) and the UniProt search (multi-FASTA database,
    'read,pos_pred,score' rows like mpi_sw_solve_uniprot.cpp:151-168) against the reference's goldens."""
    import os
    import subprocess
    from conftest import ROOT, GOLDEN
    ddir = os.path.join(ROOT, "parallel-genomeseq_b200", "drivers")
    if not (os.path.isfile(os.path.join(ddir, "sw_solve_big")) and os.path.isfile(os.path.join(ddir, "sw_search_uniprot"))):
        subprocess.check_call(["make", "-C", ddir])
    ref, truth = data_small
    fa = tmp_path / "custom_ref_1.fa"
    fa.write_text(ref + "\n")
    reads_csv = os.path.join(GOLDEN, "data_small", "data_small_ground_truth.csv")
    for args, gold in ((["0", "2"], "data_small_sw_skewed.csv"), (["2", "3"], "data_small_p4.csv")):
        r = subprocess.run([os.path.join(ddir, "sw_solve_big")] + args + ["--fa", str(fa), "--reads", reads_csv], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "GCUPS avg:" in r.stdout and "[INFO] Average SW iter_ad_read times:" in r.stdout
        g0 = read_golden_csv(gold)[0]
        assert f"[INFO] first read: pos {g0['pos']} score {g0['score']}" in r.stdout, r.stdout
    q = tmp_path / "query.fasta"
    q.write_text(">query\n" + c4_sample["query"] + "\n")
    db = tmp_path / "db.fasta"
    with open(db, "w") as f:
        for k, e in enumerate(c4_sample["entries"]):
            f.write(f">sp|P{k:05d}|synthetic\n")
            for o_ in range(0, len(e["x"]), 60):
                f.write(e["x"][o_:o_ + 60] + "\n")
    out_csv = tmp_path / "out.csv"
    r = subprocess.run([os.path.join(ddir, "sw_search_uniprot"), str(q), str(db), str(out_csv), "--blosum62", str(c4_sample["gap"])], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "GCUPS" in r.stdout, r.stdout + r.stderr
    rows = out_csv.read_text().strip().split("\n")
    assert rows[0] == "read,pos_pred,score" and len(rows) == len(c4_sample["entries"]) + 1
    for line, e in zip(rows[1:], c4_sample["entries"]):
        f_ = line.split(", ")
        assert f_[0] == e["x"][:126] and int(f_[1]) == e["pos"] and float(f_[2]) == e["score"]


def test_driver_fine_grain_benchmark_surface(tmp_path):
    """omp_sw_solve_small replacement: same argv, align-output rows and timing-CSV schema
    (omp_sw_solve_small.cpp:66-73,233-239); long reads (row strips) in both arithmetic flavours vs the oracle."""
    import os
    import subprocess
    from conftest import ROOT
    drv = os.path.join(ROOT, "parallel-genomeseq_b200", "drivers", "omp_sw_solve_small")
    if not os.path.isfile(drv):
        subprocess.check_call(["make", "-C", os.path.dirname(drv)])
    ref = synth.c3_reference(20_000, seed=5)
    reads = synth.mutated_reads(ref, 3, 3_000, seed=6, sub=0.02, ins=0.002, dele=0.002)
    fa = tmp_path / "ref.fa"
    fa.write_text(ref[:9000] + "\n" + ref[9000:] + "\n")          # every line is concatenated, no header
    rcsv = tmp_path / "reads.csv"
    rcsv.write_text("index,QNAME,SEQ,POS\n" + "".join(f"{i},r{i},{x},0\n" for i, x in enumerate(reads)))
    timing = tmp_path / "timing.csv"
    for fg, omode in ((1, o.MODE_EXACT), (-1, o.MODE_SAT_U8)):
        out_csv = tmp_path / f"align_{fg}.csv"
        r = subprocess.run([drv, "solve_small", "2", "4", str(fg), str(timing), str(fa), str(rcsv), str(out_csv)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "GCUPS" in r.stdout, r.stdout + r.stderr
        rows = out_csv.read_text().strip().split("\n")
        assert rows[0] == "index,QNAME,SEQ,POS,pos_pred,score" and len(rows) == 3      # the first n_reads = 2 reads
        for line, x in zip(rows[1:], reads):
            w = o.align(x, ref, mode=omode)
            f_ = line.split(", ")
            assert int(f_[-2]) == w["pos"] and float(f_[-1]) == w["score"]
    t = timing.read_text().strip().split("\n")
    assert t[0] == "n_reads,n_threads,finegrain_type,avg_t_calcscore,avg_t_adread,avg_t_adisum" and len(t) == 3
    assert t[1].startswith("2,4,1,") and t[2].startswith("2,4,-1,")


def test_ragged_read_lengths_keep_efficient_geometry(engine, pkg):
    """Reads trimmed to 100..150 bp (many distinct lengths) must keep the sub-warp geometries (8 x 16 / 8 x 19),
    not fall back to thin strips, and stay bit-exact."""
    rng = np.random.default_rng(41)
    y = "".join(rng.choice(list("ACGT"), size=30_000))
    xs = []
    for k in range(6000):
        m = int(rng.integers(100, 151)); s0 = int(rng.integers(0, 30_000 - m))
        xs.append(y[s0:s0 + m])
    engine.set_scoring_match(pkg.MODE_SAT_U8, 3, -3, 2)
    engine.set_reference(y)
    r = engine.align(xs, consensus=True)
    st = engine.stats()
    assert st["lanes_per_pair"] == 8 and st["rows_per_lane"] in (16, 19), st
    for i in range(0, len(xs), 37):
        _check(r, i, o.align(xs[i], y, mode=o.MODE_SAT_U8), tag=("ragged", i, len(xs[i])))


def test_randomised_fuzz_against_oracle():
    """tools/fuzz_parity.py: random shapes (1..3500 columns, 1..2600 rows incl. row strips), alphabets, match and
    tabulated scorings (incl. zero gap / zero mismatch), both modes, chunking, and kernel knobs (select, columns
    per step, sub-batch size) against the oracle.  Longer runs: `python tools/fuzz_parity.py 1500 <seed>`."""
    import os
    import subprocess
    import sys as _sys
    from conftest import ROOT
    env = {k: v for k, v in os.environ.items() if not k.startswith("SWB_")}
    r = subprocess.run([_sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), "200", "777"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "fuzz ok" in r.stdout, r.stdout + r.stderr


def test_randomised_fuzz_query_stationary():
    """The same fuzzer with the query-stationary kernels forced on wherever they apply (unchunked, reference of at
    most 1024 symbols): random and asymmetric tables, both modes, zero gaps, ragged batches, against the oracle."""
    import os
    import subprocess
    import sys as _sys
    from conftest import ROOT
    env = {k: v for k, v in os.environ.items() if not k.startswith("SWB_")}
    env["SWB_QSTAT"] = "1"
    r = subprocess.run([_sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), "300", "2024"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "fuzz ok" in r.stdout, r.stdout + r.stderr


def test_query_stationary_database_search(engine, pkg, c4_sample, monkeypatch):
    """BASELINE config 4 through the query-stationary kernels (sw_qs.cuh): the golden C4 sample (forced on: it has
    fewer sequences than the automatic threshold), then a database large enough to switch the mode on by itself
    (BLOSUM62 and an asymmetric table), checked against the oracle incl. the arg-max cell."""
    monkeypatch.setenv("SWB_QSTAT", "1")
    engine.set_scoring_table(pkg.MODE_EXACT, synth.blosum62_table(), c4_sample["gap"])
    engine.set_reference(c4_sample["query"])
    ents = c4_sample["entries"]
    r = engine.align([e["x"] for e in ents], cons_stride=6000)
    st = engine.stats()
    assert (st["lanes_per_pair"], st["rows_per_lane"], st["kernel_launches"]) == (16, 19, 2), st   # one score + one trace launch
    for i, e in enumerate(ents):
        _check(r, i, e, tag="c4-qs")
    monkeypatch.delenv("SWB_QSTAT")
    rng = np.random.default_rng(77)
    query = synth.c4_queries(1, 130, seed=5)[0]
    db = synth.c4_database(2500, seed=6)
    asym = rng.integers(-5, 4, size=(256, 256)).astype(np.int32)
    np.fill_diagonal(asym, rng.integers(2, 10, size=256))
    for table, gap in ((synth.blosum62_table(), 10), (asym, 3)):
        engine.set_scoring_table(pkg.MODE_EXACT, table, gap)
        engine.set_reference(query)
        r = engine.align(db, cons_stride=3000)
        st = engine.stats()
        assert st["lanes_per_pair"] * st["rows_per_lane"] >= len(query) and st["kernel_launches"] == 2, st   # automatic
        for i in range(0, len(db), 9):
            w = o.align(db[i], query, mode=o.MODE_EXACT, table=table, gap=gap)
            if w["score"] == 0:
                assert int(r["score"][i]) == 0 and int(r["len"][i]) == 0
                continue
            _check(r, i, w, tag=("qs-auto", gap, i, len(db[i])))
            assert tuple(r["end"][i]) == w["end"]


def test_reference_sharded_align_single_rank(engine, pkg, tmp_path):
    """sharding.reference_sharded_align with the engine as the piece aligner (world_size 1 over gloo: one piece =
    the whole reference): the N > 1 logic is covered on CPU (tests/test_host_cpu.py), this checks the engine glue."""
    import torch.distributed as dist
    sh = __import__("importlib").import_module("parallel-genomeseq_b200.sharding")
    ref = synth.c3_reference(20_000, seed=61)
    reads = synth.mutated_reads(ref, 6, 150, seed=62)
    engine.set_scoring_match(pkg.MODE_EXACT, 3, -3, 2)
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("gloo", init_method="file://" + str(tmp_path / "rendezvous"), rank=0, world_size=1)   # no port needed
    try:
        sc, ps, win = sh.reference_sharded_align(sh.engine_aligner(engine, cons_stride=400), reads, ref, 2.0, pkg.make_string_range)
    finally:
        if created:
            dist.destroy_process_group()
    for i, x in enumerate(reads):
        w = o.align(x, ref, mode=o.MODE_EXACT)
        assert (int(sc[i]), int(ps[i]), int(win[i])) == (w["score"], w["pos"], 0)


# ---- parity at the sizes and geometries that are TIMED (golden vectors of the linear-memory oracle) ---------------
def _load_json(name):
    import json
    import os
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def c3_10k():
    doc = _load_json("c3_10k.json")
    ref = synth.c3_reference(doc["ref_len"])
    assert hashlib.sha256(ref.encode()).hexdigest() == doc["ref_sha256"]
    reads = synth.c3_reads(ref, doc["n_reads"])
    assert hashlib.sha256("".join(reads).encode()).hexdigest() == doc["reads_sha256"]
    return ref, reads, doc["rows"]


def _check_rows(r, rows, idx, tag):
    for k, i in enumerate(idx):
        d = hashlib.sha256((r["cx"][k] + "|" + r["cy"][k]).encode()).hexdigest()[:16]
        got = [int(r["score"][k]), int(r["pos"][k]), int(r["end"][k][0]), int(r["end"][k][1]), int(r["len"][k]), d]
        assert got == rows[i] and r["flags"][k] == 0, (tag, i, got, rows[i])


@pytest.mark.parametrize("B", [1024, 2048])
def test_timed_geometry_parity(engine, pkg, c3_sample, c3_10k, monkeypatch, B):
    """The geometry bench.py times (8 lanes x 19 rows, checkpoint period 1024 / 2048, several sub-batches that reuse
    the HBM work buffers) on the golden C3 sample and 200 further reads — score, pos, arg-max cell, both consensus
    strings.  The knobs only force what a 75 776-pair batch gets by itself."""
    ref, reads, rows = c3_10k
    monkeypatch.setenv("SWB_FORCE_L", "8")
    monkeypatch.setenv("SWB_FORCE_R", "19")
    monkeypatch.setenv("SWB_FORCE_B", str(B))
    monkeypatch.setenv("SWB_CHUNK_PAIRS", "64")          # 224 reads = 112 pairs -> two sub-batches
    engine.set_scoring_match(pkg.MODE_SAT_U8, 3, -3, 2)
    engine.set_reference(ref)
    batch = [e["x"] for e in c3_sample["reads"]] + reads[:200]
    r = engine.align(batch)
    st = engine.stats()
    assert (st["lanes_per_pair"], st["rows_per_lane"], st["block_steps"]) == (8, 19, B), st
    for i, e in enumerate(c3_sample["reads"]):
        _check(r, i, e, tag=("timed-geometry", B))
    n0 = len(c3_sample["reads"])
    sub = dict(score=r["score"][n0:], pos=r["pos"][n0:], end=r["end"][n0:], len=r["len"][n0:], flags=r["flags"][n0:], cx=r["cx"][n0:], cy=r["cy"][n0:])
    _check_rows(sub, rows, range(200), ("timed-geometry", B))


def test_c3_ten_thousand_reads(engine, pkg, c3_10k):
    """BASELINE config 3 on 10 000 reads (indels included) against the linear-memory oracle's golden rows."""
    ref, reads, rows = c3_10k
    engine.set_scoring_match(pkg.MODE_SAT_U8, 3, -3, 2)
    engine.set_reference(ref)
    r = engine.align(reads)
    _check_rows(r, rows, range(len(reads)), "c3-10k")


def test_c5_full_size(engine, pkg):
    """BASELINE config 5 at its STATED size: 10 kbp reads against the 51 Mbp reference, EXACT (omp_sw_solve_small.cpp:167)
    and SAT_U8 (MTSIMD, :164) — every read the golden file holds, all of score / pos / arg-max cell / consensus."""
    doc = _load_json("c5_full.json")
    ref = synth.c5_reference(doc["ref_len"])
    assert hashlib.sha256(ref.encode()).hexdigest() == doc["ref_sha256"]
    reads = synth.c5_reads(ref, 16, doc["read_len"])[:doc["n_reads"]]
    assert [hashlib.sha256(x.encode()).hexdigest() for x in reads] == doc["reads_sha256"]
    for mode, key in ((pkg.MODE_EXACT, "exact"), (pkg.MODE_SAT_U8, "sat_u8")):
        engine.set_scoring_match(mode, 3, -3, 2)
        engine.set_reference(ref)
        r = engine.align(reads, cons_stride=25_000)
        for i, e in enumerate(doc[key]):
            got = (int(r["score"][i]), int(r["pos"][i]), [int(v) for v in r["end"][i]], int(r["len"][i]),
                   hashlib.sha256(r["cx"][i].encode()).hexdigest(), hashlib.sha256(r["cy"][i].encode()).hexdigest())
            assert got == (e["score"], e["pos"], e["end"], e["len"], e["cx_sha256"], e["cy_sha256"]), (key, i, got[:4], e["score"], e["pos"], e["end"], e["len"])
            assert r["flags"][i] == 0


def test_reference_unit_test_with_type_swap():
    """The reference's OWN test/test_localaligner.cpp (score 13, pos 2, consensus CAGTTG / CA-TTG), compiled unmodified
    against the reference headers and its vendored googletest with only the aligner type swapped to the CUDA shims
    (-DSWB_WITH_REFERENCE_HEADERS, the mode INTEGRATION.md tells maintainers to use).  The binary is built in the build
    container (tests/cpp/build_ref_tests.py needs /root/reference) and travels with the snapshot."""
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "tests", "cpp", "ref_test_localaligner")
    if not os.path.isfile(exe):
        if not os.path.isdir("/root/reference"):
            pytest.skip("ref_test_localaligner was not built (needs /root/reference at build time)")
        subprocess.check_call([os.sys.executable, os.path.join(ROOT, "tests", "cpp", "build_ref_tests.py")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "[  PASSED  ] 2 tests" in r.stdout, r.stdout + r.stderr


def test_drivers_over_several_gpus(tmp_path, data_small, c4_sample):
    """--gpus N of the batched drivers (one host thread + one context per device behind the C ABI, reads block-
    partitioned like mpi_sw_solve_small.cpp:52-55, database partitioned by residues): same CSV as the goldens.  Runs
    with every device count the box has (1 on a single-GPU box: the threaded path with one worker)."""
    import os
    import subprocess
    from conftest import ROOT, GOLDEN
    ddir = os.path.join(ROOT, "parallel-genomeseq_b200", "drivers")
    import torch
    ndev = torch.cuda.device_count()
    want = read_golden_csv("data_small_sw_skewed.csv")
    for g in sorted({0, min(2, ndev), ndev}):
        csv_out = str(tmp_path / f"align_{g}.csv")
        r = subprocess.run([os.path.join(ddir, "sw_solve_small"), os.path.join(GOLDEN, "data_small", "genome.chr22.5K.fa"), os.path.join(GOLDEN, "data_small", "data_small_ground_truth.csv"), csv_out, "--gpus", str(g)],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "GCUP:" in r.stdout and (g == 1 or "reads divided over" in r.stdout), r.stdout + r.stderr
        rows = open(csv_out).read().strip().split("\n")
        assert len(rows) == 1171
        for line, gd in zip(rows[1:], want):
            f_ = line.split(", ")
            assert int(f_[-2]) == gd["pos"] and float(f_[-1]) == gd["score"]
    q = tmp_path / "query.fasta"
    q.write_text(">query\n" + c4_sample["query"] + "\n")
    db = tmp_path / "db.fasta"
    with open(db, "w") as f:
        for k, e in enumerate(c4_sample["entries"]):
            f.write(f">sp|P{k:05d}|synthetic\n" + e["x"] + "\n")
    out_csv = tmp_path / "out.csv"
    r = subprocess.run([os.path.join(ddir, "sw_search_uniprot"), str(q), str(db), str(out_csv), "--blosum62", str(c4_sample["gap"]), "--gpus", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "partitioned by residues over" in r.stdout, r.stdout + r.stderr
    rows = out_csv.read_text().strip().split("\n")
    for line, e in zip(rows[1:], c4_sample["entries"]):
        f_ = line.split(", ")
        assert int(f_[1]) == e["pos"] and float(f_[2]) == e["score"]
    # the packed 5-bit database blob: same rows in the original database order, no FASTA parsing per run
    packed = tmp_path / "db.swbdb"
    r = subprocess.run([os.path.join(ddir, "sw_search_uniprot"), "--pack", str(db), str(packed)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out2 = tmp_path / "out2.csv"
    r = subprocess.run([os.path.join(ddir, "sw_search_uniprot"), str(q), str(packed), str(out2), "--blosum62", str(c4_sample["gap"])], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "packed blob" in r.stdout, r.stdout + r.stderr
    assert out2.read_text() == out_csv.read_text()


def test_rebind_reference_keeps_the_database_resident(engine, pkg):
    """swb_batch_rebind_reference: many queries against one staged database (query-stationary mode); every query's result
    equals a fresh stage, and a batch that is not in that mode refuses."""
    db = synth.c4_database(3000, seed=9)
    qs = synth.c4_queries(3, 300, seed=10) + synth.c4_queries(1, 250, seed=11)
    t = synth.blosum62_table()
    engine.set_scoring_table(pkg.MODE_EXACT, t, 10)
    engine.set_reference(qs[0])
    engine.stage(db, consensus=False)
    for q in qs:
        engine.rebind_reference(q)
        engine.run()
        r = engine.fetch()
        for i in range(0, len(db), 97):
            w = o.align(db[i], q, mode=o.MODE_EXACT, table=t, gap=10)
            if w["score"] == 0:
                assert int(r["score"][i]) == 0
                continue
            assert (int(r["score"][i]), int(r["pos"][i]), tuple(int(v) for v in r["end"][i])) == (w["score"], w["pos"], tuple(w["end"])), (len(q), i)
        assert engine.stats()["cells_reference"] == sum(len(p) for p in db) * len(q)
    engine.set_scoring_match(pkg.MODE_SAT_U8, 3, -3, 2)
    engine.set_reference("ACGT" * 300)
    engine.stage(["ACGTACGTTT" * 10] * 4)
    with pytest.raises(pkg.SwbError) as ei:
        engine.rebind_reference("ACGT" * 100)
    assert ei.value.code == -6


def test_checked_build_runs_clean():
    """libswb200_checked.so = the same sources with -DSWB_CHECKED: every store into a work buffer (checkpoints, block maxima,
    strip boundary rows, local checkpoints), every access of the pass-2 ring and the unmasked reference loads carry an index
    check that turns into an error code.  compute-sanitizer is closed on the GPU pool (profiles/sanitizer_r02_closed.txt);
    this is the stand-in for its memcheck: a batch through every kernel family and a fuzz run must pass with no check firing."""
    import os
    import subprocess
    import sys as _sys
    from conftest import ROOT
    lib = os.path.join(ROOT, "parallel-genomeseq_b200", "libswb200_checked.so")
    if not os.path.isfile(lib):
        subprocess.check_call([_sys.executable, os.path.join(ROOT, "parallel-genomeseq_b200", "build.py"), "--checked"])
    env = {k: v for k, v in os.environ.items() if not k.startswith("SWB_")}
    env["SWB_LIB_PATH"] = lib
    r = subprocess.run([_sys.executable, os.path.join(ROOT, "tools", "sanitize_case.py")], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "sanitize case ok" in r.stdout and "SWB_CHECKED" in r.stdout, r.stdout + r.stderr
    r = subprocess.run([_sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), "150", "4711"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "fuzz ok" in r.stdout, r.stdout + r.stderr
