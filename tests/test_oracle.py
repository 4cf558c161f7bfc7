"""Pin the CPU oracle (oracle/sw_oracle.c) against the reference's own known answers and against the
golden vectors dumped from the compiled reference (tests/golden/make_golden.py).  CPU only."""
import hashlib
import os

import numpy as np
import pytest

import pyoracle as o
import synth
from conftest import GOLDEN, read_golden_csv


# ---- the reference's own unit tests -------------------------------------------------------------------
def test_known_answer_score_pos():
    """test/test_localaligner.cpp:8-27: SWAligner<Skewed>("GGTTGACTA","TGTTACGG") -> score 13, pos 2."""
    for mode in (o.MODE_SAT_U8, o.MODE_EXACT):
        r = o.align("GGTTGACTA", "TGTTACGG", mode=mode)
        assert (r["score"], r["pos"]) == (13, 2)


def test_known_answer_consensus():
    """test/test_localaligner.cpp:53-59: consensus_x "CAGTTG", consensus_y "CA-TTG" (end -> start)."""
    for mode in (o.MODE_SAT_U8, o.MODE_EXACT):
        r = o.align("GGTTGACTA", "TGTTACGG", mode=mode)
        assert (r["cx"], r["cy"]) == ("CAGTTG", "CA-TTG")


def test_known_answer_matrix():
    """test/test_localaligner.cpp:31-42 (commented-out golden H for the Wikipedia pair)."""
    want = np.array([
        [0, 0, 0, 0, 0, 0, 0, 0, 0],
        [0, 0, 3, 1, 0, 0, 0, 3, 3],
        [0, 0, 3, 1, 0, 0, 0, 3, 6],
        [0, 3, 1, 6, 4, 2, 0, 1, 4],
        [0, 3, 1, 4, 9, 7, 5, 3, 2],
        [0, 1, 6, 4, 7, 6, 4, 8, 6],
        [0, 0, 4, 3, 5, 10, 8, 6, 5],
        [0, 0, 2, 1, 3, 8, 13, 11, 9],
        [0, 3, 1, 5, 4, 6, 11, 10, 8],
        [0, 1, 0, 3, 2, 7, 9, 8, 7]], dtype=np.int32)
    for mode in (o.MODE_SAT_U8, o.MODE_EXACT):
        assert (o.matrix("GGTTGACTA", "TGTTACGG", mode=mode) == want).all()


def test_index_round_trip():
    """test/test_skewedmatrix.cpp:5-37: raw2true o true2raw = id for a 9x7 pair, both orientations."""
    for (m, n) in ((9, 7), (7, 9)):
        seen = set()
        for ti in range(n + 1):
            for tj in range(m + 1):
                ri, rj = o.true2raw(ti, tj, m, n)
                assert o.raw2true(ri, rj, m, n) == (ti, tj)
                seen.add((ri, rj))
        assert len(seen) == (m + 1) * (n + 1)


def test_skewed_equals_plain_cells():
    """test/test_skewedmatrix.cpp:39-66: every cell of the two SMTs agrees for GGTTGACTA / TGTTACG."""
    for x, y in (("GGTTGACTA", "TGTTACG"), ("TGTTACG", "GGTTGACTA")):
        assert (o.matrix(x, y, mode=o.MODE_SAT_U8) == o.matrix(x, y, mode=o.MODE_EXACT)).all()


def test_make_string_range_values():
    """plocalaligner.cpp:44-67; values from SURVEY §8a-9 (probed on the reference)."""
    assert o.make_string_range(4, 125, 4980, 2.0) == [(0, 1432), (1182, 2614), (2364, 3796), (3546, 4980)]
    r17 = o.make_string_range(17, 125, 4980, 2.0)
    assert r17[0] == (0, 528) and r17[1] == (278, 806) and r17[-1] == (4448, 4980) and len(r17) == 17
    assert o.make_string_range(1, 10, 100, 2.0) == [(0, 100)]
    assert o.make_string_range(4, 100, 150, 2.0) == -1  # assert(overlaplength <= piecelength) :52


def test_saturate():
    """similaritymatrix.cpp:376-384."""
    assert [o.saturate(v) for v in (-1.0, 0.0, 2.9, 3.0, 255.0, 255.5, 1000.0)] == [0, 0, 2, 3, 255, 255, 255]


# ---- golden vectors dumped from the compiled reference ----------------------------------------------------
def test_golden_files_intact():
    with open(os.path.join(GOLDEN, "SHA256SUMS")) as f:
        for line in f:
            h, rel = line.split()
            with open(os.path.join(GOLDEN, rel), "rb") as g:
                assert hashlib.sha256(g.read()).hexdigest() == h, rel


@pytest.mark.parametrize("name,mode,npiece", [("data_small_sw_skewed.csv", o.MODE_SAT_U8, 0), ("data_small_p4.csv", o.MODE_SAT_U8, 4),
                                              ("data_small_p17.csv", o.MODE_SAT_U8, 17), ("data_small_sw_float.csv", o.MODE_EXACT, 0)])
def test_data_small_against_reference(data_small, name, mode, npiece):
    """BASELINE configs 1 and 2: all 1170 data_small reads, (score, pos, consensus) identical to the reference."""
    ref, truth = data_small
    gold = read_golden_csv(name)
    assert len(gold) == len(truth) == 1170
    step = 1 if mode == o.MODE_SAT_U8 else 6  # the int32 EXACT fill is slower; every 6th read keeps the CPU suite short
    for (idx, _, seq, _), g in list(zip(truth, gold))[::step]:
        r = o.align_chunked(seq, ref, npiece, 2.0, mode=mode) if npiece else o.align(seq, ref, mode=mode)
        assert (r["score"], r["pos"], r["cx"], r["cy"]) == (g["score"], g["pos"], g["cx"], g["cy"]), idx


def test_data_small_scores_saturate(data_small):
    """every data_small read scores exactly 255 on the u8 path."""
    assert all(g["score"] == 255 for g in read_golden_csv("data_small_sw_skewed.csv"))


def test_random_pairs_against_reference(random_pairs):
    for c in random_pairs:
        sc = c["scoring"]
        mode = o.MODE_SAT_U8 if c["smt"] == 0 else o.MODE_EXACT
        kw = dict(mode=mode, match=sc["match"], mismatch=sc["mismatch"], gap=sc["gap"])
        r = o.align_chunked(c["x"], c["y"], c["npiece"], c["ratio"], **kw) if c["npiece"] else o.align(c["x"], c["y"], **kw)
        assert (r["score"], r["pos"], r["cx"], r["cy"]) == (c["score"], c["pos"], c["cx"], c["cy"]), (len(c["x"]), len(c["y"]), c["smt"], sc, c["npiece"])


def test_c3_sample_against_reference(c3_sample):
    ref = synth.c3_reference(c3_sample["ref_len"])
    assert hashlib.sha256(ref.encode()).hexdigest() == c3_sample["ref_sha256"]
    for e in c3_sample["reads"][:4]:  # 150 x 10^6 int32 cells each; 4 keep the CPU suite short
        r = o.align(e["x"], ref, mode=o.MODE_SAT_U8)
        assert (r["score"], r["pos"], r["cx"], r["cy"]) == (e["score"], e["pos"], e["cx"], e["cy"])


def test_c4_sample_against_reference(c4_sample):
    table = synth.blosum62_table()
    for e in c4_sample["entries"]:
        r = o.align(e["x"], c4_sample["query"], mode=o.MODE_EXACT, table=table, gap=c4_sample["gap"])
        assert (r["score"], r["pos"], r["cx"], r["cy"]) == (e["score"], e["pos"], e["cx"], e["cy"])


@pytest.mark.skipif(o.ref() is None or not os.path.isdir("/root/reference"), reason="compiled reference not available")
def test_live_reference_random_cells():
    """Cell-by-cell: restatement == reference operator()(row, col) on fresh random non-square shapes."""
    rng = np.random.default_rng(7)
    for _ in range(60):
        m, n = int(rng.integers(1, 90)), int(rng.integers(1, 140))
        if m == n:
            n += 1
        x = "".join(rng.choice(list("ACGT"), size=m))
        y = "".join(rng.choice(list("ACGT"), size=n))
        for smt, mode in ((0, o.MODE_SAT_U8), (1, o.MODE_EXACT)):
            want = o.ref_matrix(x, y, smt=smt, scoring_kind=1, match=4, mismatch=-2, gap=3).astype(np.int32)
            got = o.matrix(x, y, mode=mode, match=4, mismatch=-2, gap=3)
            assert (want == got).all(), (m, n, smt)


@pytest.mark.skipif(o.ref() is None or not os.path.isdir("/root/reference"), reason="compiled reference not available")
def test_live_reference_saturating_scores():
    """Saturation with a large match score (hits 255 within a few rows) and scoring clamp sat()."""
    rng = np.random.default_rng(8)
    for _ in range(40):
        m, n = int(rng.integers(20, 70)), int(rng.integers(80, 200))
        y = "".join(rng.choice(list("ACGT"), size=n))
        s = int(rng.integers(0, n - m))
        x = y[s:s + m]
        want = o.ref_align(x, y, smt=0, scoring_kind=1, match=40, mismatch=-7, gap=9)
        got = o.align(x, y, mode=o.MODE_SAT_U8, match=40, mismatch=-7, gap=9)
        assert (got["score"], got["pos"], got["cx"], got["cy"]) == (want["score"], want["pos"], want["cx"], want["cy"])
        assert got["score"] == 255


@pytest.mark.skipif(o.ref() is None or not os.path.isdir("/root/reference"), reason="compiled reference not available")
def test_live_reference_asymmetric_table_orientation():
    """A scoring callback with fn(a, b) != fn(b, a): pins which argument is the x (row) character and which the y
    (column) character in the restatement — plain SWAligner and the chunked aligner (whose final alignment uses the
    DEFAULT scoring) against the compiled reference.  The query-stationary CUDA kernels compute the
    matrix transposed and are checked against the oracle, so this orientation must be the reference's."""
    rng = np.random.default_rng(12)
    alpha = list("ARNDCQEGHILKMFPSTWYV")
    table = rng.integers(-6, 5, size=(256, 256)).astype(np.int32)
    np.fill_diagonal(table, rng.integers(2, 9, size=256))
    assert (table != table.T).any()
    checked = 0
    for _ in range(40):
        m, n = int(rng.integers(5, 80)), int(rng.integers(30, 160))
        if m == n:
            n += 1
        y = "".join(rng.choice(alpha, size=n))
        s0 = int(rng.integers(0, max(1, n - m)))
        x = list((y * 4)[s0:s0 + m])
        for q in range(len(x)):
            if rng.random() < 0.3:
                x[q] = str(rng.choice(alpha))
        x = "".join(x)
        got = o.align(x, y, mode=o.MODE_EXACT, table=table, gap=3)
        if got["score"] == 0:
            continue
        want = o.ref_align(x, y, smt=1, scoring_kind=2, table=table, gap=3)
        assert (got["score"], got["pos"], got["cx"], got["cy"]) == (want["score"], want["pos"], want["cx"], want["cy"]), (m, n)
        H = o.ref_matrix(x, y, smt=1, scoring_kind=2, table=table, gap=3).astype(np.int32)
        assert (H == o.matrix(x, y, mode=o.MODE_EXACT, table=table, gap=3)).all()
        checked += 1
        if n > 4 * m and m >= 10:
            try:
                wantc = o.ref_align(x, y, smt=1, scoring_kind=2, table=table, gap=3, npiece=3, ratio=1.5)
            except AssertionError:
                continue
            gotc = o.align_chunked(x, y, 3, 1.5, mode=o.MODE_EXACT, table=table, gap=3)
            if gotc.get("err", 0) == 0 and gotc["score"] > 0:
                assert (gotc["score"], gotc["pos"], gotc["cx"], gotc["cy"]) == (wantc["score"], wantc["pos"], wantc["cx"], wantc["cy"]), (m, n, "chunked")
    assert checked >= 25


# ---- the linear-memory restatement (oracle/sw_oracle_linear.c) -----------------------------------------------
def _same(a, b):
    return (a["score"], a["pos"], a["cx"], a["cy"], a["end"], a["err"]) == (b["score"], b["pos"], b["cx"], b["cy"], b["end"], b["err"])


def test_linear_oracle_equals_full_matrix_oracle():
    """Differential test: anti-diagonal / three-diagonal restatement == full-matrix restatement on random shapes
    (both orientations, related and unrelated sequences, saturating and zero-gap scorings, a random asymmetric table,
    the all-zero case), incl. shapes large enough for several diagonal checkpoints and window doublings."""
    rng = np.random.default_rng(2025)
    table = rng.integers(-6, 9, size=(256, 256)).astype(np.int32)
    for k in range(260):
        m, n = int(rng.integers(1, 200)), int(rng.integers(1, 600))
        if m == n:
            n += 1
        if k % 7 == 0:
            m, n = n, m
        y = "".join(rng.choice(list("ACGT"), size=n))
        if k % 2 and m < n:
            s = int(rng.integers(0, n - m))
            x = "".join(c if rng.random() > 0.1 else str(rng.choice(list("ACGT"))) for c in y[s:s + m])
        else:
            x = "".join(rng.choice(list("ACGT"), size=m))
        for mode in (o.MODE_SAT_U8, o.MODE_EXACT):
            for (ma, mi, g) in ((3, -3, 2), (40, -7, 9), (5, -1, 0)):
                kw = dict(mode=mode, match=ma, mismatch=mi, gap=g)
                assert _same(o.align(x, y, **kw), o.align(x, y, linear=True, **kw)), (k, m, n, mode, ma, mi, g)
        assert _same(o.align(x, y, mode=o.MODE_EXACT, table=table, gap=3), o.align(x, y, mode=o.MODE_EXACT, table=table, gap=3, linear=True))
    assert o.align("AAAA", "CCCCCC", linear=True)["err"] == -2 == o.align("AAAA", "CCCCCC")["err"]
    ref = synth.c3_reference(40_000, seed=5)
    for x in synth.mutated_reads(ref, 2, 3_000, seed=6, sub=0.02, ins=0.002, dele=0.002):
        for mode in (o.MODE_SAT_U8, o.MODE_EXACT):
            assert _same(o.align(x, ref, mode=mode), o.align(x, ref, mode=mode, linear=True))


def test_linear_oracle_against_reference_goldens(data_small, random_pairs, c3_sample, c4_sample):
    """The linear-memory restatement against the vectors dumped from the compiled reference: data_small (both SMTs),
    the random pairs, the whole C3 sample (24 reads x 1 Mbp — the full-matrix oracle only affords 4) and the C4 sample."""
    ref, truth = data_small
    for name, mode, step in (("data_small_sw_skewed.csv", o.MODE_SAT_U8, 3), ("data_small_sw_float.csv", o.MODE_EXACT, 3)):
        gold = read_golden_csv(name)
        for (idx, _, seq, _), g in list(zip(truth, gold))[::step]:
            r = o.align(seq, ref, mode=mode, linear=True)
            assert (r["score"], r["pos"], r["cx"], r["cy"]) == (g["score"], g["pos"], g["cx"], g["cy"]), (name, idx)
    for c in random_pairs:
        if c["npiece"]:
            continue
        sc = c["scoring"]
        r = o.align(c["x"], c["y"], mode=o.MODE_SAT_U8 if c["smt"] == 0 else o.MODE_EXACT, match=sc["match"], mismatch=sc["mismatch"], gap=sc["gap"], linear=True)
        assert (r["score"], r["pos"], r["cx"], r["cy"]) == (c["score"], c["pos"], c["cx"], c["cy"])
    y = synth.c3_reference(c3_sample["ref_len"])
    for e in c3_sample["reads"]:
        r = o.align(e["x"], y, mode=o.MODE_SAT_U8, linear=True)
        assert (r["score"], r["pos"], r["cx"], r["cy"]) == (e["score"], e["pos"], e["cx"], e["cy"])
    table = synth.blosum62_table()
    for e in c4_sample["entries"]:
        r = o.align(e["x"], c4_sample["query"], mode=o.MODE_EXACT, table=table, gap=c4_sample["gap"], linear=True)
        assert (r["score"], r["pos"], r["cx"], r["cy"]) == (e["score"], e["pos"], e["cx"], e["cy"])


@pytest.mark.skipif(o.ref() is None or not os.path.isdir("/root/reference"), reason="compiled reference not available")
def test_linear_oracle_against_live_reference():
    """Fresh shapes straight against the compiled reference (both SMTs, custom scoring)."""
    rng = np.random.default_rng(91)
    for _ in range(40):
        m, n = int(rng.integers(5, 120)), int(rng.integers(20, 400))
        if m == n:
            n += 1
        y = "".join(rng.choice(list("ACGT"), size=n))
        s0 = int(rng.integers(0, max(1, n - m)))
        x = "".join(c if rng.random() > 0.15 else str(rng.choice(list("ACGT"))) for c in (y * 3)[s0:s0 + m])
        for smt, mode in ((0, o.MODE_SAT_U8), (1, o.MODE_EXACT)):
            got = o.align(x, y, mode=mode, match=4, mismatch=-2, gap=3, linear=True)
            if got["score"] == 0:
                continue
            want = o.ref_align(x, y, smt=smt, scoring_kind=1, match=4, mismatch=-2, gap=3)
            assert (got["score"], got["pos"], got["cx"], got["cy"]) == (want["score"], want["pos"], want["cx"], want["cy"]), (m, n, smt)


def test_linear_goldens_reproducible():
    """tests/golden/c3_10k.json and c5_full.json (make_golden_linear.py) are what the current oracle computes: the
    inputs regenerate from their seeds (sha256) and a sample of the rows is recomputed (C5 at full size is 200 s per
    alignment, so only the C3 rows are recomputed here; the C5 file is checked for shape and seed consistency)."""
    import json
    with open(os.path.join(GOLDEN, "c3_10k.json")) as f:
        c3 = json.load(f)
    ref = synth.c3_reference(c3["ref_len"])
    assert hashlib.sha256(ref.encode()).hexdigest() == c3["ref_sha256"]
    reads = synth.c3_reads(ref, c3["n_reads"])
    assert hashlib.sha256("".join(reads).encode()).hexdigest() == c3["reads_sha256"] and len(c3["rows"]) == 10_000
    for i in range(0, 10_000, 1250):
        w = o.align(reads[i], ref, mode=o.MODE_SAT_U8, linear=True)
        d = hashlib.sha256((w["cx"] + "|" + w["cy"]).encode()).hexdigest()[:16]
        assert [w["score"], w["pos"], w["end"][0], w["end"][1], len(w["cx"]), d] == c3["rows"][i], i
    with open(os.path.join(GOLDEN, "c5_full.json")) as f:
        c5 = json.load(f)
    assert c5["ref_len"] == 51_000_000 and c5["n_reads"] == len(c5["exact"]) == len(c5["sat_u8"]) >= 2
    assert all(e["score"] == 255 for e in c5["sat_u8"]) and all(25_000 < e["score"] <= 30_000 for e in c5["exact"])
    for e in c5["exact"][:2]:
        assert hashlib.sha256(e["cx"].encode()).hexdigest() == e["cx_sha256"] and len(e["cx"]) == e["len"]
