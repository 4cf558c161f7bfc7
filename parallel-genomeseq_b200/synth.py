"""Seeded synthetic workloads of the shapes BASELINE.json names.

Host-side data preparation only (the counterpart of the reference's py/ompfg_data_prep.py:31-122,
which samples reads from a reference and mutates them).  Pure numpy; no GPU, no oracle.

  C3  batched read mapping : iid uniform ACGT reference, reads = windows of it with 1 % substitutions,
                             0.1 % insertions, 0.1 % deletions, padded/trimmed to read_len.
  C4  protein DB search    : residues iid over the 20 amino acids; DB length = clamp(round(exp(N(5.65,
                             0.65))), 10, 5000) (Swiss-Prot-like, mean ~350); BLOSUM62 table.
  C5  long pair            : same generator as C3 with 2 % substitutions / 0.2 % indels.
"""
import numpy as np

DNA = np.frombuffer(b"ACGT", dtype=np.uint8)
AMINO = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)

_BLOSUM62_ORDER = "ARNDCQEGHILKMFPSTWYVBZX*"
_BLOSUM62_ROWS = """
 4 -1 -2 -2  0 -1 -1  0 -2 -1 -1 -1 -1 -2 -1  1  0 -3 -2  0 -2 -1  0 -4
-1  5  0 -2 -3  1  0 -2  0 -3 -2  2 -1 -3 -2 -1 -1 -3 -2 -3 -1  0 -1 -4
-2  0  6  1 -3  0  0  0  1 -3 -3  0 -2 -3 -2  1  0 -4 -2 -3  3  0 -1 -4
-2 -2  1  6 -3  0  2 -1 -1 -3 -4 -1 -3 -3 -1  0 -1 -4 -3 -3  4  1 -1 -4
 0 -3 -3 -3  9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1 -3 -3 -2 -4
-1  1  0  0 -3  5  2 -2  0 -3 -2  1  0 -3 -1  0 -1 -2 -1 -2  0  3 -1 -4
-1  0  0  2 -4  2  5 -2  0 -3 -3  1 -2 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4
 0 -2  0 -1 -3 -2 -2  6 -2 -4 -4 -2 -3 -3 -2  0 -2 -2 -3 -3 -1 -2 -1 -4
-2  0  1 -1 -3  0  0 -2  8 -3 -3 -1 -2 -1 -2 -1 -2 -2  2 -3  0  0 -1 -4
-1 -3 -3 -3 -1 -3 -3 -4 -3  4  2 -3  1  0 -3 -2 -1 -3 -1  3 -3 -3 -1 -4
-1 -2 -3 -4 -1 -2 -3 -4 -3  2  4 -2  2  0 -3 -2 -1 -2 -1  1 -4 -3 -1 -4
-1  2  0 -1 -3  1  1 -2 -1 -3 -2  5 -1 -3 -1  0 -1 -3 -2 -2  0  1 -1 -4
-1 -1 -2 -3 -1  0 -2 -3 -2  1  2 -1  5  0 -2 -1 -1 -1 -1  1 -3 -1 -1 -4
-2 -3 -3 -3 -2 -3 -3 -3 -1  0  0 -3  0  6 -4 -2 -2  1  3 -1 -3 -3 -1 -4
-1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4  7 -1 -1 -4 -3 -2 -2 -1 -2 -4
 1 -1  1  0 -1  0  0  0 -1 -2 -2  0 -1 -2 -1  4  1 -3 -2 -2  0  0  0 -4
 0 -1  0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1  1  5 -2 -2  0 -1 -1  0 -4
-3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1  1 -4 -3 -2 11  2 -3 -4 -3 -2 -4
-2 -2 -2 -3 -2 -1 -2 -3  2 -1 -1 -2 -1  3 -3 -2 -2  2  7 -1 -3 -2 -1 -4
 0 -3 -3 -3 -1 -2 -2 -3 -3  3  1 -2  1 -1 -2 -2  0 -3 -1  4 -3 -2 -1 -4
-2 -1  3  4 -3  0  1 -1  0 -3 -4  0 -3 -3 -2  0 -1 -4 -3 -3  4  1 -1 -4
-1  0  0  1 -3  3  4 -2  0 -3 -3  1 -1 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4
 0 -1 -1 -1 -2 -1 -1 -1 -1 -1 -1 -1 -1 -1 -2  0  0 -2 -1 -1 -1 -1 -1 -4
-4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4  1
"""


def blosum62_table():
    """256x256 int32 table[a][b] = BLOSUM62(a, b) for upper-case residue bytes; -4 for anything else.

    This is the tabulated form of the scoring callback std::function<float(const char&, const char&)>
    that SWAligner's constructor takes (smithwaterman.h:16-17).  The reference ships no BLOSUM data
    (SURVEY F4); the same table is handed to the reference and to the CUDA path in the parity tests.
    """
    rows = [[int(v) for v in line.split()] for line in _BLOSUM62_ROWS.strip().split("\n")]
    b = np.array(rows, dtype=np.int32)
    assert b.shape == (24, 24) and (b == b.T).all()
    t = np.full((256, 256), -4, dtype=np.int32)
    idx = [ord(c) for c in _BLOSUM62_ORDER]
    t[np.ix_(idx, idx)] = b
    return t


def match_table(match=3, mismatch=-3):
    """256x256 table of the reference's default callback a == b ? 3 : -3 (smithwaterman.cpp:8)."""
    t = np.full((256, 256), mismatch, dtype=np.int32)
    np.fill_diagonal(t, match)
    return t


def c3_reference(n=1_000_000, seed=22):
    rng = np.random.default_rng(seed)
    return DNA[rng.integers(0, 4, size=n)].tobytes().decode("ascii")


def mutated_reads(ref, count, read_len, seed, sub=0.01, ins=0.001, dele=0.001):
    """`count` reads of exactly read_len bases sampled uniformly from `ref` and mutated."""
    rng = np.random.default_rng(seed)
    r = np.frombuffer(ref.encode("ascii"), dtype=np.uint8)
    n = len(r)
    slack = max(8, int(read_len * (ins + dele) * 8) + 8)
    starts = rng.integers(0, n - read_len - slack + 1, size=count)
    out = []
    for s in starts:
        w = r[s:s + read_len + slack].copy()
        u = rng.random(len(w))
        subm = u < sub
        w[subm] = DNA[rng.integers(0, 4, size=int(subm.sum()))]
        keep = rng.random(len(w)) >= dele
        w = w[keep]
        insm = np.nonzero(rng.random(len(w)) < ins)[0]
        if len(insm):
            w = np.insert(w, insm, DNA[rng.integers(0, 4, size=len(insm))])
        out.append(w[:read_len].tobytes().decode("ascii"))
    return out


def mutated_reads_fast(ref_u8, count, read_len, seed, sub=0.01):
    """Vectorised read sampler for large batches (substitutions only): returns (count, read_len) uint8.

    Used by bench.py for the 10^5..10^6-read batches where the per-read python loop above is too slow;
    indels do not change the cell count or the kernel's work."""
    rng = np.random.default_rng(seed)
    n = len(ref_u8)
    starts = rng.integers(0, n - read_len + 1, size=count)
    idx = starts[:, None] + np.arange(read_len)[None, :]
    reads = ref_u8[idx]
    subm = rng.random(reads.shape) < sub
    reads[subm] = DNA[rng.integers(0, 4, size=int(subm.sum()))]
    return np.ascontiguousarray(reads)


def mutated_reads_vec(ref_u8, count, read_len, seed, sub=0.01, ins=0.001, dele=0.001):
    """Vectorised read sampler WITH indels for the 10^5..10^6-read batches bench.py times: (count, read_len) uint8.

    Same error model as mutated_reads (py/ompfg_data_prep.py:98-104 samples windows; the 1 % / 0.1 % / 0.1 % rates are
    SURVEY §8d's C3 workload): a window of the reference is walked base by base; a base is dropped with probability
    `dele`, otherwise emitted (substituted with probability `sub`) and followed, with probability `ins`, by one random
    base; the first read_len emitted symbols are the read."""
    rng = np.random.default_rng(seed)
    n = len(ref_u8)
    slack = max(8, int(read_len * (ins + dele) * 8) + 8)
    w = read_len + slack
    starts = rng.integers(0, n - w + 1, size=count)
    src = ref_u8[starts[:, None] + np.arange(w)[None, :]]
    subm = rng.random(src.shape) < sub
    src[subm] = DNA[rng.integers(0, 4, size=int(subm.sum()))]
    keep = rng.random(src.shape) >= dele
    insm = keep & (rng.random(src.shape) < ins)
    emitted = keep.astype(np.int32) + insm.astype(np.int32)          # symbols this source position contributes
    first = np.cumsum(emitted, axis=1) - emitted                     # output index of its first symbol
    out = np.zeros((count, read_len), dtype=np.uint8)
    rows = np.broadcast_to(np.arange(count)[:, None], src.shape)
    sel = keep & (first < read_len)
    out[rows[sel], first[sel]] = src[sel]
    sel = insm & (first + 1 < read_len)
    out[rows[sel], first[sel] + 1] = DNA[rng.integers(0, 4, size=int(sel.sum()))]
    assert (out != 0).all()
    return out


def c3_reads(ref, count, read_len=150, seed=23):
    return mutated_reads(ref, count, read_len, seed, sub=0.01, ins=0.001, dele=0.001)


def c4_queries(count=64, length=300, seed=24):
    rng = np.random.default_rng(seed)
    return [AMINO[rng.integers(0, 20, size=length)].tobytes().decode("ascii") for _ in range(count)]


def c4_database(count=500_000, seed=25):
    rng = np.random.default_rng(seed)
    lens = np.clip(np.rint(np.exp(rng.normal(5.65, 0.65, size=count))), 10, 5000).astype(np.int64)
    return [AMINO[rng.integers(0, 20, size=int(l))].tobytes().decode("ascii") for l in lens]


def c5_reference(n=51_000_000, seed=26):
    return c3_reference(n, seed)


def c5_reads(ref, count=16, read_len=10_000, seed=27):
    return mutated_reads(ref, count, read_len, seed, sub=0.02, ins=0.002, dele=0.002)
