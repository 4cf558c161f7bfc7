"""ctypes bindings for the CPU oracle (oracle/sw_oracle.c) and the compiled reference (oracle/_ref).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MODE_SAT_U8, MODE_EXACT = 0, 1
_lib = None
_ref = None


def build_oracle(force=False):
    """liboracle.so = the full-matrix restatement (sw_oracle.c) + the linear-memory one (sw_oracle_linear.c)."""
    srcs = [os.path.join(HERE, "sw_oracle.c"), os.path.join(HERE, "sw_oracle_linear.c")]
    lib = os.path.join(HERE, "liboracle.so")
    if force or not os.path.isfile(lib) or any(os.path.getmtime(lib) < os.path.getmtime(s) for s in srcs):
        # x86-64-v3 (AVX2): the column loops of the linear oracle vectorise; the file built here also runs on the GPU box
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-fPIC", "-shared", "-std=c11", "-Wall", "-Wextra", "-o", lib] + srcs)
    return lib


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle())
        _lib.swo_align.restype = C.c_int64
        _lib.swo_align_linear.restype = C.c_int64
        _lib.swo_align_chunked.restype = C.c_int64
        _lib.swo_traceback.restype = C.c_int64
    return _lib


def _u8(s):
    if isinstance(s, str):
        s = s.encode("latin-1")
    return np.frombuffer(bytes(s), dtype=np.uint8)


def _ptr(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def default_table(match=3, mismatch=-3):
    t = np.full((256, 256), mismatch, dtype=np.int32)
    np.fill_diagonal(t, match)
    return t


def saturate(v):
    return int(lib().swo_saturate(C.c_float(v)))


def align(x, y, mode=MODE_SAT_U8, match=3, mismatch=-3, gap=2, table=None, cap=None, linear=False):
    """SWAligner<SMT>(x, y, fn, gap).calculateScore() restated.  Returns dict(score,pos,cx,cy,end).
    linear=True: the linear-memory restatement (sw_oracle_linear.c) — same result, O(m + window) memory."""
    xs, ys = _u8(x), _u8(y)
    m, n = len(xs), len(ys)
    cap = cap or (m + n + 2)
    cx, cy = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
    score, pos = C.c_int32(0), C.c_uint32(0)
    ex, ey = C.c_int64(0), C.c_int64(0)
    if mode == MODE_SAT_U8:
        p0, p1, g, tp = saturate(match), saturate(-mismatch), saturate(gap), None
    else:
        t = np.ascontiguousarray(table if table is not None else default_table(match, mismatch), dtype=np.int32)
        p0, p1, g, tp = 0, 0, int(gap), _ptr(t)
    fn = lib().swo_align_linear if linear else lib().swo_align
    ln = fn(mode, _ptr(xs), C.c_int64(m), _ptr(ys), C.c_int64(n), p0, p1, tp, g,
            C.byref(score), C.byref(pos), C.byref(ex), C.byref(ey), _ptr(cx), _ptr(cy), C.c_int64(cap))
    if ln < 0:
        return dict(score=score.value, pos=0, cx="", cy="", end=(ex.value, ey.value), err=int(ln))
    return dict(score=score.value, pos=pos.value, cx=cx[:ln].tobytes().decode("latin-1"),
                cy=cy[:ln].tobytes().decode("latin-1"), end=(ex.value, ey.value), err=0)


def align_chunked(x, y, npiece, ratio, mode=MODE_SAT_U8, match=3, mismatch=-3, gap=2, table=None):
    """Serial OMPParallelLocalAligner<SMT, SWAligner<SMT>>(x, y, npiece, ratio, fn, gap) restated."""
    xs, ys = _u8(x), _u8(y)
    m, n = len(xs), len(ys)
    cap = m + n + 2
    cx, cy = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
    score, pos, bp = C.c_int32(0), C.c_uint32(0), C.c_int(0)
    if mode == MODE_SAT_U8:
        p0, p1, g, tp = saturate(match), saturate(-mismatch), saturate(gap), None
    else:
        t = np.ascontiguousarray(table if table is not None else default_table(match, mismatch), dtype=np.int32)
        p0, p1, g, tp = 0, 0, int(gap), _ptr(t)
    ln = lib().swo_align_chunked(mode, _ptr(xs), C.c_int64(m), _ptr(ys), C.c_int64(n), int(npiece), C.c_float(ratio),
                                 p0, p1, tp, g, C.byref(score), C.byref(pos), C.byref(bp), _ptr(cx), _ptr(cy), C.c_int64(cap))
    if ln < 0:
        return dict(score=score.value, pos=0, cx="", cy="", piece=bp.value, err=int(ln))
    return dict(score=score.value, pos=pos.value, cx=cx[:ln].tobytes().decode("latin-1"),
                cy=cy[:ln].tobytes().decode("latin-1"), piece=bp.value, err=0)


def matrix(x, y, mode=MODE_SAT_U8, match=3, mismatch=-3, gap=2, table=None):
    xs, ys = _u8(x), _u8(y)
    m, n = len(xs), len(ys)
    H = np.zeros((m + 1, n + 1), np.int32)
    if mode == MODE_SAT_U8:
        lib().swo_matrix(mode, _ptr(xs), C.c_int64(m), _ptr(ys), C.c_int64(n), saturate(match), saturate(-mismatch), None, saturate(gap), _ptr(H))
    else:
        t = np.ascontiguousarray(table if table is not None else default_table(match, mismatch), dtype=np.int32)
        lib().swo_matrix(mode, _ptr(xs), C.c_int64(m), _ptr(ys), C.c_int64(n), 0, 0, _ptr(t), int(gap), _ptr(H))
    return H


def argmax(H, mode):
    H = np.ascontiguousarray(H, np.int32)
    m, n = H.shape[0] - 1, H.shape[1] - 1
    ix, iy, mx = C.c_int64(0), C.c_int64(0), C.c_int32(0)
    f = lib().swo_argmax_skewed if mode == MODE_SAT_U8 else lib().swo_argmax_colmajor
    f(_ptr(H), C.c_int64(m), C.c_int64(n), C.byref(ix), C.byref(iy), C.byref(mx))
    return ix.value, iy.value, mx.value


def true2raw(ti, tj, m, n):
    ri, rj = C.c_int64(0), C.c_int64(0)
    lib().swo_true2raw(C.c_int64(ti), C.c_int64(tj), C.c_int64(m), C.c_int64(n), C.byref(ri), C.byref(rj))
    return ri.value, rj.value


def raw2true(ri, rj, m, n):
    ti, tj = C.c_int64(0), C.c_int64(0)
    lib().swo_raw2true(C.c_int64(ri), C.c_int64(rj), C.c_int64(m), C.c_int64(n), C.byref(ti), C.byref(tj))
    return ti.value, tj.value


def make_string_range(npiece, shortlen, longlen, ratio):
    left = np.zeros(max(npiece, 1), np.int64)
    right = np.zeros(max(npiece, 1), np.int64)
    k = lib().swo_make_string_range(int(npiece), C.c_int64(shortlen), C.c_int64(longlen), C.c_float(ratio), _ptr(left), _ptr(right))
    if k < 0:
        return k
    return [(int(left[i]), int(right[i])) for i in range(k)]


# ----------------------------------------------------------------------------------------------
# Compiled reference (oracle/_ref/libref_aligner.so).  Optional: absent => ref() returns None.
# ----------------------------------------------------------------------------------------------
def ref():
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libref_aligner.so")
        if not os.path.isfile(path):
            try:
                import importlib.util
                spec = importlib.util.spec_from_file_location("build_ref", os.path.join(HERE, "build_ref.py"))
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                path = mod.build()
            except Exception:
                return None
        _ref = C.CDLL(path)
        _ref.ref_bench_reads.restype = C.c_double
    return _ref


def ref_align(x, y, smt=0, scoring_kind=0, match=3.0, mismatch=-3.0, gap=2.0, table=None, npiece=0, ratio=0.0):
    """Run the reference's own SWAligner / OMPParallelLocalAligner.  smt 0 = Skewed(u8), 1 = plain f32."""
    r = ref()
    xs, ys = _u8(x), _u8(y)
    cap = len(xs) + len(ys) + 2
    cx, cy = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
    score, pos, ln, us = C.c_float(0), C.c_uint(0), C.c_int(0), C.c_float(0)
    tp = None
    if table is not None:
        tf = np.ascontiguousarray(table, dtype=np.float32)
        tp = _ptr(tf)
    rc = r.ref_align(int(smt), _ptr(xs), C.c_int64(len(xs)), _ptr(ys), C.c_int64(len(ys)), int(scoring_kind),
                     C.c_float(match), C.c_float(mismatch), C.c_float(gap), tp, int(npiece), C.c_float(ratio),
                     C.byref(score), C.byref(pos), _ptr(cx), _ptr(cy), cap, C.byref(ln), C.byref(us))
    assert rc == 0, rc
    return dict(score=int(score.value), pos=pos.value, cx=cx[:ln.value].tobytes().decode("latin-1"),
                cy=cy[:ln.value].tobytes().decode("latin-1"), iterate_us=us.value)


def ref_matrix(x, y, smt=0, scoring_kind=0, match=3.0, mismatch=-3.0, gap=2.0, table=None):
    r = ref()
    xs, ys = _u8(x), _u8(y)
    out = np.zeros((len(xs) + 1, len(ys) + 1), np.float32)
    tp = None
    if table is not None:
        tf = np.ascontiguousarray(table, dtype=np.float32)
        tp = _ptr(tf)
    r.ref_matrix(int(smt), _ptr(xs), C.c_int64(len(xs)), _ptr(ys), C.c_int64(len(ys)), int(scoring_kind),
                 C.c_float(match), C.c_float(mismatch), C.c_float(gap), tp, _ptr(out))
    return out


def ref_make_string_range(npiece, shortlen, longlen, ratio):
    r = ref()
    left, right = np.zeros(npiece, np.int64), np.zeros(npiece, np.int64)
    k = r.ref_make_string_range(int(npiece), C.c_int64(shortlen), C.c_int64(longlen), C.c_float(ratio), _ptr(left), _ptr(right))
    return [(int(left[i]), int(right[i])) for i in range(k)]


def ref_bench_reads(reads, y, smt=0, npiece=0, ratio=0.0, nthreads=1):
    """Time the reference aligner over a list of reads (harness-level OpenMP over reads).
    Returns (wall_seconds, iterate_us_sum, pos[], score[])."""
    r = ref()
    blobs = [bytes(_u8(s)) for s in reads]
    offs = np.zeros(len(blobs) + 1, np.int64)
    offs[1:] = np.cumsum([len(b) for b in blobs])
    blob = np.frombuffer(b"".join(blobs), dtype=np.uint8)
    ys = _u8(y)
    pos, score = np.zeros(len(blobs), np.uint32), np.zeros(len(blobs), np.float32)
    us = C.c_double(0)
    wall = r.ref_bench_reads(int(smt), _ptr(blob), _ptr(offs), C.c_int64(len(blobs)), _ptr(ys), C.c_int64(len(ys)),
                             int(npiece), C.c_float(ratio), int(nthreads), C.byref(us), _ptr(pos), _ptr(score))
    return wall, us.value, pos, score


# ----------------------------------------------------------------------------------------------
# The reference's own -DUSEOMP build (oracle/_ref/libref_aligner_omp.so): TIMING ONLY.
# ----------------------------------------------------------------------------------------------
_ref_omp = None


def ref_omp():
    global _ref_omp
    if _ref_omp is None:
        path = os.path.join(HERE, "_ref", "libref_aligner_omp.so")
        if not os.path.isfile(path):
            try:
                import importlib.util
                spec = importlib.util.spec_from_file_location("build_ref", os.path.join(HERE, "build_ref.py"))
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                path = mod.build(useomp=True)
            except Exception:
                return None
        _ref_omp = C.CDLL(path)
        _ref_omp.ref_omp_chunked_bench.restype = C.c_double
        _ref_omp.ref_omp_finegrain.restype = C.c_double
    return _ref_omp


def ref_omp_chunked_bench(reads, y, npiece, ratio=2.0):
    """the reference's USEOMP OMPParallelLocalAligner<Skewed>(read, ref, npiece, ratio) per read
    (threads = npiece inside).  Returns (wall_seconds, iterate_us_sum).  Results are racy and are not returned."""
    r = ref_omp()
    blobs = [bytes(_u8(s)) for s in reads]
    offs = np.zeros(len(blobs) + 1, np.int64)
    offs[1:] = np.cumsum([len(b) for b in blobs])
    blob = np.frombuffer(b"".join(blobs), dtype=np.uint8)
    ys = _u8(y)
    us = C.c_double(0)
    wall = r.ref_omp_chunked_bench(_ptr(blob), _ptr(offs), C.c_int64(len(blobs)), _ptr(ys), C.c_int64(len(ys)), int(npiece), C.c_float(ratio), C.byref(us))
    return wall, us.value


def ref_omp_finegrain(x, y, finegrain_type=1, nthreads=1):
    """SWAligner<Similarity_Matrix> with the driver-set fine-grain fields (omp_sw_solve_small.cpp:164-189).
    Returns dict(wall, score, pos, iterate_us, adsum_us)."""
    r = ref_omp()
    xs, ys = _u8(x), _u8(y)
    score, pos = C.c_float(0), C.c_uint(0)
    t = np.zeros(2, np.float32)
    wall = r.ref_omp_finegrain(_ptr(xs), C.c_int64(len(xs)), _ptr(ys), C.c_int64(len(ys)), int(finegrain_type), int(nthreads), C.byref(score), C.byref(pos), _ptr(t))
    return dict(wall=wall, score=int(score.value), pos=pos.value, iterate_us=float(t[0]), adsum_us=float(t[1]))
