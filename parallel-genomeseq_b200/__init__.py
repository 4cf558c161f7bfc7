"""parallel-genomeseq_b200 — B200-native Smith-Waterman engine behind the parallel-genomeseq aligner API.

The product is libswb200.so (hand-written sm_100a CUDA kernels behind the C ABI of include/swb200.h)
plus the C++ shims in cpp/ that mirror the reference's LocalAligner / ParallelLocalAligner classes.
This Python module is a thin ctypes binding of that C ABI for the tests and bench.py; it contains no
alignment logic and has NO CPU fallback: if the library or a CUDA device is missing it raises.
"""
import ctypes as C
import os

import numpy as np

from . import synth  # noqa: F401  (synthetic workloads of the BASELINE shapes)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libswb200.so")

MODE_SAT_U8, MODE_EXACT = 0, 1
FLAG_CONSENSUS = 1
ERR_NAMES = {-1: "SWB_ERR_CUDA", -2: "SWB_ERR_ARG", -3: "SWB_ERR_RANGE", -4: "SWB_ERR_SCORING", -5: "SWB_ERR_UNSUPPORTED", -6: "SWB_ERR_STATE"}

# every symbol include/swb200.h declares
EXPORTS = ["swb_device_count", "swb_create", "swb_destroy", "swb_last_error", "swb_version", "swb_set_scoring", "swb_set_scoring_match",
           "swb_set_reference", "swb_align_batch", "swb_batch_stage", "swb_batch_run", "swb_batch_fetch",
           "swb_batch_device_results", "swb_batch_rebind_reference", "swb_last_stats", "swb_make_string_range", "swb_matrix"]


class SwbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [("cells_reference", C.c_uint64), ("cells_executed", C.c_uint64), ("cells_pass2", C.c_uint64),
                ("kernel_launches", C.c_uint32), ("lanes_per_pair", C.c_uint32), ("rows_per_lane", C.c_uint32),
                ("block_steps", C.c_uint32), ("pass1_us", C.c_float), ("pass2_us", C.c_float),
                ("cols_per_step", C.c_uint32), ("kernel_kind", C.c_uint32)]


_lib = None


def load_library():
    """dlopen libswb200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("SWB_LIB_PATH") or LIB_PATH      # SWB_LIB_PATH: A/B runs of an alternative build of the same library
    if not os.path.isfile(path):
        raise ImportError(f"{path} is missing: run `python __graft_entry__.py build` (there is no CPU fallback)")
    lib = C.CDLL(path)
    lib.swb_last_error.restype = C.c_char_p
    lib.swb_version.restype = C.c_char_p
    lib.swb_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.swb_destroy.argtypes = [C.c_void_p]
    lib.swb_last_error.argtypes = [C.c_void_p]
    lib.swb_set_scoring.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_float]
    lib.swb_set_scoring_match.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float]
    lib.swb_set_reference.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.swb_align_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_uint,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                    C.c_void_p, C.c_void_p]
    lib.swb_batch_stage.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_uint, C.c_size_t]
    lib.swb_batch_run.argtypes = [C.c_void_p, C.c_void_p]
    lib.swb_batch_fetch.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    lib.swb_batch_device_results.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.swb_batch_rebind_reference.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.swb_last_stats.argtypes = [C.c_void_p, C.c_void_p]
    lib.swb_make_string_range.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_void_p]
    lib.swb_matrix.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    _lib = lib
    return lib


def make_string_range(npiece, shortlen, longlen, ratio):
    """_make_string_range (plocalaligner.cpp:44-67) through the C ABI; returns a list of (left, right) or raises."""
    lib = load_library()
    left = np.zeros(max(1, npiece), np.int64)
    right = np.zeros(max(1, npiece), np.int64)
    k = lib.swb_make_string_range(int(npiece), int(shortlen), int(longlen), float(ratio), left.ctypes.data, right.ctypes.data)
    if k < 0:
        raise SwbError(k, "_make_string_range precondition failed")
    return [(int(left[i]), int(right[i])) for i in range(k)]


def pack_sequences(seqs):
    """list of str/bytes (or a 2-D uint8 array) -> (blob uint8[], offsets uint64[n+1])."""
    if isinstance(seqs, np.ndarray) and seqs.ndim == 2:
        n, m = seqs.shape
        return np.ascontiguousarray(seqs, np.uint8).reshape(-1), (np.arange(n + 1, dtype=np.uint64) * np.uint64(m))
    bs = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in seqs]
    offs = np.zeros(len(bs) + 1, np.uint64)
    if bs:
        offs[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    return np.frombuffer(b"".join(bs), dtype=np.uint8).copy(), offs


class Engine:
    """One swb_ctx.  Mirrors the reference's constructor surface:
    scoring callback + gap (smithwaterman.h:14-17) -> set_scoring*, sequence_y -> set_reference,
    the per-read aligner loop (sw_solve_small.cpp:56-101) -> align()."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.swb_create(int(device), C.byref(h))
        if rc != 0:
            raise SwbError(rc, "swb_create failed: no usable CUDA device (there is no CPU fallback)")
        self.h = h
        self.n_seqs = 0
        self.cons_stride = 0
        self.flags = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.swb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise SwbError(rc, self.lib.swb_last_error(self.h).decode())

    def set_scoring_match(self, mode, match=3.0, mismatch=-3.0, gap=2.0):
        self._check(self.lib.swb_set_scoring_match(self.h, int(mode), float(match), float(mismatch), float(gap)))

    def set_scoring_table(self, mode, table, gap):
        t = np.ascontiguousarray(table, dtype=np.float32)
        assert t.shape == (256, 256)
        self._check(self.lib.swb_set_scoring(self.h, int(mode), t.ctypes.data, float(gap)))

    def set_reference(self, y):
        yb = y.encode("latin-1") if isinstance(y, str) else bytes(y)
        self._ref_keepalive = yb
        self._check(self.lib.swb_set_reference(self.h, yb, len(yb)))

    # ---- split API (inputs resident in HBM) ---------------------------------------------------------
    def stage(self, seqs, npiece=0, ratio=0.0, consensus=True, cons_stride=None):
        blob, offs = seqs if isinstance(seqs, tuple) else pack_sequences(seqs)
        n = len(offs) - 1
        if cons_stride is None:
            lens = np.diff(offs).astype(np.int64)
            cons_stride = int(2 * lens.max() + 64) if n else 64
        self._blob, self._offs = blob, offs
        self.n_seqs, self.cons_stride = n, int(cons_stride)
        self.flags = FLAG_CONSENSUS if consensus else 0
        self._check(self.lib.swb_batch_stage(self.h, blob.ctypes.data, offs.ctypes.data, n, int(npiece), float(ratio), self.flags, self.cons_stride))

    def rebind_reference(self, y):
        """Swap the reference under a batch staged in query-stationary mode (database search over many queries)."""
        yb = y.encode("latin-1") if isinstance(y, str) else bytes(y)
        self._ref_keepalive = yb
        self._check(self.lib.swb_batch_rebind_reference(self.h, yb, len(yb)))

    def run(self):
        us = C.c_float(0)
        self._check(self.lib.swb_batch_run(self.h, C.byref(us)))
        return us.value

    def fetch(self):
        n = self.n_seqs
        out = dict(score=np.zeros(n, np.int32), pos=np.zeros(n, np.uint32), end=np.zeros((n, 2), np.uint32),
                   len=np.zeros(n, np.uint32), flags=np.zeros(n, np.uint32))
        cx = cy = None
        if self.flags & FLAG_CONSENSUS:
            cx = np.zeros((n, self.cons_stride), np.uint8)
            cy = np.zeros((n, self.cons_stride), np.uint8)
        self._check(self.lib.swb_batch_fetch(self.h, out["score"].ctypes.data, out["pos"].ctypes.data, out["end"].ctypes.data,
                                             cx.ctypes.data if cx is not None else None, cy.ctypes.data if cy is not None else None,
                                             out["len"].ctypes.data, out["flags"].ctypes.data))
        out["cx_raw"], out["cy_raw"] = cx, cy
        return out

    def device_results(self):
        ds, dp = C.c_void_p(), C.c_void_p()
        self._check(self.lib.swb_batch_device_results(self.h, C.byref(ds), C.byref(dp)))
        return ds.value, dp.value

    def stats(self):
        s = Stats()
        self._check(self.lib.swb_last_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def matrix(self, x):
        """Dense H, (len(x)+1) x (len(reference)+1) int32 — Abstract_Similarity_Matrix::operator()(row, col)
        (similaritymatrix.h:13-24) through the device path; tests and small inputs only."""
        xb = x.encode("latin-1") if isinstance(x, str) else bytes(x)
        n = len(self._ref_keepalive)
        out = np.zeros((len(xb) + 1, n + 1), np.int32)
        self._check(self.lib.swb_matrix(self.h, xb, len(xb), out.ctypes.data))
        return out

    # ---- one-call API (host buffers in, host buffers out) ----------------------------------------------
    def align(self, seqs, npiece=0, ratio=0.0, consensus=True, cons_stride=None, decode=True):
        """Returns dict(score, pos, end, len, flags, cx, cy, device_us); cx/cy are lists of str (end -> start).
        decode=False skips the per-read Python string decoding and returns the raw arenas as cx_raw / cy_raw."""
        blob, offs = seqs if isinstance(seqs, tuple) else pack_sequences(seqs)
        n = len(offs) - 1
        if cons_stride is None:
            cons_stride = int(2 * np.diff(offs).astype(np.int64).max() + 64)
        flags = FLAG_CONSENSUS if consensus else 0
        # the consensus arenas (~100 MB for a large batch) are kept between calls of the same shape — re-allocating
        # them per call costs page faults inside the end-to-end time; cx_raw / cy_raw are therefore only valid
        # until the next align() of the same shape.  The small per-read arrays are fresh on every call.
        out = dict(score=np.zeros(n, np.int32), pos=np.zeros(n, np.uint32), end=np.zeros((n, 2), np.uint32),
                   len=np.zeros(n, np.uint32), flags=np.zeros(n, np.uint32))
        key = (n, cons_stride)
        if consensus and getattr(self, "_arena_key", None) != key:
            self._arena_key = key
            self._arena = (np.zeros((n, cons_stride), np.uint8), np.zeros((n, cons_stride), np.uint8))
        cx, cy = self._arena if consensus else (None, None)
        us = C.c_float(0)
        self._check(self.lib.swb_align_batch(self.h, blob.ctypes.data, offs.ctypes.data, n, int(npiece), float(ratio), flags,
                                             out["score"].ctypes.data, out["pos"].ctypes.data, out["end"].ctypes.data,
                                             cx.ctypes.data if consensus else None, cy.ctypes.data if consensus else None,
                                             out["len"].ctypes.data, cons_stride, out["flags"].ctypes.data, C.byref(us)))
        self.n_seqs, self.cons_stride, self.flags = n, cons_stride, flags
        out["device_us"] = us.value
        out["cx_raw"], out["cy_raw"] = cx, cy
        if consensus and decode:
            out["cx"] = [cx[i, :min(out["len"][i], cons_stride)].tobytes().decode("latin-1") for i in range(n)]
            out["cy"] = [cy[i, :min(out["len"][i], cons_stride)].tobytes().decode("latin-1") for i in range(n)]
        return out
