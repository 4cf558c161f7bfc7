/*
 * sw_oracle.c — CPU restatement of the parallel-genomeseq Smith-Waterman alignment path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under parallel-genomeseq_b200/ may include, link, load or
 * call this file; it exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg can check the CUDA path.  It is a plain, full-matrix, scalar restatement written from the
 * behaviour of the reference (file:line citations are relative to /root/reference/) and is
 * pinned in tests/test_oracle.py against
 *   (a) the reference's own known-answer tests (test/test_localaligner.cpp:25-26,54-58,31-42;
 *       test/test_skewedmatrix.cpp:17-23,58-65), and
 *   (b) golden vectors dumped from the reference itself, compiled here as oracle/_ref
 *       (tests/golden/make_golden.py).
 *
 * Conventions: x = first constructor argument (read / DB protein, length m),
 * y = second (reference / query, length n).  H is (m+1) x (n+1), row i <-> x[i-1],
 * column j <-> y[j-1], H[0][*] = H[*][0] = 0.  H is stored row-major: H[i*(n+1)+j].
 *
 * Two arithmetic modes, both exact in int32:
 *   SAT_U8 — Similarity_Matrix_Skewed semantics (uint8 saturating, byte-equality scoring);
 *   EXACT  — Similarity_Matrix semantics (FP32 holding exact integers, tabulated callback).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SWO_MODE_SAT_U8 0
#define SWO_MODE_EXACT 1

typedef int32_t cell_t;

/* _saturate(), src/aligner/similaritymatrix.cpp:376-384: clamp to [0,255], truncate toward zero. */
int swo_saturate(float a) {
  if (a < 0) return 0;
  if (a > 255) return 255;
  return (int)(uint8_t)a;
}

/*
 * SAT_U8 fill.  Similarity_Matrix_Skewed::iterate, similaritymatrix.cpp:386-561 with
 * dp_func(__m256i...) :75-81.  Parameters: M = sat(fn('A','A')), X = sat(-fn('A','T')),
 * G = sat(gap) (:389-392).  Per interior cell (byte equality, :415-417):
 *     diag = eq ? adds_epu8(NW, M) : subs_epu8(NW, X)      -> min(255, NW+M)  or  max(0, NW-X)
 *     H    = max(diag, subs_epu8(W, G), subs_epu8(N, G))
 * The three phases only change the traversal order, not the recurrence (verified
 * against the compiled reference on non-square shapes — square shapes hit a reference defect,
 * SURVEY F9, and are excluded from parity).
 */
void swo_fill_sat_u8(const uint8_t* x, int64_t m, const uint8_t* y, int64_t n, int M, int X, int G, cell_t* H) {
  const int64_t ld = n + 1;
  for (int64_t j = 0; j <= n; ++j) H[j] = 0;
  for (int64_t i = 1; i <= m; ++i) {
    H[i * ld] = 0;
    for (int64_t j = 1; j <= n; ++j) {
      int nw = H[(i - 1) * ld + (j - 1)], w = H[i * ld + (j - 1)], no = H[(i - 1) * ld + j];
      int diag;
      if (x[i - 1] == y[j - 1]) { diag = nw + M; if (diag > 255) diag = 255; }
      else { diag = nw - X; if (diag < 0) diag = 0; }
      int a = w - G; if (a < 0) a = 0;
      int b = no - G; if (b < 0) b = 0;
      int h = diag; if (a > h) h = a; if (b > h) h = b;
      H[i * ld + j] = h;
    }
  }
}

/*
 * EXACT fill.  Similarity_Matrix::iterate serial branch, similaritymatrix.cpp:99-116,247-257 with
 * the scalar dp_func :49-54: H = max(NW + fn(x[i-1], y[j-1]), W - g, N - g, 0).  The callback is
 * tabulated: table[a*256 + b] = fn(a, b) for byte values a (from x) and b (from y) — argument
 * order (x char, y char) as at :252-254.  Integer-valued scores/gap make FP32 exact (< 2^24).
 */
void swo_fill_exact(const uint8_t* x, int64_t m, const uint8_t* y, int64_t n, const int32_t* table, int gap, cell_t* H) {
  const int64_t ld = n + 1;
  for (int64_t j = 0; j <= n; ++j) H[j] = 0;
  for (int64_t i = 1; i <= m; ++i) {
    H[i * ld] = 0;
    const int32_t* trow = table + (int64_t)x[i - 1] * 256;
    for (int64_t j = 1; j <= n; ++j) {
      int h = H[(i - 1) * ld + (j - 1)] + trow[y[j - 1]];
      int a = H[i * ld + (j - 1)] - gap; if (a > h) h = a;
      int b = H[(i - 1) * ld + j] - gap; if (b > h) h = b;
      if (h < 0) h = 0;
      H[i * ld + j] = h;
    }
  }
}

/*
 * Skewed raw-index map.  _trueindex2rawindex, similaritymatrix.cpp:353-364, with the constructor's
 * role swap (:274-289): len_x = n+1 (reference / y), len_y = m+1 (read / x), nrows = min,
 * ncols = max; true index (ti, tj) = (column j of H, row i of H).
 */
void swo_true2raw(int64_t ti, int64_t tj, int64_t m, int64_t n, int64_t* ri, int64_t* rj) {
  const int64_t len_x = n + 1, len_y = m + 1;
  const int64_t nrows = len_x < len_y ? len_x : len_y, ncols = len_x < len_y ? len_y : len_x;
  if (ti + tj < nrows - 1) { *ri = ti; *rj = ti + tj; }
  else if (ti + tj > ncols - 1) { *ri = ti - ncols + len_y; *rj = ti + tj - (ncols - 1) - 1; }
  else { *ri = (len_x <= len_y) ? ti : len_y - 1 - tj; *rj = ti + tj; }
}

/* _rawindex2trueindex, similaritymatrix.cpp:330-346 (same role swap). */
void swo_raw2true(int64_t ri, int64_t rj, int64_t m, int64_t n, int64_t* ti, int64_t* tj) {
  const int64_t len_x = n + 1, len_y = m + 1;
  const int64_t nrows = len_x < len_y ? len_x : len_y;
  if (rj < nrows - 1) {
    if (ri <= rj) { *ti = ri; *tj = rj - ri; }
    else { *ti = len_x - nrows + ri; *tj = len_y - ri + rj; }
  } else {
    if (len_x <= len_y) { *ti = ri; *tj = rj - ri; }
    else { *ti = rj - (nrows - 1) + ri; *tj = nrows - 1 - ri; }
  }
}

/*
 * Arg-max, SAT_U8 path.  Similarity_Matrix_Skewed::find_index_of_maximum, similaritymatrix.cpp:291-299:
 * Eigen 3.3.7 maxCoeff(&r,&c) is a scalar visitor over the RAW (skewed, wrapped) storage in
 * column-major order with a strict '>' (Eigen/src/Core/Visitor.h:44-56,173-185), so the winner is
 * the maximal cell with the smallest raw key (rj, ri).  All padding cells of the raw matrix are 0,
 * so when the maximum is 0 the visitor stays at raw (0,0) -> true (0,0).
 * Returns index_x (row i of H, into x), index_y (column j of H, into y) as the reference does (:298).
 */
void swo_argmax_skewed(const cell_t* H, int64_t m, int64_t n, int64_t* index_x, int64_t* index_y, int32_t* maxv) {
  const int64_t ld = n + 1;
  cell_t best = 0; int64_t bri = 0, brj = 0, bi = 0, bj = 0;
  for (int64_t i = 0; i <= m; ++i) for (int64_t j = 0; j <= n; ++j) {
    cell_t v = H[i * ld + j];
    if (v < best || v == 0) continue;
    int64_t ri, rj; swo_true2raw(j, i, m, n, &ri, &rj);
    if (v > best || rj < brj || (rj == brj && ri < bri)) { best = v; bri = ri; brj = rj; bi = i; bj = j; }
  }
  *index_x = bi; *index_y = bj; *maxv = best;
}

/*
 * Arg-max, EXACT path.  Similarity_Matrix::find_index_of_maximum, similaritymatrix.cpp:21-28: the
 * same visitor over the plain column-major (m+1) x (n+1) matrix => smallest column j, then
 * smallest row i, among the maximal cells.
 */
void swo_argmax_colmajor(const cell_t* H, int64_t m, int64_t n, int64_t* index_x, int64_t* index_y, int32_t* maxv) {
  const int64_t ld = n + 1;
  cell_t best = H[0]; int64_t bi = 0, bj = 0;
  for (int64_t j = 0; j <= n; ++j) for (int64_t i = 0; i <= m; ++i)
    if (H[i * ld + j] > best) { best = H[i * ld + j]; bi = i; bj = j; }
  *index_x = bi; *index_y = bj; *maxv = best;
}

/*
 * Greedy traceback.  SWAligner::traceback, src/aligner/smithwaterman.cpp:40-78.  Looks only at the
 * three neighbours' VALUES: stop if any is 0 (emit both characters, pos = index_y); else NW if
 * n1>=n2 && n1>=n3; else W if n2>=n1 && n2>=n3 ('-' in consensus_x); else N ('-' in consensus_y).
 * Strings are emitted end -> start (smithwaterman.h:50-53).  Returns the consensus length, or -1 if
 * cap is too small, or -2 for the reference's undefined all-zero case (index 0).
 */
int64_t swo_traceback(const cell_t* H, const uint8_t* x, int64_t m, const uint8_t* y, int64_t n,
                      int64_t ix, int64_t iy, char* cx, char* cy, int64_t cap, uint32_t* pos) {
  (void)m;
  const int64_t ld = n + 1;
  int64_t len = 0;
  if (ix <= 0 || iy <= 0) return -2;
  for (;;) {
    cell_t n1 = H[(ix - 1) * ld + (iy - 1)], n2 = H[ix * ld + (iy - 1)], n3 = H[(ix - 1) * ld + iy];
    if (len >= cap) return -1;
    if (n1 == 0 || n2 == 0 || n3 == 0) {
      cx[len] = (char)x[ix - 1]; cy[len] = (char)y[iy - 1]; ++len;
      *pos = (uint32_t)iy;
      return len;
    }
    if (n1 >= n2 && n1 >= n3) { cx[len] = (char)x[ix - 1]; cy[len] = (char)y[iy - 1]; --ix; --iy; }
    else if (n2 >= n1 && n2 >= n3) { cx[len] = '-'; cy[len] = (char)y[iy - 1]; --iy; }
    else { cx[len] = (char)x[ix - 1]; cy[len] = '-'; --ix; }
    ++len;
  }
}

/*
 * One full alignment = SWAligner<SMT>::calculateScore, smithwaterman.cpp:80-108:
 * iterate -> find_index_of_maximum -> traceback.  mode selects the SMT.
 *   SAT_U8: p0 = M, p1 = X, gap = G already saturated by the caller (swo_saturate);  table unused.
 *   EXACT : table[256*256] int32, gap integer.
 * Returns consensus length (>=1), -1 cap too small, -2 all-zero matrix (reference UB), -3 alloc.
 */
int64_t swo_align(int mode, const uint8_t* x, int64_t m, const uint8_t* y, int64_t n,
                  int p0, int p1, const int32_t* table, int gap,
                  int32_t* score, uint32_t* pos, int64_t* end_x, int64_t* end_y,
                  char* cx, char* cy, int64_t cap) {
  cell_t* H = (cell_t*)malloc((size_t)(m + 1) * (size_t)(n + 1) * sizeof(cell_t));
  if (!H) return -3;
  int64_t ix, iy; int32_t mx;
  if (mode == SWO_MODE_SAT_U8) { swo_fill_sat_u8(x, m, y, n, p0, p1, gap, H); swo_argmax_skewed(H, m, n, &ix, &iy, &mx); }
  else { swo_fill_exact(x, m, y, n, table, gap, H); swo_argmax_colmajor(H, m, n, &ix, &iy, &mx); }
  if (score) *score = mx;
  if (end_x) *end_x = ix;
  if (end_y) *end_y = iy;
  int64_t len = swo_traceback(H, x, m, y, n, ix, iy, cx, cy, cap, pos);
  free(H);
  return len;
}

/* Dense matrix for cell-by-cell tests (test/test_skewedmatrix.cpp:39-66 compares the two SMTs). */
int swo_matrix(int mode, const uint8_t* x, int64_t m, const uint8_t* y, int64_t n,
               int p0, int p1, const int32_t* table, int gap, int32_t* H) {
  if (mode == SWO_MODE_SAT_U8) swo_fill_sat_u8(x, m, y, n, p0, p1, gap, H);
  else swo_fill_exact(x, m, y, n, table, gap, H);
  return 0;
}

/*
 * _make_string_range, src/aligner/plocalaligner.cpp:44-67.  overlap = (Index)(shortlen * ratio) is an
 * Index->float conversion, a float multiply and a truncation; piece = (longlen + (npiece-1)*overlap)
 * / npiece (integer division).  The reference's three asserts are live in its Release build
 * (SURVEY §5); here they become return codes: -1 overlap > piece, -2 right >= longlen before the
 * last piece, -3 npiece < 1.  Returns the number of ranges written.
 */
int swo_make_string_range(int npiece, int64_t shortlen, int64_t longlen, float ratio, int64_t* left, int64_t* right) {
  if (npiece < 1) return -3;
  int64_t ov = (int64_t)((float)shortlen * ratio);
  if (npiece == 1) { left[0] = 0; right[0] = longlen; return 1; }
  int64_t piece = (longlen + (int64_t)(npiece - 1) * ov) / npiece;
  if (ov > piece) return -1;
  int64_t l = 0, r = piece; int k = 0;
  left[k] = l; right[k] = r; ++k;
  while (k < npiece - 1) {
    l = r - ov; if (l < 0) l = 0;
    r = l + piece; if (r > longlen) r = longlen;
    left[k] = l; right[k] = r; ++k;
  }
  if (!(r < longlen)) return -2;
  l = r - ov; if (l < 0) l = 0;
  left[k] = l; right[k] = longlen; ++k;
  return k;
}

/*
 * OMPParallelLocalAligner<SMT, SWAligner<SMT>>::calculateScore, plocalaligner.cpp:106-143, in its
 * deterministic serial semantic: iterate every piece with the CONSTRUCTOR's scoring
 * (:113-115); pick the lowest-index piece with the strictly greatest maximum (:122-129); re-run a
 * full SWAligner on that piece with the DEFAULT scoring (+3/-3, gap 2 — :135 drops the custom
 * callback); pos = la.getPos() + left (:137).
 *   piece scoring: (p0, p1, table, gap) as in swo_align; default scoring is fixed here.
 */
int64_t swo_align_chunked(int mode, const uint8_t* x, int64_t m, const uint8_t* y, int64_t n,
                          int npiece, float ratio, int p0, int p1, const int32_t* table, int gap,
                          int32_t* score, uint32_t* pos, int* best_piece, char* cx, char* cy, int64_t cap) {
  int64_t* left = (int64_t*)malloc(sizeof(int64_t) * (size_t)(npiece > 0 ? npiece : 1) * 2);
  if (!left) return -3;
  int64_t* right = left + (npiece > 0 ? npiece : 1);
  int k = swo_make_string_range(npiece, m, n, ratio, left, right);
  if (k < 0) { free(left); return -10 + k; }
  int32_t best = -1; int bp = 0;
  for (int p = 0; p < k; ++p) {
    int64_t pn = right[p] - left[p];
    cell_t* H = (cell_t*)malloc((size_t)(m + 1) * (size_t)(pn + 1) * sizeof(cell_t));
    if (!H) { free(left); return -3; }
    int64_t ix, iy; int32_t mx;
    if (mode == SWO_MODE_SAT_U8) { swo_fill_sat_u8(x, m, y + left[p], pn, p0, p1, gap, H); swo_argmax_skewed(H, m, pn, &ix, &iy, &mx); }
    else { swo_fill_exact(x, m, y + left[p], pn, table, gap, H); swo_argmax_colmajor(H, m, pn, &ix, &iy, &mx); }
    free(H);
    if (mx > best) { best = mx; bp = p; }
  }
  int64_t len;
  uint32_t lp = 0;
  if (mode == SWO_MODE_SAT_U8) {
    len = swo_align(mode, x, m, y + left[bp], right[bp] - left[bp], 3, 3, NULL, 2, score, &lp, NULL, NULL, cx, cy, cap);
  } else {
    int32_t* deft = (int32_t*)malloc(sizeof(int32_t) * 65536);
    if (!deft) { free(left); return -3; }
    for (int a = 0; a < 256; ++a) for (int b = 0; b < 256; ++b) deft[a * 256 + b] = (a == b) ? 3 : -3;
    len = swo_align(mode, x, m, y + left[bp], right[bp] - left[bp], 0, 0, deft, 2, score, &lp, NULL, NULL, cx, cy, cap);
    free(deft);
  }
  if (pos) *pos = lp + (uint32_t)left[bp];
  if (best_piece) *best_piece = bp;
  free(left);
  return len;
}
