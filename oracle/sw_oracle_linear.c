/*
 * sw_oracle_linear.c — LINEAR-MEMORY CPU restatement of the parallel-genomeseq alignment path.
 *
 * TEST INFRASTRUCTURE ONLY (same rule as sw_oracle.c): only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it; the product never does.
 *
 * Why it exists: the reference materialises the whole matrix
 * (similaritymatrix.cpp:17 MatrixXf (m+1)x(n+1); :287 u8 (nrows+32) x ncols), so neither it nor the
 * full-matrix restatement in sw_oracle.c can check BASELINE config 3 on many reads (183 MB per read) or
 * config 5 at its stated size (10 kbp x 51 Mbp = 2 TB).  This file computes the SAME function
 *     SWAligner<SMT>::calculateScore()  =  iterate -> find_index_of_maximum -> traceback
 *     (smithwaterman.cpp:80-108)
 * in O(m + window) memory.  It is a second, independently arranged restatement (anti-diagonal sweep over three
 * rotating diagonals; sw_oracle.c fills a full row-major matrix) and is pinned in tests/test_oracle.py by
 * differential tests against sw_oracle.c, against the compiled reference (oracle/_ref) and against every golden
 * vector of tests/golden/.
 *
 *   pass 1  sweep the anti-diagonals d = i + j = 2 .. m + n keeping three diagonals of m+1 cells (every cell of a
 *           diagonal depends only on the two diagonals before it, which is also how the reference's skewed matrix
 *           walks, similaritymatrix.cpp:403-557); track the maximum and the cell the reference's arg-max would
 *           return (first maximum in the raw storage order: keys below); keep diagonals d-1 and d for every K-th d.
 *   pass 2  recompute from the checkpoint before the window [je - W, je] x [0, ie] and keep that window densely,
 *           then walk back literally as smithwaterman.cpp:40-78 does; if the walk needs a column left of the
 *           window, double W and repeat (exact for any path length).
 *
 * Recurrences (same citations as sw_oracle.c):
 *   SAT_U8  similaritymatrix.cpp:75-81,415-417:  diag = eq ? min(255, NW + M) : max(0, NW - X);
 *           H = max(diag, max(0, W - G), max(0, N - G))
 *   EXACT   similaritymatrix.cpp:49-54,252-254:  H = max(NW + fn(x[i-1], y[j-1]), W - g, N - g, 0)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SWO_MODE_SAT_U8 0
#define SWO_MODE_EXACT 1

/* _trueindex2rawindex with the constructor's role swap (similaritymatrix.cpp:274-289,353-364): Eigen's column-major
 * visitor with strict '>' returns the maximal cell with the smallest (rj, ri) (Visitor.h:44-56,173-185). */
static inline void skew_key(int64_t i, int64_t j, int64_t m, int64_t n, int64_t* rj, int64_t* ri) {
  const int64_t len_x = n + 1, len_y = m + 1;
  const int64_t nrows = len_x < len_y ? len_x : len_y, ncols = len_x < len_y ? len_y : len_x;
  const int64_t ti = j, tj = i;
  if (ti + tj < nrows - 1) { *ri = ti; *rj = ti + tj; }
  else if (ti + tj > ncols - 1) { *ri = ti - ncols + len_y; *rj = ti + tj - (ncols - 1) - 1; }
  else { *ri = (len_x <= len_y) ? ti : len_y - 1 - tj; *rj = ti + tj; }
}

typedef struct {
  int mode;
  const uint8_t* x; int64_t m;
  const uint8_t* yr; int64_t n;     /* yr = y reversed: y[j-1] = yr[n-j], so a diagonal reads it with unit stride */
  int M, X, G;                      /* SAT_U8; EXACT match-shaped tables use M / X as well */
  const int32_t* table;             /* EXACT, general callback */
  int shaped;                       /* EXACT: table is a == b ? M : -X */
} lin_ctx;

/*
 * Diagonal d from the two before it.  Arrays are indexed by the row i; cell (i, j = d - i).
 *   A = diagonal d-2 (NW = A[i-1]), B = diagonal d-1 (W = B[i], N = B[i-1]).
 * Interior rows: i in [max(1, d-n), min(m, d-1)].  Row 0 (index 0) and the column-0 cell of each diagonal (index
 * ihi+1) are never written and stay 0 from calloc (the valid ranges only grow at that end), indices below ilo are
 * never read (see the range argument in the file header of tests/test_oracle.py::test_linear_oracle).
 * Returns the diagonal's maximum.
 */
static int32_t next_diagonal(const lin_ctx* c, int64_t d, const int32_t* A, const int32_t* B, int32_t* C) {
  const int64_t m = c->m, n = c->n;
  const int64_t ilo = d - n > 1 ? d - n : 1, ihi = d - 1 < m ? d - 1 : m;
  const uint8_t* x = c->x - 1;                 /* x[i] = character of row i */
  const uint8_t* yr = c->yr;
  const int64_t off = n - d;                   /* yr[off + i] = y[d - i - 1] = character of column j = d - i */
  const int32_t G = c->G;
  int32_t mx = 0;
  if (c->mode == SWO_MODE_SAT_U8) {
    const int32_t M = c->M, X = c->X;
    for (int64_t i = ilo; i <= ihi; ++i) {
      int32_t dg = A[i - 1] + (x[i] == yr[off + i] ? M : -X);            /* byte equality, :415-417 */
      dg = dg < 0 ? 0 : dg; dg = dg > 255 ? 255 : dg;               /* adds_epu8 / subs_epu8 */
      int32_t a = B[i] - G; a = a < 0 ? 0 : a;
      int32_t b = B[i - 1] - G; b = b < 0 ? 0 : b;
      int32_t h = dg > a ? dg : a; h = h > b ? h : b;
      C[i] = h;
      mx = h > mx ? h : mx;
    }
  } else if (c->shaped) {
    const int32_t M = c->M, X = c->X;
    for (int64_t i = ilo; i <= ihi; ++i) {
      int32_t h = A[i - 1] + (x[i] == yr[off + i] ? M : -X);
      int32_t a = B[i] - G, b = B[i - 1] - G;
      h = h > a ? h : a; h = h > b ? h : b; h = h < 0 ? 0 : h;
      C[i] = h;
      mx = h > mx ? h : mx;
    }
  } else {
    const int32_t* T = c->table;
    for (int64_t i = ilo; i <= ihi; ++i) {
      int32_t h = A[i - 1] + T[(int32_t)x[i] * 256 + yr[off + i]];       /* fn(x char, y char), :252-254 */
      int32_t a = B[i] - G, b = B[i - 1] - G;
      h = h > a ? h : a; h = h > b ? h : b; h = h < 0 ? 0 : h;
      C[i] = h;
      mx = h > mx ? h : mx;
    }
  }
  return mx;
}

/*
 * Same contract as swo_align (sw_oracle.c): returns the consensus length (>= 1), -1 cap too small,
 * -2 all-zero matrix (reference UB), -3 allocation failure.
 * SAT_U8: p0 = M, p1 = X, gap = G (already saturated); EXACT: table[256*256], integer gap.
 */
int64_t swo_align_linear(int mode, const uint8_t* x, int64_t m, const uint8_t* y, int64_t n,
                         int p0, int p1, const int32_t* table, int gap,
                         int32_t* score, uint32_t* pos, int64_t* end_x, int64_t* end_y,
                         char* cx, char* cy, int64_t cap) {
  lin_ctx c; memset(&c, 0, sizeof c);
  c.mode = mode; c.x = x; c.m = m; c.n = n; c.M = p0; c.X = p1; c.G = gap; c.table = table;
  if (mode == SWO_MODE_EXACT) {
    const int32_t dv = table[0], ov = table[1];
    c.shaped = 1;
    for (int a = 0; a < 256 && c.shaped; ++a) for (int b = 0; b < 256; ++b) if (table[a * 256 + b] != (a == b ? dv : ov)) { c.shaped = 0; break; }
    if (c.shaped) { c.M = dv; c.X = -ov; }
  }
  const int64_t D = m + n;                             /* last diagonal */
  /* checkpoint period: a power of two >= 256 keeping all checkpoints (two diagonals each) within 512 MB */
  int64_t K = 256;
  while ((D / K + 1) * 2 * (m + 1) * 4 > ((int64_t)512 << 20)) K <<= 1;
  const int64_t nck = D / K + 1;                       /* checkpoint q holds diagonals q*K - 1 and q*K (q = 0: zeros) */
  const size_t W1 = (size_t)(m + 2);
  uint8_t* yr = (uint8_t*)malloc((size_t)n + (size_t)m + 64);
  int32_t* ck = (int32_t*)calloc((size_t)nck * 2 * W1, sizeof(int32_t));
  int32_t* buf = (int32_t*)calloc(3 * W1, sizeof(int32_t));
  if (!yr || !ck || !buf) { free(yr); free(ck); free(buf); return -3; }
  for (int64_t j = 1; j <= n; ++j) yr[n - j] = y[j - 1];
  c.yr = yr;

  /* ---- pass 1: maximum and the reference's arg-max cell -------------------------------------------------- */
  int32_t best = 0; int64_t bi = 0, bj = 0, k1 = 0, k2 = 0;
  int32_t *A = buf, *B = buf + W1, *Cd = buf + 2 * W1;
  for (int64_t d = 2; d <= D; ++d) {
    const int32_t mx = next_diagonal(&c, d, A, B, Cd);
    if (mx >= best && mx > 0) {
      const int64_t ilo = d - n > 1 ? d - n : 1, ihi = d - 1 < m ? d - 1 : m;
      /* ties (saturated plateaus are full of them): skip the diagonal when none of its keys can beat the winner's.
       * SAT_U8: rj depends on d alone (every branch of the index map is a function of ti + tj); EXACT: j >= d - ihi. */
      int64_t kmin;
      if (mode == SWO_MODE_SAT_U8) { int64_t r0; skew_key(ilo, d - ilo, m, n, &kmin, &r0); }
      else kmin = d - ihi;
      if (mx == best && kmin > k1) goto next_d;
      for (int64_t i = ilo; i <= ihi; ++i) {
        if (Cd[i] != mx) continue;
        const int64_t j = d - i;
        int64_t a, b;
        if (mode == SWO_MODE_SAT_U8) skew_key(i, j, m, n, &a, &b);     /* skewed raw storage: (rj, ri) */
        else { a = j; b = i; }                                         /* plain column-major storage, similaritymatrix.cpp:21-28 */
        if (mx > best || a < k1 || (a == k1 && b < k2)) { best = mx; bi = i; bj = j; k1 = a; k2 = b; }
      }
    }
  next_d:
    if ((d % K) == 0) {
      memcpy(ck + (size_t)(d / K) * 2 * W1, B, sizeof(int32_t) * W1);
      memcpy(ck + (size_t)(d / K) * 2 * W1 + W1, Cd, sizeof(int32_t) * W1);
    }
    int32_t* t = A; A = B; B = Cd; Cd = t;
  }
  if (score) *score = best;
  if (end_x) *end_x = bi;
  if (end_y) *end_y = bj;
  if (bi <= 0 || bj <= 0) { free(yr); free(ck); free(buf); return -2; }

  /* ---- pass 2: dense window [c_lo, bj] x [0, bi], literal traceback (smithwaterman.cpp:40-78) ------------ */
  /* The window is kept diagonal-major (one contiguous row of `rows` cells per anti-diagonal c_lo .. bi + bj), so
   * that storing a diagonal is one memcpy; cells that are never written (row 0, column 0, outside the matrix)
   * read as 0 from calloc. */
  int64_t len = -3;
  for (int64_t W = 256;; W *= 2) {
    const int64_t c_lo = bj - W > 0 ? bj - W : 0;
    const int64_t q = c_lo / K;                        /* every window cell with i >= 1 lies on a diagonal > q*K */
    const int64_t rows = bi + 1, ndiag = bi + bj - c_lo + 1;
    int32_t* win = (int32_t*)calloc((size_t)rows * (size_t)ndiag, sizeof(int32_t));
    if (!win) { len = -3; break; }
    /* fresh zeroed buffers: the stale-index argument of next_diagonal needs the never-written cells to be 0 */
    memset(buf, 0, sizeof(int32_t) * 3 * W1);
    A = buf; B = buf + W1; Cd = buf + 2 * W1;
    memcpy(A, ck + (size_t)q * 2 * W1, sizeof(int32_t) * W1);
    memcpy(B, ck + (size_t)q * 2 * W1 + W1, sizeof(int32_t) * W1);
    for (int64_t d = q * K + 1; d <= bi + bj; ++d) {
      if (d >= 2) {
        next_diagonal(&c, d, A, B, Cd);
        const int64_t ilo = d - n > 1 ? d - n : 1, ihi = d - 1 < m ? d - 1 : m;
        int64_t lo = d - bj > ilo ? d - bj : ilo;      /* j <= bj */
        int64_t hi = d - c_lo < ihi ? d - c_lo : ihi;  /* j >= c_lo */
        if (hi > bi) hi = bi;
        if (d >= c_lo && hi >= lo) memcpy(win + (size_t)(d - c_lo) * (size_t)rows + (size_t)lo, Cd + lo, sizeof(int32_t) * (size_t)(hi - lo + 1));
      }
      int32_t* t = A; A = B; B = Cd; Cd = t;
    }
#define HW(i, j) win[(size_t)((i) + (j) - c_lo) * (size_t)rows + (size_t)(i)]
    int64_t ix = bi, iy = bj;
    int retry = 0;
    len = 0;
    for (;;) {
      if (iy - 1 < c_lo) { retry = 1; break; }         /* only with c_lo > 0: column 0 is inside any window that starts at 0 */
      const int32_t n1 = HW(ix - 1, iy - 1), n2 = HW(ix, iy - 1), n3 = HW(ix - 1, iy);
      if (len >= cap) { len = -1; break; }
      if (n1 == 0 || n2 == 0 || n3 == 0) {
        cx[len] = (char)x[ix - 1]; cy[len] = (char)y[iy - 1]; ++len;
        *pos = (uint32_t)iy;
        break;
      }
      if (n1 >= n2 && n1 >= n3) { cx[len] = (char)x[ix - 1]; cy[len] = (char)y[iy - 1]; --ix; --iy; }
      else if (n2 >= n1 && n2 >= n3) { cx[len] = '-'; cy[len] = (char)y[iy - 1]; --iy; }
      else { cx[len] = (char)x[ix - 1]; cy[len] = '-'; --ix; }
      ++len;
    }
#undef HW
    free(win);
    if (!retry) break;
  }
  free(yr); free(ck); free(buf);
  return len;
}
