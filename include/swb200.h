/*
 * swb200.h — C ABI of the B200-native Smith-Waterman engine (libswb200.so).
 *
 * This is the drop-in boundary for the alignment hot path of kosta777/parallel-genomeseq.  The
 * reference has no FFI layer; its seam is the C++ aligner interface, so every entry point below
 * names the reference interface it stands in for (paths relative to the reference repository):
 *
 *   swb_set_scoring / _match     the scoring callback + gap penalty taken by the constructors
 *                                src/aligner/smithwaterman.h:14-17, src/aligner/plocalaligner.h:9-12
 *   swb_set_reference            the second constructor argument (sequence_y: reference / query)
 *                                src/aligner/smithwaterman.h:14, src/sw_solve_small.cpp:84
 *   swb_align_batch              the per-read loop "construct aligner; calculateScore(); getPos();
 *                                getScore(); getConsensus_x/y(); getTimings()" of the drivers:
 *                                src/sw_solve_small.cpp:56-101, src/mpi_sw_solve_uniprot.cpp:95-138
 *                                (LocalAligner / ParallelLocalAligner, src/aligner/localaligner.h:7-28)
 *   swb_make_string_range        _make_string_range, src/aligner/plocalaligner.cpp:44-67
 *
 * Plain pointers and sizes only; the caller owns every buffer; no exceptions cross the boundary;
 * every function returns 0 on success or a negative SWB_ERR_* code (swb_last_error gives the text).
 * One context per host thread.  There is NO CPU fallback: without a CUDA device swb_create fails.
 */
#ifndef SWB200_H_
#define SWB200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct swb_ctx swb_ctx;

/* Arithmetic modes: */
#define SWB_MODE_SAT_U8 0 /* Similarity_Matrix_Skewed: uint8 saturating, byte-equality scoring, skewed arg-max order */
#define SWB_MODE_EXACT 1  /* Similarity_Matrix: exact integers, tabulated callback, column-major arg-max order     */

#define SWB_OK 0
#define SWB_ERR_CUDA -1      /* CUDA runtime error (text in swb_last_error) */
#define SWB_ERR_ARG -2       /* bad argument */
#define SWB_ERR_RANGE -3     /* _make_string_range precondition (the reference asserts, plocalaligner.cpp:52,63,65) */
#define SWB_ERR_SCORING -4   /* scoring not integer-valued / out of the 16-bit lane range */
#define SWB_ERR_UNSUPPORTED -5 /* shape not supported by this build (see text) */
#define SWB_ERR_STATE -6     /* call order (no reference / no staged batch) */

#define SWB_FLAG_CONSENSUS 1u /* produce consensus_x / consensus_y */

/* Per-alignment result flags (out_flags) */
#define SWB_RES_CONS_TRUNCATED 1u /* consensus longer than cons_stride: strings truncated, score/end still exact */

int swb_device_count(void); /* CUDA devices visible to this process (0 when there is none) */
int swb_create(int device, swb_ctx** out);
void swb_destroy(swb_ctx* ctx);
const char* swb_last_error(const swb_ctx* ctx);
const char* swb_version(void);

/*
 * Scoring.  `table` is the callback tabulated by the host: table[a * 256 + b] = fn((char)a, (char)b),
 * a from sequence_x, b from sequence_y (argument order of similaritymatrix.cpp:252-254).
 *   SWB_MODE_SAT_U8: only fn('A','A'), -fn('A','T') and gap are used, each clamped by the reference's
 *                    _saturate() (similaritymatrix.cpp:376-392); cells are scored by byte equality.
 *   SWB_MODE_EXACT : every entry is used; entries and gap must be integer-valued.
 */
int swb_set_scoring(swb_ctx* ctx, int mode, const float* table, float gap);
/* Convenience for the default-shaped callback a == b ? match : mismatch (smithwaterman.cpp:8). */
int swb_set_scoring_match(swb_ctx* ctx, int mode, float match, float mismatch, float gap);

/* sequence_y for all following batches (copied to the device). */
int swb_set_reference(swb_ctx* ctx, const char* y, size_t n);

/*
 * Align n_seqs sequences x_r = seqs[offsets[r] .. offsets[r+1]) against the reference.
 *   npiece <= 0 : SWAligner<SMT>(x_r, y, fn, gap).calculateScore()
 *   npiece >= 1 : OMPParallelLocalAligner<SMT, SWAligner<SMT>>(x_r, y, npiece, ratio, fn, gap).calculateScore()
 *                 in its deterministic serial semantic, including the re-alignment of the
 *                 winning piece with the DEFAULT scoring (plocalaligner.cpp:135).
 * Outputs (any may be NULL): score[r] = getScore() (exact integer), pos[r] = getPos() (1-based),
 * cons_x/cons_y + r*cons_stride = getConsensus_x/y() (end -> start, not NUL-terminated), cons_len[r],
 * end_xy[2r..2r+1] = (index_x, index_y) of find_index_of_maximum (of the winning piece when chunked),
 * *device_us = CUDA-event time of all kernels of the batch (the getTimings()[0] analogue).
 */
int swb_align_batch(swb_ctx* ctx, const char* seqs, const uint64_t* offsets, size_t n_seqs,
                    int npiece, float ratio, unsigned flags,
                    int32_t* score, uint32_t* pos, uint32_t* end_xy,
                    char* cons_x, char* cons_y, uint32_t* cons_len, size_t cons_stride,
                    uint32_t* out_flags, float* device_us);

/*
 * The same call split in three so that a driver (or bench.py) can keep inputs resident in HBM:
 * stage = host preparation + H2D; run = kernels only (may be repeated); fetch = D2H of the results.
 */
int swb_batch_stage(swb_ctx* ctx, const char* seqs, const uint64_t* offsets, size_t n_seqs, int npiece, float ratio, unsigned flags, size_t cons_stride);
int swb_batch_run(swb_ctx* ctx, float* device_us);
int swb_batch_fetch(swb_ctx* ctx, int32_t* score, uint32_t* pos, uint32_t* end_xy,
                    char* cons_x, char* cons_y, uint32_t* cons_len, uint32_t* out_flags);
/*
 * Database search (mpi_sw_solve_uniprot.cpp:95-138 aligns every database protein against the query; a search over many
 * queries repeats that loop): swap sequence_y while the staged batch stays resident in HBM.  Only valid when the batch
 * was staged in query-stationary mode and the new reference fits the staged lane geometry; otherwise SWB_ERR_STATE
 * (call swb_set_reference + swb_batch_stage instead).
 */
int swb_batch_rebind_reference(swb_ctx* ctx, const char* y, size_t n);
/* Device pointers of the last run's per-sequence (score, pos) arrays, n_seqs entries each (for a NCCL gather). */
int swb_batch_device_results(swb_ctx* ctx, const int32_t** d_score, const uint32_t** d_pos);

/* Work done by the last swb_batch_run. */
typedef struct swb_stats {
  uint64_t cells_reference;  /* sum len(x) * len(y): the reference drivers' GCUPS numerator (sw_solve_small.cpp:89) */
  uint64_t cells_executed;   /* DP cells actually computed in pass 1 (padding, chunk overlap, re-alignment included) */
  uint64_t cells_pass2;      /* DP cells recomputed by locate + traceback (upper bound) */
  uint32_t kernel_launches;  /* kernels launched by the last run */
  uint32_t lanes_per_pair;   /* L of the (last) launch class */
  uint32_t rows_per_lane;    /* R */
  uint32_t block_steps;      /* B */
  float pass1_us, pass2_us;  /* CUDA-event split of device_us */
  uint32_t cols_per_step;    /* C: columns a lane advances per wavefront step */
  uint32_t kernel_kind;      /* pass-1 kernel of the (last) launch class: 0 batched, 1 pipelined strips, 2 query-stationary */
} swb_stats;
int swb_last_stats(const swb_ctx* ctx, swb_stats* out);

/* _make_string_range (plocalaligner.cpp:44-67).  Returns the number of ranges or SWB_ERR_RANGE. */
int swb_make_string_range(int npiece, int64_t shortlen, int64_t longlen, float ratio, int64_t* left, int64_t* right);

/* Dense H (row-major (m+1) x (n+1) int32) through the device path — the operator()(row, col) surface of
 * Abstract_Similarity_Matrix (similaritymatrix.h:13-24) for tests and small inputs only. */
int swb_matrix(swb_ctx* ctx, const char* x, size_t m, int32_t* out);

#ifdef __cplusplus
}
#endif
#endif /* SWB200_H_ */
