/*
 * ozimmu_b200.h -- C-ABI of the B200-native Ozaki-scheme DGEMM (drop-in for enp1s0/ozIMMU).
 *
 * Plain C: pointers, sizes and ints only.  `stream` arguments are a cudaStream_t passed as
 * void*.  Every function returns 0 on success; kernel launchers return the cudaError_t
 * value otherwise, host-API functions return 1 for an invalid argument (as the reference
 * does) or a negative number for a CUDA / internal failure.  Nothing here throws, nothing
 * synchronises the device unless stated.
 *
 * Two layers live in the one shared library (ozimmu_b200/lib/libozimmu.so):
 *
 *  1. ozk_*   -- the kernel ABI: thin launchers over the hand-written sm_100a kernels
 *                (ozimmu_b200/csrc/split.cu, gemm_fused.cu).  Raw device pointers.
 *  2. ozimmu_* -- the host API: a C spelling of the reference's public C++ interface
 *                (reference include/ozimmu/ozimmu.hpp:47-100).  The same library also
 *                exports that C++ interface itself (namespace mtk::ozimmu, declared in
 *                include/ozimmu/ozimmu.hpp of this repo) and the cuBLAS interposers
 *                (cublasDgemm_v2, cublasGemmEx, ... -- reference src/cublas.cu:103-513),
 *                so LD_PRELOAD=libozimmu.so behaves like the reference's libozimmu.so.
 *
 * Each declaration cites the reference interface it replaces (paths relative to the
 * reference tree).
 */
#ifndef OZIMMU_B200_H
#define OZIMMU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------
 * Enumerations (values identical to reference include/ozimmu/ozimmu.hpp:12-45)
 * ------------------------------------------------------------------------------------- */
enum { OZIMMU_OP_N = 0, OZIMMU_OP_T = 1 };                      /* operation_t   :12 */
enum {                                                          /* compute_mode_t :14-37 */
  OZIMMU_SGEMM = 0,
  OZIMMU_DGEMM = 1,
  OZIMMU_FP64_INT8_3 = 2, /* fp64_int8_S == S - 1 for S = 3..18 */
  OZIMMU_FP64_INT8_18 = 17,
  OZIMMU_FP64_INT8_AUTO = 18
};
enum { OZIMMU_MALLOC_SYNC = 0, OZIMMU_MALLOC_ASYNC = 1 };       /* malloc_mode_t :41 */
enum { OZIMMU_REAL = 0, OZIMMU_COMPLX = 1 };                    /* element_kind_t :43-46 */

typedef struct ozimmu_handle_s *ozimmu_handle_t;                /* mtk::ozimmu::handle_t :9-11 */

/* ---------------------------------------------------------------------------------------
 * 1. Kernel ABI (device pointers; asynchronous on `stream`)
 * ------------------------------------------------------------------------------------- */

/* reference src/split.cu:520-536 get_bits_per_int8 */
uint32_t ozk_bits_per_int8(uint32_t k);

/* int8 slice layout shared by the split and GEMM kernels ("blocked SW128"): one operand is
 * [num_split][row tile][k block][128 rows][128 bytes] -- every 128-row x 128-byte tile is 16 KB contiguous
 * and pre-swizzled for shared memory (16-byte chunk c of row r sits at chunk c ^ (r & 7)), rows are padded
 * to a multiple of 256 and k to a multiple of 128 with zeros.  (The reference keeps [slice][row][k4] and pads
 * k to 4, src/utils.hpp:30-39; the padding is zero in both, so products are identical.)
 * ozk_slice_pitch(k): k rounded up to 128.  ozk_slices_bytes: bytes of one operand. */
size_t ozk_slice_pitch(size_t k);
size_t ozk_slices_bytes(size_t rows, size_t k, unsigned num_split);

/* reference src/split.cu:193-283 (split_int8_kernel + split_int8_A/split_int8):
 * per-"row" max exponent scan and FP64 -> num_split x int8 mantissa split.
 *   in        : rows x len view of op(X).  col_major != 0: element (r,c) = in[c*ld + r]
 *               (op_n A / op_t B), else in[r*ld + c] (op_t A / op_n B).
 *   out       : ozk_slices_bytes(rows, len, num_split) bytes in the blocked layout above; pitch =
 *               ozk_slice_pitch(len); all padding (rows up to a multiple of 256, k up to pitch) is zeroed.
 *   max_exp   : [rows] doubles, 2 * 2^(emax-1023) (reference :191,202-204,234-241).
 *   scratch   : [rows] uint32 device scratch (only used when col_major != 0).
 */
int ozk_split_int8(int8_t *out, size_t pitch, double *max_exp, uint32_t *scratch, size_t rows,
                   size_t len, const double *in, size_t ld, int col_major, unsigned num_split,
                   unsigned bits_per_int8, void *stream);

/* ozk_split_int8 on one plane of an interleaved complex matrix (reference src/split.cu:69-152,
 * 211-216: real and imaginary parts are scaled and cut independently): `in` points at the plane's
 * first double, elem_stride = 2, ld counts complex elements.  elem_stride = 1 is ozk_split_int8. */
int ozk_split_int8_strided(int8_t *out, size_t pitch, double *max_exp, uint32_t *scratch, size_t rows,
                           size_t len, const double *in, size_t ld, int col_major, unsigned num_split,
                           unsigned bits_per_int8, unsigned elem_stride, void *stream);

/* Row block [row0, row0 + rows) of an operand whose slice planes hold plane_rows rows (the host-operand
 * pipeline splits blocks of A / B as they arrive over PCIe; the split scales every row on its own, reference
 * src/split.cu:193-242, so a block-wise split is bit-identical to a whole-matrix one).  out = base of the
 * operand's slices; max_exp / scratch / in point at row row0's entries.  row0 must be a multiple of 256 and
 * the block must end on a multiple of 256 or at plane_rows; the block that ends the plane zeroes the padding. */
int ozk_split_int8_block(int8_t *out, size_t pitch, size_t plane_rows, size_t row0, double *max_exp,
                         uint32_t *scratch, size_t rows, size_t len, const double *in, size_t ld,
                         int col_major, unsigned num_split, unsigned bits_per_int8, unsigned elem_stride,
                         void *stream);

/* ozk_split_int8 for `batch` operands of the same shape in ONE launch (strided-batched GEMMs): entry e reads
 * in + e*in_stride (doubles) and writes its slices at out + e*out_stride (bytes), its row scales at
 * max_exp + e*max_stride (doubles) and uses scratch + e*scr_stride (uint32).  batch <= 65535. */
int ozk_split_int8_batched(int8_t *out, size_t out_stride, size_t pitch, double *max_exp, size_t max_stride,
                           uint32_t *scratch, size_t scr_stride, size_t rows, size_t len, const double *in,
                           size_t ld, size_t in_stride, int col_major, unsigned num_split,
                           unsigned bits_per_int8, size_t batch, void *stream);
/* the same on one plane of interleaved complex matrices (elem_stride = 2; in_stride still counts doubles) */
int ozk_split_int8_batched_strided(int8_t *out, size_t out_stride, size_t pitch, double *max_exp, size_t max_stride,
                                   uint32_t *scratch, size_t scr_stride, size_t rows, size_t len, const double *in,
                                   size_t ld, size_t in_stride, int col_major, unsigned num_split,
                                   unsigned bits_per_int8, unsigned elem_stride, size_t batch, void *stream);

/* reference src/gemm.cu:266-334 (matmul_core -> cublasGemmEx int8) + :77-102
 * (accumulate_in_f64) + :104-122 (init_accumulator_buffer) + :124-158 (axby), fused:
 * for every (i,j) of the reference pair order (src/config.cu:85-92) an exact
 * int8 x int8 -> int32 product on tcgen05 tensor cores, accumulated per element in FP64 in
 * the reference's op order, then scaled by 2^-44 * amax[r] * bmax[c] and written as
 * C = alpha*x (+ beta*C).  C is column-major with leading dimension ldc; alpha/beta by value.
 */
int ozk_gemm_i8_fused(size_t m, size_t n, size_t k, const int8_t *a_slices, const int8_t *b_slices,
                      size_t pitch, const double *amax, const double *bmax, unsigned num_split,
                      unsigned bits_per_int8, double alpha, double beta, double *c, size_t ldc,
                      void *stream);

/* ozk_gemm_i8_fused on the m x n block of C whose first element is (row0, col0) (both multiples of 256):
 * a_slices / b_slices are the bases of operands split for a_plane_rows / b_plane_rows rows, amax / bmax / c
 * point at the block's first entries.  flags: OZK_FUSED_NO_LOCKSTEP = do not pace the CTA pairs against
 * each other (for launches that share the GPU with another launch). */
#define OZK_FUSED_NO_LOCKSTEP 1u
/* one CTA pair per tile (non-persistent launch): concurrent launches fill the SMs tile by tile through the hardware
 * CTA scheduler, and a higher-priority kernel gets SMs whenever a tile ends (the block pipelines use this) */
#define OZK_FUSED_ONE_TILE_PER_PAIR 2u
int ozk_gemm_i8_fused_block(size_t m, size_t n, size_t k, const int8_t *a_slices, size_t a_plane_rows,
                            size_t row0, const int8_t *b_slices, size_t b_plane_rows, size_t col0,
                            size_t pitch, const double *amax, const double *bmax, unsigned num_split,
                            unsigned bits_per_int8, double alpha, double beta, double *c, size_t ldc,
                            unsigned flags, void *stream);

/* Grouped launch for a strided batch (reference src/cublas.cu:315-472 loops one GEMM per entry, :380-406):
 * `batch` independent m x n x k products in ONE persistent launch, tiles of all entries in one queue so that
 * small entries fill the GPU together.  Entry e uses a_slices + e*a_batch_bytes, b_slices + e*b_batch_bytes,
 * amax + e*amax_batch, bmax + e*bmax_batch and c + e*c_batch (the last three counted in doubles). */
int ozk_gemm_i8_fused_batched(size_t m, size_t n, size_t k, size_t batch, const int8_t *a_slices,
                              size_t a_batch_bytes, const int8_t *b_slices, size_t b_batch_bytes, size_t pitch,
                              const double *amax, size_t amax_batch, const double *bmax, size_t bmax_batch,
                              unsigned num_split, unsigned bits_per_int8, double alpha, double beta, double *c,
                              size_t ldc, size_t c_batch, void *stream);

/* The general form of the fused launch -- real or complex C, whole matrix or block, single or strided batch,
 * scalars by value or in device memory; the ozk_gemm_i8_fused* functions above fill this struct.  Zero-initialise
 * it and set what applies.
 *   complex_c != 0: a complex GEMM (reference src/gemm.cu:412-521) in ONE launch.  a_slices / b_slices hold two planes
 *   each (real part, imaginary part, split independently: src/split.cu:69-152) a_plane_bytes / b_plane_bytes apart,
 *   amax / bmax two row-scale vectors amax_plane / bmax_plane doubles apart; c is cuDoubleComplex*, ldc / c_batch in
 *   complex elements.  Every tile runs the reference's four real plane products in its order (im,im) -> -alpha,
 *   (re,re) -> +alpha, (im,re), (re,im) -> i*alpha (:479-518), each folded into C with y = fma(x, coef, y)
 *   (axy_complex, :160-186) after C = beta*C (init_c_complex, :188-239; SURVEY App. B.6 fixed: Im uses the original Re).
 *   alpha_dev / beta_dev (both or neither): the scalars are read from device memory in stream order (cuBLAS device
 *   pointer mode; 1 double each, 2 for complex) and alpha[] / beta[] are ignored. */
typedef struct {
  size_t m, n, k, pitch;
  const int8_t *a_slices, *b_slices;   /* base of the operands' slices (entry 0, real plane) */
  size_t a_plane_rows, b_plane_rows;   /* rows the slice planes were split for (0: m / n) */
  size_t row0, col0;                   /* origin of the m x n block inside the planes, multiples of 256 */
  const double *amax, *bmax;           /* row scales of the block's first row / column */
  unsigned num_split, bits_per_int8;
  int complex_c;
  size_t a_plane_bytes, b_plane_bytes, amax_plane, bmax_plane;
  double alpha[2], beta[2];
  const double *alpha_dev, *beta_dev;
  void *c;                             /* the block's first element */
  size_t ldc;
  size_t batch;                        /* 0 or 1: a single product */
  size_t a_batch_bytes, b_batch_bytes, amax_batch, bmax_batch, c_batch;
  unsigned flags;                      /* OZK_FUSED_* */
} ozk_fused_args_t;
int ozk_gemm_i8_fused_ex(const ozk_fused_args_t *args, void *stream);

/* Degenerate k == 0 product: C = beta * C (C not read when beta == 0; reference
 * src/gemm.cu:143-147 applied to an all-zero accumulator). */
int ozk_scale_c(size_t m, size_t n, double beta, double *c, size_t ldc, void *stream);
/* the same for complex C (complex_c != 0: beta = {re, im}, c cuDoubleComplex*, ldc in complex elements: the beta
 * pre-scale of the complex path) and / or beta read from device memory (beta_dev != NULL) */
int ozk_scale_c_ex(size_t m, size_t n, const double beta[2], const double *beta_dev, int complex_c, void *c,
                   size_t ldc, void *stream);

/* Complex GEMM, second form: the four plane products computed as ordinary REAL products into scratch (alpha = 1,
 * beta = 0), then folded into C here with exactly the operations and order of the fused complex epilogue (reference
 * src/gemm.cu:160-239,479-518).  x4 = [4][n][m] doubles, plane (A plane) + 2 * (B plane) with 0 = real, 1 = imaginary.
 * The host picks this form when 4 x as many, 4 x shorter work items fill the GPU's CTA pairs in fewer rounds. */
int ozk_zgemm_combine(size_t m, size_t n, const double *x4, const double alpha[2], const double beta[2],
                      const double *alpha_dev, const double *beta_dev, void *c, size_t ldc, void *stream);

/* Test/tuning hook: force the tile of the fused kernel: (0, w), w in {128, 192, 208, 224, 240, 256} = 256 rows x w
 * columns per CTA pair; (64, 128) = 128 x 128 (64 rows per CTA, UMMA M = 128: the small-problem tile); anything else
 * restores the per-problem choice.  OZIMMU_B200_TILE_N=w forces a width from the environment. */
int ozk_set_cluster_shape(int cm, int cn);

/* Diagnostic (no GPU needed): the tile the fused kernel would use for an m x n x k problem of `batch` entries --
 * rows of C per CTA pair (256 or 128) and tile width -- from the measured per-round cost table (persistent launch:
 * first + (rounds - 1) * next; one_tile_per_pair != 0: SM time of the tiles).  sms <= 0: the current device's SM count
 * (148 without a device).  Ignores ozk_set_cluster_shape / OZIMMU_B200_TILE_N.  Returns 0, or 1 for a null pointer. */
int ozk_fused_tile_choice(size_t m, size_t n, size_t k, size_t batch, int one_tile_per_pair, int sms, int *rows,
                          int *width);

/* Debug/verification launcher: the raw int32 product of ONE slice pair (1-based ids), written
 * column-major with ld = m -- what the reference's cublasGemmEx call produces
 * (src/gemm.cu:315-329).  Same tcgen05 main loop as ozk_gemm_i8_fused. */
int ozk_gemm_i8_pair(size_t m, size_t n, size_t k, const int8_t *a_slices, const int8_t *b_slices,
                     size_t pitch, unsigned num_split, unsigned a_id, unsigned b_id, int32_t *c_i32,
                     void *stream);

/* reference src/split.cu:302-380 (init counter + calculate_mantissa_loss_kernel):
 * adds, for num_split = 3..18, sum over elements of max(0, (emax+1-e)+53 - num_split*bits)
 * into counters[16] (device, uint64).  Does not zero the counters. */
int ozk_mantissa_loss(unsigned long long *counters16, uint32_t *scratch, size_t rows, size_t len,
                      const double *in, size_t ld, int col_major, unsigned bits_per_int8,
                      void *stream);

/* ozk_mantissa_loss on one plane of an interleaved complex matrix (elem_stride = 2). */
int ozk_mantissa_loss_strided(unsigned long long *counters16, uint32_t *scratch, size_t rows, size_t len,
                              const double *in, size_t ld, int col_major, unsigned bits_per_int8,
                              unsigned elem_stride, void *stream);

/* ozk_mantissa_loss_strided for `batch` operands of the same shape in one launch: counters16 is [batch][16], entry e
 * reads in + e*in_stride (doubles) and uses scratch + e*scr_stride.  batch <= 65535. */
int ozk_mantissa_loss_batched(unsigned long long *counters16, uint32_t *scratch, size_t scr_stride, size_t rows,
                              size_t len, const double *in, size_t ld, size_t in_stride, int col_major,
                              unsigned bits_per_int8, unsigned elem_stride, size_t batch, void *stream);

/* ---------------------------------------------------------------------------------------
 * 2. Host API (C spelling of reference include/ozimmu/ozimmu.hpp)
 * ------------------------------------------------------------------------------------- */
int ozimmu_create(ozimmu_handle_t *handle, int malloc_mode);                     /* :47  */
int ozimmu_destroy(ozimmu_handle_t handle);                                      /* :48  */
int ozimmu_set_cuda_stream(ozimmu_handle_t handle, void *stream);                /* :49-50 */
int ozimmu_enable_profiling(ozimmu_handle_t handle);                             /* :52  */
int ozimmu_disable_profiling(ozimmu_handle_t handle);                            /* :53  */
int ozimmu_print_profiler_result(ozimmu_handle_t handle, const char *tag, int csv); /* :54-55 */
int ozimmu_clear_profiler_result(ozimmu_handle_t handle);                        /* :56  */
int ozimmu_set_auto_mantissa_loss_threshold(ozimmu_handle_t handle, double t);   /* :58-59 */
double ozimmu_get_auto_mantissa_loss_threshold(ozimmu_handle_t handle);          /* :60  */

/* :68-74 -- returns the new size in bytes if the workspace grew, else 0 */
size_t ozimmu_reallocate_working_memory(ozimmu_handle_t handle, int op_a, int op_b, size_t m,
                                        size_t n, size_t k, int element_kind, int compute_mode);
size_t ozimmu_reallocate_working_memory_bytes(ozimmu_handle_t handle, size_t size_in_byte);

/* :76-83 -- alpha/beta are HOST pointers (double for real, double[2] for complex); a/b/c are
 * device pointers; BLAS column-major. 0 ok, 1 invalid argument, <0 CUDA failure. */
int ozimmu_gemm(ozimmu_handle_t handle, int op_a, int op_b, size_t m, size_t n, size_t k,
                const void *alpha, const void *a, size_t lda, const void *b, size_t ldb,
                const void *beta, void *c, size_t ldc, int compute_mode, int element_kind);

/* Strided batch of real DGEMMs (what the reference's cublasGemmStridedBatchedEx / cublasDgemmStridedBatched
 * interposers compute entry by entry, src/cublas.cu:315-492): C_e = alpha*op(A_e)*op(B_e) + beta*C_e with
 * X_e = X + e*stride_x (strides in doubles).  compute_mode: fp64_int8_3..18, fp64_int8_auto (decided per
 * entry) or dgemm.  Entries are split on two streams and multiplied by one grouped launch; every entry is
 * bit-identical to a separate ozimmu_gemm call. */
int ozimmu_gemm_strided_batched(ozimmu_handle_t handle, int op_a, int op_b, size_t m, size_t n, size_t k,
                                const double *alpha, const double *a, size_t lda, long long stride_a,
                                const double *b, size_t ldb, long long stride_b, const double *beta, double *c,
                                size_t ldc, long long stride_c, size_t batch, int compute_mode);

/* The same for real or complex data (element_kind; alpha / beta as for ozimmu_gemm; strides in complex elements for
 * complex data -- what cublasZgemmStridedBatched computes, reference src/cublas.cu:494-512).  A complex batch is one
 * grouped launch as well: every tile runs its four plane products back to back. */
int ozimmu_gemm_strided_batched_ex(ozimmu_handle_t handle, int op_a, int op_b, size_t m, size_t n, size_t k,
                                   const void *alpha, const void *a, size_t lda, long long stride_a, const void *b,
                                   size_t ldb, long long stride_b, const void *beta, void *c, size_t ldc,
                                   long long stride_c, size_t batch, int compute_mode, int element_kind);

/* on_device != 0: alpha / beta of the following ozimmu_gemm / ozimmu_gemm_strided_batched* calls are DEVICE pointers
 * (cuBLAS device pointer mode).  The fp64_int8_S modes read them on the device in stream order; the other modes
 * fetch them with one blocking read.  The reference dereferences them on the host regardless (src/gemm.cu:405). */
int ozimmu_set_scalar_pointer_mode(ozimmu_handle_t handle, int on_device);

/* ozimmu_gemm (real) when B becomes valid column panel by column panel -- the multi-GPU path broadcasts B that
 * way -- so that split(A) and the products of the panels that have landed overlap the rest of the
 * transfer.  Panel p = columns [col_edges[p], col_edges[p+1]) of op(B) and C: col_edges has num_panels + 1
 * entries, col_edges[0] = 0, col_edges[num_panels] = n, inner edges multiples of 256, num_panels <= 16;
 * ready_events[p] is a cudaEvent_t the caller recorded (on any stream) after panel p was written.
 * Bit-identical to ozimmu_gemm; asynchronous on the handle's stream. */
int ozimmu_gemm_streamed_b(ozimmu_handle_t handle, int op_a, int op_b, size_t m, size_t n, size_t k,
                           const double *alpha, const double *a, size_t lda, const double *b, size_t ldb,
                           const double *beta, double *c, size_t ldc, int compute_mode, size_t num_panels,
                           const size_t *col_edges, void *const *ready_events);

/* :85-94 -- returns the selected compute mode (OZIMMU_FP64_INT8_3.. or OZIMMU_DGEMM);
 * negative on failure.  counters16 (host, optional) receives the 16 loss totals. */
int ozimmu_auto_mode_select(ozimmu_handle_t handle, int op_a, int op_b, size_t m, size_t n,
                            size_t k, const void *a, size_t lda, const void *b, size_t ldb,
                            int element_kind, double mantissa_loss_threshold,
                            unsigned long long *counters16);

const char *ozimmu_get_compute_mode_name_str(int compute_mode);                  /* :96  */
uint32_t ozimmu_get_bits_per_int8(uint32_t k);                                   /* :102 */

/* Same as ozimmu_gemm but a/b/c are HOST buffers (pinned for full speed): stages H2D copies,
 * split, products and the D2H copy of C on internal streams and returns when C is complete.
 * This is the end-to-end entry `bench.py` times as "e2e". */
int ozimmu_gemm_host(ozimmu_handle_t handle, int op_a, int op_b, size_t m, size_t n, size_t k,
                     const double *alpha, const double *a, size_t lda, const double *b, size_t ldb,
                     const double *beta, double *c, size_t ldc, int compute_mode);

/* ---------------------------------------------------------------------------------------
 * 3. Multi-GPU (one process per GPU; not in the reference, which is single-GPU -- SURVEY 8e, BASELINE config 4)
 *
 * Rank g owns a row block of A and of C (ozimmu_row_block), B is replicated from its owner rank by NCCL broadcasts
 * over NVLink / NVSwitch, every rank runs the single-GPU path on its block: no reduction between ranks (K is never
 * split), so each block is bit-identical to the same rows of a single-GPU ozimmu_gemm.  NCCL is resolved at run time
 * (the libnccl the process has loaded, else libnccl.so.2); the library has no link-time dependency on it.
 * ------------------------------------------------------------------------------------- */
typedef struct ozimmu_comm_s *ozimmu_comm_t;
/* rank 0: 128 bytes to hand to every rank (ncclGetUniqueId) by whatever means the application has */
int ozimmu_comm_unique_id(void *id128);
/* collective: this process joins as `rank` of `nranks` on its current CUDA device (ncclCommInitRank) */
int ozimmu_comm_create(ozimmu_comm_t *comm, int nranks, int rank, const void *id128);
/* wrap a ncclComm_t the application already owns (not destroyed by ozimmu_comm_destroy) */
int ozimmu_comm_adopt(ozimmu_comm_t *comm, void *nccl_comm);
int ozimmu_comm_destroy(ozimmu_comm_t comm);
int ozimmu_comm_rank(ozimmu_comm_t comm);
int ozimmu_comm_size(ozimmu_comm_t comm);
/* rows [*row0, *row0 + *rows) of an m-row matrix that rank `rank` of `nranks` owns: blocks of ceil(m / nranks) rows,
 * the last one short (possibly empty) */
void ozimmu_row_block(size_t m, int nranks, int rank, size_t *row0, size_t *rows);

/* C_block = alpha * op(A_block) * op(B) + beta * C_block on every rank, device operands, asynchronous on the handle's
 * stream.  a_block / c_block: this rank's m_local rows (column-major); b: a device buffer for the full B on every
 * rank whose CONTENT is taken from rank src_rank (the broadcast overwrites the other ranks' copies).  max_panels > 1
 * (op_n B, fp64_int8_S): B travels in up to max_panels column panels on the communicator's stream and each panel of C
 * starts as soon as its columns have landed; max_panels <= 1: one broadcast, then one product launch.  Collective. */
int ozimmu_gemm_sharded(ozimmu_handle_t handle, ozimmu_comm_t comm, int op_a, int op_b, size_t m_local, size_t n, size_t k,
                        const double *alpha, const double *a_block, size_t lda, double *b, size_t ldb, const double *beta,
                        double *c_block, size_t ldc, int compute_mode, int src_rank, unsigned max_panels);

/* The same with HOST operands (pinned for full speed): a_block / c_block are this rank's rows in host memory, b is read
 * on rank src_rank only (NULL elsewhere).  Every rank uploads its blocks of A over its own PCIe link; the owner uploads
 * B block by block in between and broadcasts each block over NVLink as it lands; products start per block pair as in
 * ozimmu_gemm_host.  Returns when this rank's block of C is complete.  Collective. */
int ozimmu_gemm_sharded_host(ozimmu_handle_t handle, ozimmu_comm_t comm, int op_a, int op_b, size_t m_local, size_t n,
                             size_t k, const double *alpha, const double *a_block, size_t lda, const double *b, size_t ldb,
                             const double *beta, double *c_block, size_t ldc, int compute_mode, int src_rank);

/* Diagnostic: the column-panel boundaries ozimmu_gemm_sharded broadcasts B in (ascending, edges[0] = 0, last = n, inner
 * edges multiples of 256).  Writes min(count, capacity) entries, returns count. */
size_t ozimmu_sharded_panel_edges(size_t n, size_t max_panels, size_t *edges, size_t capacity);

/* Diagnostic: the block boundaries ozimmu_gemm_host cuts an operand of `extent` rows into for a requested block
 * edge `want` (0 = one block) -- ascending, edges[0] = 0, last = extent, inner edges multiples of 256, at most 17
 * entries.  Writes min(count, capacity) entries, returns count. */
size_t ozimmu_host_block_edges(size_t extent, size_t want, int taper, size_t *edges, size_t capacity);

/* Number of kernels this library launched since load (bench.py's gpu_launches). */
unsigned long long ozimmu_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* OZIMMU_B200_H */
