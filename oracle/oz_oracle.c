/*
 * oz_oracle.c -- CPU restatement of reference ozIMMU's Ozaki-scheme DGEMM hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (ozimmu_b200/lib/libozimmu.so) never links, loads or calls anything in oracle/.
 *
 * Parity status: PINNED.  The restatement is checked bit-for-bit against golden
 * vectors produced by the UNMODIFIED reference (built from /root/reference by
 * oracle/Makefile into oracle/_ref/ and run on a B200 by tests/golden/make_golden.py);
 * see tests/test_oracle_golden.py.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Plain C99 + unsigned __int128, scalar, single threaded.
 * Build with -ffp-contract=off: every fused multiply-add the GPU code performs is
 * spelled fma() here; everything else must stay unfused.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

static inline uint64_t f2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double u2f(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

#define EXP_MASK  0x7FF0000000000000ULL  /* cutf fp.hpp:86-91 mask_exponent */
#define MANT_MASK 0x000FFFFFFFFFFFFFULL  /* cutf fp.hpp:79-84 mask_mantissa */

/* src/split.cu:520-536  get_bits_per_int8 */
uint32_t oz_bits_per_int8(uint32_t k) {
  if (k == 0) return 0;
  uint32_t lg = 0;
  while ((1u << (lg + 1)) <= k) lg++;
  if ((1u << lg) != k) lg++;
  uint32_t v = (31 - lg) / 2;
  return v < 7 ? v : 7;
}

/* src/config.cu:85-92  ordered (A_id,B_id) pair list; returns the number of pairs */
int oz_pair_list(int num_split, int *a_id, int *b_id) {
  int cnt = 0;
  for (int sum = 2; sum <= num_split + 1; sum++)
    for (int j = 1; j < sum; j++) {
      if (j > num_split || sum - j > num_split) continue;
      if (a_id) { a_id[cnt] = j; b_id[cnt] = sum - j; }
      cnt++;
    }
  return cnt;
}

/* src/utils.hpp:30-39  padded_ld<int8>: round up to a multiple of 4 */
uint32_t oz_slice_ld(uint32_t k) { return ((k + 3) / 4) * 4; }

/* element (row r, position c) of a "row" of op(X): src/split.cu:203,210
 * col_major: in[c*ld + r]   else: in[c + r*ld] */
static inline double oz_at_s(const double *in, size_t ld, int col_major, size_t r, size_t c, size_t es) {
  return col_major ? in[(c * ld + r) * es] : in[(c + r * ld) * es];
}
static inline double oz_at(const double *in, size_t ld, int col_major, size_t r, size_t c) {
  return oz_at_s(in, ld, col_major, r, c, 1);
}

/* src/split.cu:14-67 + :191,202-204: max over the row of the exponent-only value, times 2.
 * (cutf::math::max on doubles -> fmax; operands are never NaN unless exponent==0x7FF with
 *  mantissa cleared => +Inf, so fmax semantics are irrelevant.) */
double oz_row_max_exp_s(const double *in, size_t ld, int col_major, size_t r, size_t len, size_t es) {
  double mx = 0.0;
  for (size_t c = 0; c < len; c++) {
    double v = u2f(f2u(oz_at_s(in, ld, col_major, r, c, es)) & EXP_MASK);
    if (v > mx) mx = v;
  }
  return mx * 2.0;
}
double oz_row_max_exp(const double *in, size_t ld, int col_major, size_t r, size_t len) {
  return oz_row_max_exp_s(in, ld, col_major, r, len, 1);
}

/* src/split.cu:155-185 cut_int8_core<double,__uint128_t>.  out[t*inc] for t<num_split.
 * The GPU lowering of the 128-bit right shift yields 0 for shift amounts >= 128
 * (SURVEY App. C.2); C leaves that undefined, so it is clamped explicitly. */
void oz_cut_int8(int8_t *out, size_t inc, double a, double max_exp, unsigned num_split,
                 unsigned L) {
  const int sign = (a > 0) ? 1 : -1;
  const uint64_t ea = f2u(a) & EXP_MASK;
  const uint64_t implicit = ea ? 1ull : 0ull;
  const u128 mant = ((u128)((f2u(a) & MANT_MASK) | (implicit << 52))) << (64 + 11);
  const uint64_t off = (f2u(max_exp) - ea) >> 52;
  u128 sh = (off >= 128) ? (u128)0 : (mant >> off);
  for (unsigned t = 0; t < num_split; t++) {
    const int8_t top = (int8_t)(sh >> (128 - L));
    out[t * inc] = (int8_t)(top * sign);
    sh <<= L;
  }
}

/* src/split.cu:193-242 split_int8_kernel (real) + :244-283 host wrappers.
 * rows x len view of op(X); out is [num_split][rows][ldo] int8, K contiguous, columns
 * len..ldo-1 zero-filled (:222-232); max_exp[rows] (:234-241). */
/* es = 1: real matrix.  es = 2: one plane of an interleaved complex matrix (src/split.cu:69-152,
 * 211-216: each plane has its own row maximum and is cut independently); `in` points at the plane. */
void oz_split_s(int8_t *out, uint32_t ldo, double *max_exp, size_t rows, size_t len,
                const double *in, size_t ld, int col_major, unsigned num_split, unsigned L, size_t es) {
  const size_t N = rows * (size_t)ldo;
  for (size_t r = 0; r < rows; r++) {
    const double mx = oz_row_max_exp_s(in, ld, col_major, r, len, es);
    for (size_t c = 0; c < len; c++)
      oz_cut_int8(out + r * ldo + c, N, oz_at_s(in, ld, col_major, r, c, es), mx, num_split, L);
    for (size_t c = len; c < ldo; c++)
      for (unsigned t = 0; t < num_split; t++) out[r * ldo + c + t * N] = 0;
    max_exp[r] = mx;
  }
}
void oz_split(int8_t *out, uint32_t ldo, double *max_exp, size_t rows, size_t len,
              const double *in, size_t ld, int col_major, unsigned num_split, unsigned L) {
  oz_split_s(out, ldo, max_exp, rows, len, in, ld, col_major, num_split, L, 1);
}

/* src/gemm.cu:315-329: cublasGemmEx(OP_T, OP_N, m, n, k4, 1, A_i(k4 x m), B_j(k4 x n), 0,
 * C_i32 ld=m) -- exact integer product, C column-major. */
void oz_int8_gemm(int32_t *c, size_t m, size_t n, size_t k4, const int8_t *a, const int8_t *b) {
  for (size_t j = 0; j < n; j++)
    for (size_t i = 0; i < m; i++) {
      const int8_t *ar = a + i * k4, *br = b + j * k4;
      int32_t s = 0;
      for (size_t p = 0; p < k4; p++) s += (int32_t)ar[p] * (int32_t)br[p];
      c[i + j * m] = s;
    }
}

/* src/gemm.cu:91-102 scale = 2^(-rshift) built from exponent bits */
static inline double oz_scale(int32_t rshift) {
  return u2f((uint64_t)(0x3ff - rshift) << 52);
}

/* src/gemm.cu:77-89 accumulate_in_f64_kernel; nvcc contracts the += into one DFMA
 * (SURVEY App. C.2). */
void oz_accumulate(double *acc, const int32_t *p, size_t len, int32_t rshift) {
  const double scale = oz_scale(rshift);
  for (size_t t = 0; t < len; t++) {
    const double v = (double)((int64_t)p[t] * 4294967296LL); /* (int64)p << 32 */
    acc[t] = fma(v, scale, acc[t]);
  }
}

/* src/gemm.cu:124-148 axby_kernel: x = acc / 2^44 * amax[mi] * bmax[ni];
 * beta != 0: y = a*x + b*y  (contracted: DMUL b*y, DFMA(a,x,.)), else y = a*x. */
void oz_finalize(size_t m, size_t n, double alpha, const double *acc, double beta, double *c,
                 size_t ldc, const double *amax, const double *bmax) {
  for (size_t j = 0; j < n; j++)
    for (size_t i = 0; i < m; i++) {
      double x = acc[i + j * m] * 0x1p-44;
      x = x * amax[i];
      x = x * bmax[j];
      double *y = c + i + j * ldc;
      if (beta != 0) {
        const double by = beta * *y;
        *y = fma(alpha, x, by);
      } else {
        *y = alpha * x;
      }
    }
}

/* src/gemm.cu:344-410 gemm_int8<double>.  op: 0 = N, 1 = T (BLAS column-major).
 * Optional outputs (may be NULL): slices/max_exp of A and B, for golden comparison.
 * Returns 0, or 1 on allocation failure. */
int oz_gemm(int op_a, int op_b, size_t m, size_t n, size_t k, double alpha, const double *a,
            size_t lda, const double *b, size_t ldb, double beta, double *c, size_t ldc,
            unsigned num_split, int8_t *a_slices_out, int8_t *b_slices_out, double *amax_out,
            double *bmax_out) {
  const unsigned L = oz_bits_per_int8((uint32_t)k);                       /* :357 */
  const uint32_t k4 = oz_slice_ld((uint32_t)k);                           /* :369-372 */
  int8_t *as = (int8_t *)malloc((size_t)num_split * m * k4 + 1);
  int8_t *bs = (int8_t *)malloc((size_t)num_split * n * k4 + 1);
  double *amax = (double *)malloc(sizeof(double) * (m + 1));
  double *bmax = (double *)malloc(sizeof(double) * (n + 1));
  double *acc = (double *)calloc(m * n + 1, sizeof(double));              /* :367 */
  int32_t *ci = (int32_t *)malloc(sizeof(int32_t) * (m * n + 1));
  int pa[200], pb[200];
  if (!as || !bs || !amax || !bmax || !acc || !ci) return 1;
  /* A: rows of op(A); op_n => column-major source (src/split.cu:254).
   * B: wrapper swaps (m,n) and flips op (src/split.cu:277-282) => "rows" are columns of op(B):
   *    op_n B (k x n col-major, ld=ldb): col j contiguous => not col_major. */
  oz_split(as, k4, amax, m, k, a, lda, op_a == 0, num_split, L);
  oz_split(bs, k4, bmax, n, k, b, ldb, op_b != 0, num_split, L);
  const int np = oz_pair_list((int)num_split, pa, pb);
  for (int p = 0; p < np; p++) {                                          /* :387-403 */
    oz_int8_gemm(ci, m, n, k4, as + (size_t)(pa[p] - 1) * m * k4,
                 bs + (size_t)(pb[p] - 1) * n * k4);
    oz_accumulate(acc, ci, m * n, (int32_t)L * (pa[p] + pb[p] - 2) - (7 - (int32_t)L) * 2);
  }
  oz_finalize(m, n, alpha, acc, beta, c, ldc, amax, bmax);                /* :405 */
  if (a_slices_out) memcpy(a_slices_out, as, (size_t)num_split * m * k4);
  if (b_slices_out) memcpy(b_slices_out, bs, (size_t)num_split * n * k4);
  if (amax_out) memcpy(amax_out, amax, sizeof(double) * m);
  if (bmax_out) memcpy(bmax_out, bmax, sizeof(double) * n);
  free(as); free(bs); free(amax); free(bmax); free(acc); free(ci);
  return 0;
}

/* src/gemm.cu:412-521 gemm_int8<cuDoubleComplex> with :160-186 axy_complex_kernel and :188-239
 * init_c_complex (FMA contraction as in the reference's sm_100 SASS: t = c.y*b.y; c.x = fma(c.x, b.x, -t);
 * t = c.x*b.y; c.y = fma(c.y, b.x, t)).  The reference reads the UPDATED c.x in the second product (an aliasing
 * bug, SURVEY App. B.6: Im(beta*C) comes out wrong whenever Im(beta) != 0).  ref_beta_quirk != 0 reproduces that
 * (used to pin this file against the reference's golden vectors); 0 is the corrected arithmetic the product ships
 * (the two agree bit for bit when Im(beta) == 0).
 * a, b, c: interleaved (re, im) doubles; lda/ldb/ldc in complex elements; alpha/beta: {re, im}. */
int oz_gemm_complex_q(int op_a, int op_b, size_t m, size_t n, size_t k, const double *alpha, const double *a,
                      size_t lda, const double *b, size_t ldb, const double *beta, double *c, size_t ldc,
                      unsigned num_split, int ref_beta_quirk) {
  const unsigned L = oz_bits_per_int8((uint32_t)k);
  const uint32_t k4 = oz_slice_ld((uint32_t)k);
  const size_t a_plane = (size_t)num_split * m * k4, b_plane = (size_t)num_split * n * k4;
  int8_t *as = (int8_t *)malloc(2 * a_plane + 1);
  int8_t *bs = (int8_t *)malloc(2 * b_plane + 1);
  double *amax = (double *)malloc(sizeof(double) * (2 * m + 1));
  double *bmax = (double *)malloc(sizeof(double) * (2 * n + 1));
  double *acc = (double *)malloc(sizeof(double) * (m * n + 1));
  int32_t *ci = (int32_t *)malloc(sizeof(int32_t) * (m * n + 1));
  int pa[200], pb[200];
  if (!as || !bs || !amax || !bmax || !acc || !ci) return 1;
  for (int part = 0; part < 2; part++) {
    oz_split_s(as + part * a_plane, k4, amax + part * m, m, k, a + part, lda, op_a == 0, num_split, L, 2);
    oz_split_s(bs + part * b_plane, k4, bmax + part * n, n, k, b + part, ldb, op_b != 0, num_split, L, 2);
  }
  /* init_c_complex */
  for (size_t j = 0; j < n; j++)
    for (size_t i = 0; i < m; i++) {
      double *y = c + 2 * (i + j * ldc);
      if (beta[0] == 0 && beta[1] == 0) {
        y[0] = 0; y[1] = 0;
      } else {
        const double t = y[1] * beta[1];
        const double x_old = y[0];
        y[0] = fma(y[0], beta[0], -t);
        const double t2 = (ref_beta_quirk ? y[0] : x_old) * beta[1];
        y[1] = fma(y[1], beta[0], t2);
      }
    }
  const int np = oz_pair_list((int)num_split, pa, pb);
  static const int groups[4][2] = {{1, 1}, {0, 0}, {1, 0}, {0, 1}};          /* :479-480 */
  for (int g = 0; g < 4; g++) {
    const int ga = groups[g][0], gb = groups[g][1];
    memset(acc, 0, sizeof(double) * m * n);
    for (int p = 0; p < np; p++) {
      oz_int8_gemm(ci, m, n, k4, as + ga * a_plane + (size_t)(pa[p] - 1) * m * k4,
                   bs + gb * b_plane + (size_t)(pb[p] - 1) * n * k4);
      oz_accumulate(acc, ci, m * n, (int32_t)L * (pa[p] + pb[p] - 2) - (7 - (int32_t)L) * 2);
    }
    double cr, cim;                                                            /* :497-509 */
    if (ga == 0 && gb == 0) { cr = alpha[0]; cim = alpha[1]; }
    else if (ga == 1 && gb == 1) { cr = -alpha[0]; cim = -alpha[1]; }
    else { cr = -alpha[1]; cim = alpha[0]; }
    for (size_t j = 0; j < n; j++)
      for (size_t i = 0; i < m; i++) {
        double x = acc[i + j * m] * 0x1p-44;
        x = x * amax[ga * m + i];
        x = x * bmax[gb * n + j];
        double *y = c + 2 * (i + j * ldc);
        y[0] = fma(x, cr, y[0]);
        y[1] = fma(x, cim, y[1]);
      }
  }
  free(as); free(bs); free(amax); free(bmax); free(acc); free(ci);
  return 0;
}

/* the reference as compiled (beta quirk included) */
int oz_gemm_complex(int op_a, int op_b, size_t m, size_t n, size_t k, const double *alpha, const double *a,
                    size_t lda, const double *b, size_t ldb, const double *beta, double *c, size_t ldc,
                    unsigned num_split) {
  return oz_gemm_complex_q(op_a, op_b, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, num_split, 1);
}

/* src/split.cu:317-380 mantissa-loss totals, INTENDED semantics:
 * counters[16] for num_split = 3..18, accumulated (not reset) so A and B can be chained. */
void oz_mantissa_loss_s(uint64_t counters[16], size_t rows, size_t len, const double *in, size_t ld,
                        int col_major, unsigned L, size_t es) {
  for (size_t r = 0; r < rows; r++) {
    const double mx = oz_row_max_exp_s(in, ld, col_major, r, len, es);
    for (size_t c = 0; c < len; c++) {
      const double x = oz_at_s(in, ld, col_major, r, c, es);
      if (x == 0 || mx == 0) continue;                                    /* :322-324 */
      const uint64_t req = (((f2u(mx) & EXP_MASK) - (f2u(x) & EXP_MASK)) >> 52) + 53;
      for (unsigned s = 3; s <= 18; s++) {
        const uint64_t space = (uint64_t)s * L;
        if (space < req) counters[s - 3] += req - space;
      }
    }
  }
}

void oz_mantissa_loss(uint64_t counters[16], size_t rows, size_t len, const double *in, size_t ld,
                      int col_major, unsigned L) {
  oz_mantissa_loss_s(counters, rows, len, in, ld, col_major, L, 1);
}

/* complex flavour of auto_mode_select_core (src/split.cu:365-372: both planes, each against its own row
 * maximum; denominator counts complex elements, :484-493) */
int oz_auto_select_complex(int op_a, int op_b, size_t m, size_t n, size_t k, const double *a, size_t lda,
                           const double *b, size_t ldb, double threshold, uint64_t *counters_out) {
  uint64_t cnt[16] = {0};
  const unsigned L = oz_bits_per_int8((uint32_t)k);
  for (int part = 0; part < 2; part++) {
    oz_mantissa_loss_s(cnt, m, k, a + part, lda, op_a == 0, L, 2);
    oz_mantissa_loss_s(cnt, n, k, b + part, ldb, op_b != 0, L, 2);
  }
  if (counters_out) memcpy(counters_out, cnt, sizeof(cnt));
  for (int s = 3; s <= 18; s++)
    if ((double)cnt[s - 3] / (double)(m * k + k * n) <= threshold) return s;
  return 0;
}

/* src/split.cu:454-494 auto_mode_select_core: returns chosen num_split (3..18) or 0 for dgemm.
 * counters_out (optional) receives the 16 totals. */
int oz_auto_select(int op_a, int op_b, size_t m, size_t n, size_t k, const double *a, size_t lda,
                   const double *b, size_t ldb, double threshold, uint64_t *counters_out) {
  uint64_t cnt[16] = {0};
  const unsigned L = oz_bits_per_int8((uint32_t)k);
  oz_mantissa_loss(cnt, m, k, a, lda, op_a == 0, L);
  oz_mantissa_loss(cnt, n, k, b, ldb, op_b != 0, L);
  if (counters_out) memcpy(counters_out, cnt, sizeof(cnt));
  for (int s = 3; s <= 18; s++)
    if ((double)cnt[s - 3] / (double)(m * k + k * n) <= threshold) return s;
  return 0;
}
