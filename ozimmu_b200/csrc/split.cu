// split.cu -- C-ABI of the split / mantissa-loss kernels (kernels: split_kernels.cuh).  The per-split-count
// template instantiations are compiled in two halves (split_inst_lo.cu: 3..10, split_inst_hi.cu: 11..18) so
// the build parallelises.
#include "split_kernels.cuh"

namespace oz {
int split_dispatch_lo(int8_t *out, size_t pitch, size_t plane_rows, double *max_exp, uint32_t *scratch, size_t rows,
                      size_t len, const double *in, size_t ld, int col_major, unsigned num_split, unsigned L,
                      uint32_t es, cudaStream_t stream, const SplitBatch &bt);
int split_dispatch_hi(int8_t *out, size_t pitch, size_t plane_rows, double *max_exp, uint32_t *scratch, size_t rows,
                      size_t len, const double *in, size_t ld, int col_major, unsigned num_split, unsigned L,
                      uint32_t es, cudaStream_t stream, const SplitBatch &bt);
}  // namespace oz

extern "C" uint32_t ozk_bits_per_int8(uint32_t k) {
  if (k == 0) return 0;
  uint32_t lg = 0;
  while (lg < 31 && (1u << (lg + 1)) <= k) lg++;
  if ((1u << lg) != k) lg++;
  const uint32_t v = (31 - lg) / 2;
  return v < 7 ? v : 7;
}

extern "C" size_t ozk_slice_pitch(size_t k) { return oz::slice_pitch(k); }

extern "C" size_t ozk_slices_bytes(size_t rows, size_t k, unsigned num_split) { return oz::slices_bytes(rows, k, num_split); }

// Rows [row0, row0 + rows) of an operand whose slice planes hold plane_rows rows: `out` is the base of the
// operand's slices, `max_exp` / `scratch` / `in` point at row row0's entries.  row0 must be a multiple of 256 and
// the block must end on a multiple of 256 or at the end of the plane (so that blocks never share a row tile).
namespace oz {
namespace {
int split_block_impl(int8_t *out, size_t pitch, size_t plane_rows, size_t row0, double *max_exp, uint32_t *scratch,
                     size_t rows, size_t len, const double *in, size_t ld, int col_major, unsigned num_split,
                     unsigned bits_per_int8, unsigned elem_stride, cudaStream_t s, const SplitBatch &bt) {
  if (rows == 0 || len == 0 || bt.count == 0) return 0;
  if (pitch % 128 != 0 || pitch < len || bits_per_int8 == 0 || bits_per_int8 > 7 || elem_stride < 1 ||
      elem_stride > 2 || num_split < 3 || num_split > 18 || len > 0xFFFFFFF0ull || rows > 0x7FFFFFFFull ||
      (col_major && scratch == nullptr) || row0 % 256 != 0 || row0 + rows > plane_rows ||
      ((row0 + rows) % 256 != 0 && row0 + rows != plane_rows) || bt.count > 65535)
    return static_cast<int>(cudaErrorInvalidValue);
  // rows are padded to a multiple of 256 per slice; the GEMM kernel reads the padding, so it must be zero:
  // the block that ends the plane clears everything from the first partially filled 128-row tile to the end of
  // each slice plane (the kernels below then overwrite the valid rows).  Nothing to do when the plane's row
  // count is a multiple of 256.
  const size_t row_tiles = slice_row_tiles(plane_rows), full_tiles = (row0 + rows) / kTileRows;
  const size_t tile_row_bytes = pitch * kTileRows;  // one 128-row tile across all of K
  if (row0 + rows == plane_rows && full_tiles < row_tiles) {
    for (unsigned t = 0; t < num_split; t++) {
      int8_t *pad = out + (t * row_tiles + full_tiles) * tile_row_bytes;
      if (bt.count == 1)
        OZ_CUDA_TRY(cudaMemsetAsync(pad, 0, (row_tiles - full_tiles) * tile_row_bytes, s));
      else
        OZ_CUDA_TRY(cudaMemset2DAsync(pad, bt.out_stride, 0, (row_tiles - full_tiles) * tile_row_bytes, bt.count, s));
    }
  }
  int8_t *dst = out + (row0 / kTileRows) * tile_row_bytes;
  if (num_split <= 10)
    return split_dispatch_lo(dst, pitch, plane_rows, max_exp, scratch, rows, len, in, ld, col_major, num_split,
                             bits_per_int8, elem_stride, s, bt);
  return split_dispatch_hi(dst, pitch, plane_rows, max_exp, scratch, rows, len, in, ld, col_major, num_split,
                           bits_per_int8, elem_stride, s, bt);
}
}  // namespace
}  // namespace oz

// Rows [row0, row0 + rows) of an operand whose slice planes hold plane_rows rows: `out` is the base of the
// operand's slices, `max_exp` / `scratch` / `in` point at row row0's entries.  row0 must be a multiple of 256 and
// the block must end on a multiple of 256 or at the end of the plane (so that blocks never share a row tile).
extern "C" int ozk_split_int8_block(int8_t *out, size_t pitch, size_t plane_rows, size_t row0, double *max_exp,
                                    uint32_t *scratch, size_t rows, size_t len, const double *in, size_t ld,
                                    int col_major, unsigned num_split, unsigned bits_per_int8,
                                    unsigned elem_stride, void *stream) {
  return oz::split_block_impl(out, pitch, plane_rows, row0, max_exp, scratch, rows, len, in, ld, col_major, num_split,
                              bits_per_int8, elem_stride, static_cast<cudaStream_t>(stream),
                              oz::SplitBatch{1, 0, 0, 0, 0});
}

// `batch` operands of the same shape in one launch: entry e reads in + e*in_stride (doubles) and writes its slices
// at out + e*out_stride (bytes), its row scales at max_exp + e*max_stride, scratch at scratch + e*scr_stride.
extern "C" int ozk_split_int8_batched_strided(int8_t *out, size_t out_stride, size_t pitch, double *max_exp,
                                              size_t max_stride, uint32_t *scratch, size_t scr_stride, size_t rows,
                                              size_t len, const double *in, size_t ld, size_t in_stride, int col_major,
                                              unsigned num_split, unsigned bits_per_int8, unsigned elem_stride,
                                              size_t batch, void *stream) {
  if (batch == 0) return 0;
  if (batch > 65535) return static_cast<int>(cudaErrorInvalidValue);
  const oz::SplitBatch bt{static_cast<uint32_t>(batch), in_stride, out_stride, max_stride, scr_stride};
  return oz::split_block_impl(out, pitch, rows, 0, max_exp, scratch, rows, len, in, ld, col_major, num_split,
                              bits_per_int8, elem_stride, static_cast<cudaStream_t>(stream), bt);
}

extern "C" int ozk_split_int8_batched(int8_t *out, size_t out_stride, size_t pitch, double *max_exp, size_t max_stride,
                                      uint32_t *scratch, size_t scr_stride, size_t rows, size_t len, const double *in,
                                      size_t ld, size_t in_stride, int col_major, unsigned num_split,
                                      unsigned bits_per_int8, size_t batch, void *stream) {
  return ozk_split_int8_batched_strided(out, out_stride, pitch, max_exp, max_stride, scratch, scr_stride, rows, len, in,
                                        ld, in_stride, col_major, num_split, bits_per_int8, 1, batch, stream);
}

extern "C" int ozk_split_int8_strided(int8_t *out, size_t pitch, double *max_exp, uint32_t *scratch,
                                      size_t rows, size_t len, const double *in, size_t ld, int col_major,
                                      unsigned num_split, unsigned bits_per_int8, unsigned elem_stride,
                                      void *stream) {
  return ozk_split_int8_block(out, pitch, rows, 0, max_exp, scratch, rows, len, in, ld, col_major, num_split,
                              bits_per_int8, elem_stride, stream);
}

extern "C" int ozk_split_int8(int8_t *out, size_t pitch, double *max_exp, uint32_t *scratch,
                              size_t rows, size_t len, const double *in, size_t ld, int col_major,
                              unsigned num_split, unsigned bits_per_int8, void *stream) {
  return ozk_split_int8_strided(out, pitch, max_exp, scratch, rows, len, in, ld, col_major, num_split,
                                bits_per_int8, 1, stream);
}

// counters16: [batch][16]; entry e reads in + e*in_stride (doubles) and uses scratch + e*scr_stride
extern "C" int ozk_mantissa_loss_batched(unsigned long long *counters16, uint32_t *scratch, size_t scr_stride,
                                         size_t rows, size_t len, const double *in, size_t ld, size_t in_stride,
                                         int col_major, unsigned bits_per_int8, unsigned elem_stride, size_t batch,
                                         void *stream) {
  if (rows == 0 || len == 0 || batch == 0) return 0;
  if (len > 0xFFFFFFF0ull || rows > 0x7FFFFFFFull || (col_major && scratch == nullptr) || elem_stride < 1 ||
      elem_stride > 2 || batch > 65535)
    return static_cast<int>(cudaErrorInvalidValue);
  const uint32_t es = elem_stride;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  using namespace oz;
  const unsigned nb = static_cast<unsigned>(batch);
  if (col_major) {
    const SplitBatch bt{nb, in_stride, 0, 0, scr_stride};
    if (batch == 1) OZ_CUDA_TRY(cudaMemsetAsync(scratch, 0, rows * sizeof(uint32_t), s));
    else OZ_CUDA_TRY(cudaMemset2DAsync(scratch, scr_stride * sizeof(uint32_t), 0, rows * sizeof(uint32_t), batch, s));
    dim3 g(static_cast<unsigned>((rows + 255) / 256), ceil_div_u32(static_cast<uint32_t>(len), kColChunk), nb);
    rowmax_cols_kernel<<<g, 256, 0, s>>>(scratch, rows, static_cast<uint32_t>(len), in, ld, es, bt);
    loss_cols_kernel<<<g, 256, 0, s>>>(counters16, scratch, rows, static_cast<uint32_t>(len), in, ld,
                                       bits_per_int8, es, bt);
    count_launch(2);
  } else {
    loss_rows_kernel<<<dim3(static_cast<unsigned>(rows), nb), kSplitThreads, 0, s>>>(
        counters16, static_cast<uint32_t>(len), in, ld, bits_per_int8, es, in_stride);
    count_launch(1);
  }
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ozk_mantissa_loss_strided(unsigned long long *counters16, uint32_t *scratch, size_t rows,
                                         size_t len, const double *in, size_t ld, int col_major,
                                         unsigned bits_per_int8, unsigned elem_stride, void *stream) {
  return ozk_mantissa_loss_batched(counters16, scratch, 0, rows, len, in, ld, 0, col_major, bits_per_int8, elem_stride,
                                   1, stream);
}

extern "C" int ozk_mantissa_loss(unsigned long long *counters16, uint32_t *scratch, size_t rows,
                                 size_t len, const double *in, size_t ld, int col_major,
                                 unsigned bits_per_int8, void *stream) {
  return ozk_mantissa_loss_strided(counters16, scratch, rows, len, in, ld, col_major, bits_per_int8, 1, stream);
}
