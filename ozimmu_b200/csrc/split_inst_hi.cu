// split_inst_hi.cu -- instantiations of the split kernels for num_split = 11..18 (see split.cu)
#include "split_kernels.cuh"

namespace oz {
int split_dispatch_hi(int8_t *out, size_t pitch, size_t plane_rows, double *max_exp, uint32_t *scratch, size_t rows, size_t len,
                      const double *in, size_t ld, int col_major, unsigned num_split, unsigned L, uint32_t es,
                      cudaStream_t stream, const SplitBatch &bt) {
  switch (num_split) {
    case 11: return launch_split<11>(out, pitch, plane_rows, max_exp, scratch, rows, len, in, ld, col_major, L, es, stream, bt);
    case 12: return launch_split<12>(out, pitch, plane_rows, max_exp, scratch, rows, len, in, ld, col_major, L, es, stream, bt);
    case 13: return launch_split<13>(out, pitch, plane_rows, max_exp, scratch, rows, len, in, ld, col_major, L, es, stream, bt);
    case 14: return launch_split<14>(out, pitch, plane_rows, max_exp, scratch, rows, len, in, ld, col_major, L, es, stream, bt);
    case 15: return launch_split<15>(out, pitch, plane_rows, max_exp, scratch, rows, len, in, ld, col_major, L, es, stream, bt);
    case 16: return launch_split<16>(out, pitch, plane_rows, max_exp, scratch, rows, len, in, ld, col_major, L, es, stream, bt);
    case 17: return launch_split<17>(out, pitch, plane_rows, max_exp, scratch, rows, len, in, ld, col_major, L, es, stream, bt);
    case 18: return launch_split<18>(out, pitch, plane_rows, max_exp, scratch, rows, len, in, ld, col_major, L, es, stream, bt);
    default: return static_cast<int>(cudaErrorInvalidValue);
  }
}
}  // namespace oz
