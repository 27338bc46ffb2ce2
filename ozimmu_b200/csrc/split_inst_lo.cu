// split_inst_lo.cu -- instantiations of the split kernels for num_split = 3..10 (see split.cu)
#include "split_kernels.cuh"

namespace oz {
int split_dispatch_lo(int8_t *out, size_t pitch, size_t plane_rows, double *max_exp, uint32_t *scratch, size_t rows, size_t len,
                      const double *in, size_t ld, int col_major, unsigned num_split, unsigned L, uint32_t es,
                      cudaStream_t stream, const SplitBatch &bt) {
  switch (num_split) {
    case 3: return launch_split<3>(out, pitch, plane_rows, max_exp, scratch, rows, len, in, ld, col_major, L, es, stream, bt);
    case 4: return launch_split<4>(out, pitch, plane_rows, max_exp, scratch, rows, len, in, ld, col_major, L, es, stream, bt);
    case 5: return launch_split<5>(out, pitch, plane_rows, max_exp, scratch, rows, len, in, ld, col_major, L, es, stream, bt);
    case 6: return launch_split<6>(out, pitch, plane_rows, max_exp, scratch, rows, len, in, ld, col_major, L, es, stream, bt);
    case 7: return launch_split<7>(out, pitch, plane_rows, max_exp, scratch, rows, len, in, ld, col_major, L, es, stream, bt);
    case 8: return launch_split<8>(out, pitch, plane_rows, max_exp, scratch, rows, len, in, ld, col_major, L, es, stream, bt);
    case 9: return launch_split<9>(out, pitch, plane_rows, max_exp, scratch, rows, len, in, ld, col_major, L, es, stream, bt);
    case 10: return launch_split<10>(out, pitch, plane_rows, max_exp, scratch, rows, len, in, ld, col_major, L, es, stream, bt);
    default: return static_cast<int>(cudaErrorInvalidValue);
  }
}
}  // namespace oz
