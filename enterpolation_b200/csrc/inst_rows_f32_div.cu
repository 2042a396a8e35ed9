#include "inst_rows.cuh"
namespace enterp {
template <>
cudaError_t launch_bspline_rows_mode<float, false>(int dh, bool proj, int deg, bool equi, const CurveDev<float>& c,
                                               const SampleSrc<float>& s, float* out, bool vec_ok, const RowsDev<float>& rw,
                                               cudaStream_t st) {
  return launch_rows_impl<float, false>(dh, proj, deg, equi, c, s, out, vec_ok, rw, st);
}
}  // namespace enterp
