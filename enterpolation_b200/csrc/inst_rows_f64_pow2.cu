#include "inst_rows.cuh"
namespace enterp {
template <>
cudaError_t launch_bspline_rows_mode<double, true>(int dh, bool proj, int deg, bool equi, const CurveDev<double>& c,
                                               const SampleSrc<double>& s, double* out, bool vec_ok, const RowsDev<double>& rw,
                                               cudaStream_t st) {
  return launch_rows_impl<double, true>(dh, proj, deg, equi, c, s, out, vec_ok, rw, st);
}
}  // namespace enterp
