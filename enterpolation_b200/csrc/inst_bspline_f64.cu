#include "inst_bspline.cuh"
namespace enterp {
template <>
cudaError_t launch_bspline<double>(int dh, bool proj, int deg, bool equi, const CurveDev<double>& c, const SampleSrc<double>& s,
                               double* out, bool vec_ok, cudaStream_t st) {
  return launch_bspline_impl<double>(dh, proj, deg, equi, c, s, out, vec_ok, st);
}
}  // namespace enterp
