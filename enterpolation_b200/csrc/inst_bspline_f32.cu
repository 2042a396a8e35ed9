#include "inst_bspline.cuh"
namespace enterp {
template <>
cudaError_t launch_bspline<float>(int dh, bool proj, int deg, bool equi, const CurveDev<float>& c, const SampleSrc<float>& s,
                               float* out, bool vec_ok, cudaStream_t st) {
  return launch_bspline_impl<float>(dh, proj, deg, equi, c, s, out, vec_ok, st);
}
}  // namespace enterp
