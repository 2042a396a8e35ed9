#include "inst_bezier.cuh"
namespace enterp {
template <>
cudaError_t launch_bezier<float>(int dh, bool proj, int n, const CurveDev<float>& c, const SampleSrc<float>& s, float* out,
                              bool vec_ok, cudaStream_t st) {
  return launch_bezier_impl<float>(dh, proj, n, c, s, out, vec_ok, st);
}
}  // namespace enterp
