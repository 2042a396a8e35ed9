#include "inst_bezier.cuh"
namespace enterp {
template <>
cudaError_t launch_bezier<double>(int dh, bool proj, int n, const CurveDev<double>& c, const SampleSrc<double>& s, double* out,
                              bool vec_ok, cudaStream_t st) {
  return launch_bezier_impl<double>(dh, proj, n, c, s, out, vec_ok, st);
}
}  // namespace enterp
