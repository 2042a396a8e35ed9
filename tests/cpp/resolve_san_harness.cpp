// Entry point of tests/test_oracle_sanitized.py::test_host_resolution_is_clean_under_asan_and_ubsan: the product's
// host-side descriptor resolution (enterpolation_b200/csrc/resolve.cpp: validation, knot tables, span rows, bucket index,
// slot tables) compiled with AddressSanitizer + UBSan and fed the fuzz generators' descriptors. No CUDA involved.
#include "../../enterpolation_b200/csrc/resolve.hpp"

extern "C" __attribute__((visibility("default"))) int resolve_san(const enterp_curve_desc* d) {
  enterp::Resolved r;
  enterp_error_info e{};
  return (int)enterp::resolve(*d, r, &e);
}
