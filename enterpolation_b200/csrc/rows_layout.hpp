// Layout of the span-row table shared by the host builder (resolve.cpp) and the kernel (rows.cuh).
#pragma once
#if defined(__CUDACC__)
#define ENTERP_HD __host__ __device__
#else
#define ENTERP_HD
#endif

namespace enterp {

// Highest degree with span-row kernels. The default build instantiates degrees 1..4 (what the reference's examples,
// benches and BASELINE.json's configs use): a 20 MB library. `make ROWS_MAX_DEGREE=7` adds 5..7: 35 MB, 2:45 instead of
// 1:30 min of nvcc on 8 cores, and take(2^26) of a degree-5 / degree-7 `[f32;4]` curve at 60 / 61 instead of 51 / 36 G
// samples/s (f64 scalar: 84 / 41 instead of 45 / 30). Above the limit B-splines run through the generic kernel
// (compile-time degree up to 7).
#ifndef ENTERP_ROWS_MAX_DEGREE
#define ENTERP_ROWS_MAX_DEGREE 4
#endif
constexpr int ROWS_MAX_DEGREE = ENTERP_ROWS_MAX_DEGREE;
static_assert(ROWS_MAX_DEGREE >= 1 && ROWS_MAX_DEGREE <= 7, "span rows exist for degrees 1..7");
constexpr unsigned ROWS_MAX_BYTES = 64u * 1024u;  // shared-memory budget of the row table per CTA
// Bigger tables stay in global memory and serve take() only (chunked kernel variant: a CTA walks a contiguous
// sample range, so a thread's span changes rarely and the row reloads come out of L2).
constexpr unsigned long long ROWS_GLOBAL_MAX_BYTES = 64ull << 20;

// ---- compile-time schedule of the factor computations ------------------------------------------
// Factors f(r, j), r = 1..P, j = 0..P-r, use numerator index m = j + r - 1 and denominator
// kn[P + j] - kn[m]. With W lanes per operation, neighbours (j, j+1) of one r share an operation;
// the leftover factor of every r with an odd count always has m = P - 1, and leftovers are paired
// with each other (a last unpaired lane duplicates its neighbour).
struct FactorSched {
  int nops;
  int r[28][2];
  int j[28][2];
};

ENTERP_HD constexpr FactorSched make_sched(int P, int W) {
  FactorSched s{};
  int n = 0;
  if (W == 1) {
    for (int r = 1; r <= P; ++r)
      for (int j = 0; j + r <= P; ++j) {
        s.r[n][0] = r; s.j[n][0] = j; s.r[n][1] = r; s.j[n][1] = j;
        ++n;
      }
    s.nops = n;
    return s;
  }
  int sr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int ns = 0;
  for (int r = 1; r <= P; ++r) {
    const int cnt = P - r + 1;
    for (int j = 0; j + 1 < cnt; j += 2) {
      s.r[n][0] = r; s.j[n][0] = j; s.r[n][1] = r; s.j[n][1] = j + 1;
      ++n;
    }
    if (cnt % 2) sr[ns++] = r;
  }
  for (int k = 0; k < ns; k += 2) {
    const int r0 = sr[k], r1 = (k + 1 < ns) ? sr[k + 1] : sr[k];
    s.r[n][0] = r0; s.j[n][0] = P - r0; s.r[n][1] = r1; s.j[n][1] = P - r1;
    ++n;
  }
  s.nops = n;
  return s;
}

ENTERP_HD constexpr int sched_nops(int P, int W) { return make_sched(P, W).nops; }
// operation / lane that produces factor (r, j)
ENTERP_HD constexpr int sched_find(int P, int W, int r, int j) {
  const FactorSched s = make_sched(P, W);
  for (int t = 0; t < s.nops; ++t)
    for (int l = 0; l < W; ++l)
      if (s.r[t][l] == r && s.j[t][l] == j) return t * 2 + l;
  return -1;
}

// Row layout in units of R (shared by host builder and kernel):
//   [0,            NOPS*W)  left knots kl           (lane-major inside an op)
//   [NOPS*W,     2*NOPS*W)  y = RN(1/den)  (POW2 rows: the exact inverse)
//   [2*NOPS*W,   3*NOPS*W)  -den
//   [3*NOPS*W, 3*NOPS*W + (P+1)*DHP)  elements, each padded to DHP = roundup(DH, W) components
//   [hold_lo, hold_hi) — the parameter interval on which the span computation returns THIS row (-inf / +inf at
//   the clamped ends), so that a stepper can skip the search / span division while it stays inside
//   padded to a multiple of 16 bytes.
template <class R>
ENTERP_HD constexpr int lanes_of() { return sizeof(R) == 4 ? 2 : 1; }
ENTERP_HD constexpr int round_up(int a, int b) { return (a + b - 1) / b * b; }
template <class R>
ENTERP_HD constexpr int row_hold_off(int P, int DH) {
  const int W = lanes_of<R>();
  return 3 * sched_nops(P, W) * W + (P + 1) * round_up(DH, W);
}
template <class R>
ENTERP_HD constexpr int row_len(int P, int DH) {
  return round_up(row_hold_off<R>(P, DH) + 2, 16 / (int)sizeof(R));
}


}  // namespace enterp
