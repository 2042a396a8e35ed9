// FP-pipe issue-rate microbenchmark for the roofline denominators that MEASURED_PEAKS.json lacks
// (SURVEY.md §8d footnote 1): dependency-free chains of unfused FMUL/FADD (scalar and packed
// f32x2), FFMA, DMUL/DADD/DFMA, MUFU.RCP and u64->f32 conversion. Prints one JSON line.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipebench tools/pipebench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

typedef unsigned long long u64;
constexpr int ILP = 8;
constexpr int ITERS = 2048;

__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int OP>
__global__ void __launch_bounds__(256) bench(float* out, float seed, long long* clk) {
  float a[ILP];
  double d[ILP];
  u64 p[ILP];
  u64 q[ILP];
  const float m = seed * 1.0000001f, c = seed * 1e-9f;
  const double dm = (double)m, dc = (double)c;
  const u64 pm = ((u64)__float_as_uint(m) << 32) | __float_as_uint(m);
  const u64 pc = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    a[i] = seed + i + threadIdx.x;
    d[i] = a[i];
    p[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f);
    q[i] = (u64)(threadIdx.x + i) * 0x9E3779B97F4A7C15ull;
  }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) a[i] = __fmul_rn(a[i], m);
      if (OP == 1) a[i] = __fadd_rn(a[i], c);
      if (OP == 2) a[i] = __fmaf_rn(a[i], m, c);
      if (OP == 3) { a[i] = __fmul_rn(a[i], m); a[i] = __fadd_rn(a[i], c); }   // 2 instr
      if (OP == 4) p[i] = mul2(p[i], pm);
      if (OP == 5) p[i] = add2(p[i], pc);
      if (OP == 6) p[i] = fma2(p[i], pm, pc);
      if (OP == 7) d[i] = __dmul_rn(d[i], dm);
      if (OP == 8) d[i] = __dadd_rn(d[i], dc);
      if (OP == 9) d[i] = __fma_rn(d[i], dm, dc);
      if (OP == 10) a[i] = __frcp_rn(a[i]) ;                 // full IEEE reciprocal sequence
      if (OP == 11) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));  // MUFU.RCP
      if (OP == 12) { a[i] = __ull2float_rn(q[i]); q[i] += __float_as_uint(a[i]); }  // I2F.U64 + IADD
      if (OP == 13) a[i] = __fdiv_rn(a[i], m);
      if (OP == 14) d[i] = __ddiv_rn(d[i], dm);
      if (OP == 15) { a[i] = __fmul_rn(a[i], m); q[i] = q[i] * 3 + 1; }       // FMUL + integer (alu/fma mix)
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i] + (float)d[i] + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]) + (float)q[i];
  if (s == 123.456f) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

template <int OP>
double run(const char* name, int ops_per_inner, int lanes_per_op, int sms) {
  float* out; long long* clk;
  cudaMalloc(&out, 4); cudaMalloc(&clk, 8);
  const int blocks = sms * 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<OP><<<blocks, 256>>>(out, 1.0f, clk);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    bench<OP><<<blocks, 256>>>(out, 1.0f, clk);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  long long cycles; cudaMemcpy(&cycles, clk, 8, cudaMemcpyDeviceToHost);
  const double thread_instr = (double)blocks * 256 * ITERS * ILP * ops_per_inner;
  const double rate = thread_instr / (best * 1e-3);  // thread-instructions / s, whole chip
  printf("  \"%s\": {\"ms\": %.4f, \"thread_instr_per_s\": %.4e, \"lane_ops_per_s\": %.4e, \"block0_cycles\": %lld},\n",
         name, best, rate, rate * lanes_per_op, cycles);
  cudaFree(out); cudaFree(clk);
  return rate;
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount;
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("{\n  \"device\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d,\n", prop.name, sms, khz);
  run<0>("fmul", 1, 1, sms);
  run<1>("fadd", 1, 1, sms);
  run<2>("ffma", 1, 1, sms);
  run<3>("fmul_fadd_pair", 2, 1, sms);
  run<4>("fmul2", 1, 2, sms);
  run<5>("fadd2", 1, 2, sms);
  run<6>("ffma2", 1, 2, sms);
  run<7>("dmul", 1, 1, sms);
  run<8>("dadd", 1, 1, sms);
  run<9>("dfma", 1, 1, sms);
  run<10>("frcp_rn", 1, 1, sms);
  run<11>("mufu_rcp", 1, 1, sms);
  run<12>("i2f_u64_plus_iadd", 1, 1, sms);
  run<13>("fdiv_rn", 1, 1, sms);
  run<14>("ddiv_rn", 1, 1, sms);
  run<15>("fmul_plus_imad", 2, 1, sms);
  printf("  \"iters\": %d, \"ilp\": %d\n}\n", ITERS, ILP);
  return 0;
}
