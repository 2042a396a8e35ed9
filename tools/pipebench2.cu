// Issue-slot microbenchmark: does a packed f32x2 instruction (FMUL2, 2 cycles of FMA pipe) leave the
// second issue cycle free for an instruction of another pipe? Prints cycles per inner iteration
// per warp at full occupancy (8 blocks x 256 threads per SM).
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
constexpr int ILP = 8, ITERS = 2048;
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
template <int OP>
__global__ void __launch_bounds__(256) bench(float* out, float seed, unsigned kk) {
  float a[ILP], b[ILP]; u64 p[ILP]; unsigned q[ILP], q2[ILP];
  const float m = seed * 1.0000001f;
  const u64 pm = ((u64)__float_as_uint(m) << 32) | __float_as_uint(m);
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a[i] = seed + i + threadIdx.x; b[i] = a[i] * 0.5f; p[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(b[i]); q[i] = threadIdx.x + i; q2[i] = q[i] * 7; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) { p[i] = mul2(p[i], pm); }
      if (OP == 1) { p[i] = mul2(p[i], pm); asm volatile("add.u32 %0, %0, %1;" : "+r"(q[i]) : "r"(kk)); }
      if (OP == 2) { p[i] = mul2(p[i], pm); asm volatile("add.u32 %0, %0, %1;" : "+r"(q[i]) : "r"(kk)); asm volatile("xor.b32 %0, %0, %1;" : "+r"(q2[i]) : "r"(kk)); }
      if (OP == 3) { a[i] = __fmul_rn(a[i], m); b[i] = __fmul_rn(b[i], m); asm volatile("add.u32 %0, %0, %1;" : "+r"(q[i]) : "r"(kk)); }
      if (OP == 4) { a[i] = __fmul_rn(a[i], m); asm volatile("add.u32 %0, %0, %1;" : "+r"(q[i]) : "r"(kk)); }
      if (OP == 5) { asm volatile("add.u32 %0, %0, %1;" : "+r"(q[i]) : "r"(kk)); }
      if (OP == 6) { a[i] = __uint2float_rn(q[i]); q[i] += 3; }
      if (OP == 7) { q[i] = __float2uint_rz(a[i]); a[i] += 1.0f; }
      if (OP == 8) { a[i] = __ull2float_rn(((u64)q[i] << 20) + q2[i]); q[i] += 3; }
      if (OP == 9) { p[i] = mul2(p[i], pm); a[i] = __fmul_rn(a[i], m); }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i] + b[i] + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]) + (float)q[i] + (float)q2[i];
  if (s == 123.456f) out[0] = s;
}
template <int OP>
void run(const char* name, int sms, double ghz) {
  float* out; cudaMalloc(&out, 4);
  const int blocks = sms * 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<OP><<<blocks, 256>>>(out, 1.0f, 3); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); bench<OP><<<blocks, 256>>>(out, 1.0f, 3); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  // warps per SMSP = 8 blocks * 8 warps / 4 = 16; inner iterations per warp = ITERS*ILP
  const double cyc = best * 1e-3 * ghz * 1e9 / (16.0 * ITERS * ILP);
  printf("  \"%s\": {\"ms\": %.4f, \"cycles_per_inner_iter_per_smsp\": %.3f},\n", name, best, cyc);
  cudaFree(out);
}
int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount; const double ghz = 1.965;
  printf("{\n  \"assumed_ghz\": %.3f,\n", ghz);
  run<0>("fmul2", sms, ghz);
  run<1>("fmul2+iadd", sms, ghz);
  run<2>("fmul2+iadd+lop", sms, ghz);
  run<3>("fmul+fmul+iadd", sms, ghz);
  run<4>("fmul+iadd", sms, ghz);
  run<5>("iadd", sms, ghz);
  run<6>("u32_to_f32+iadd", sms, ghz);
  run<7>("f32_to_u32+fadd", sms, ghz);
  run<8>("u64_to_f32+misc", sms, ghz);
  run<9>("fmul2+fmul", sms, ghz);
  printf("  \"end\": 0\n}\n");
}
