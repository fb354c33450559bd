// FADD vs FADD2 on sm_100a: dependent-chain latency and issue throughput per SM sub-partition
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 u; asm("mov.b64 %0, {%1, %2};" : "=l"(u) : "f"(a), "f"(b)); return u; }
template <int MODE, int ILP>
__global__ void k(float *out, int iters, float seed, long long *cyc) {
  float a[ILP], b[ILP]; u64 p[ILP];
  for (int i = 0; i < ILP; i++) { a[i] = seed + i + threadIdx.x; b[i] = seed * 2 + i; p[i] = pk(a[i], b[i]); }
  u64 inc = pk(seed, seed);
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (MODE == 0) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed)); }
      if (MODE == 1) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(inc)); }
      if (MODE == 2) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(seed)); }
      if (MODE == 3) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(inc)); }
      if (MODE == 4) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(inc)); asm volatile("add.s32 %0, %0, 1;" : "+r"(*(int*)&a[i])); }
      if (MODE == 5) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(seed)); asm volatile("add.s32 %0, %0, 1;" : "+r"(*(int*)&p[i])); }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < ILP; i++) s += a[i] + b[i] + (float)p[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE, int ILP> void run(const char *name, int warps, int flops_per_op) {
  float *o; long long *c, hc; cudaMalloc(&o, 4 << 20); cudaMalloc(&c, 8);
  int iters = 4096;
  k<MODE, ILP><<<1, warps * 32>>>(o, iters, 1.0f, c); cudaDeviceSynchronize();
  k<MODE, ILP><<<1, warps * 32>>>(o, iters, 1.0f, c); cudaDeviceSynchronize();
  cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
  double per = (double)hc / iters / ILP;
  printf("%-34s warps/SM %2d ILP %d: %.2f cycles per op-group per warp; %.1f fp32 lane-results/clk/SM\n", name, warps, ILP, per,
         32.0 * flops_per_op * warps / per);
  cudaFree(o); cudaFree(c);
}
int main() {
  puts("# dependent chain (ILP 1, 1 warp): latency");
  run<0, 1>("FADD", 1, 1); run<1, 1>("FADD2", 1, 2); run<3, 1>("FMUL2", 1, 2);
  puts("# throughput: 4 warps (1 per scheduler), ILP 8");
  run<0, 8>("FADD", 4, 1); run<1, 8>("FADD2", 4, 2); run<2, 8>("2 x FADD", 4, 2);
  puts("# throughput: 16 warps, ILP 8");
  run<0, 8>("FADD", 16, 1); run<1, 8>("FADD2", 16, 2);
  puts("# mixed with integer work: {FADD2 + IADD} vs {2 FADD + IADD}, 16 warps");
  run<4, 8>("FADD2 + IADD", 16, 2); run<5, 8>("2 x FADD + IADD", 16, 2);
  return 0;
}
