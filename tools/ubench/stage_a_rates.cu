// Micro-benchmarks behind the Stage A design decisions: instruction / memory-op rates on B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stage_a_rates stage_a_rates.cu && ./stage_a_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// (a) scalar rn mul + rn add (no contraction), 8 independent chains
__global__ void k_scalar(float* out, float m, float c, int iters) {
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-3f + j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = __fadd_rn(__fmul_rn(a[j], m), c);
  }
  float s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// (b) packed: exact mul = fma(a,b,-0), exact add = fma(a,1,c) with run-time constants (ptxas must not fuse)
__global__ void k_packed(u64* out, u64 m, u64 c, u64 one, u64 nz, int iters) {
  u64 a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0x3f8000003f800000ull + threadIdx.x + j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = fma2(fma2(a[j], m, nz), one, c);
  }
  u64 s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s ^= a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// (c) scalar FFMA (contracted) for reference
__global__ void k_ffma(float* out, float m, float c, int iters) {
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-3f + j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], m, c);
  }
  float s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// (d) shared-memory atomicMin u32: mode 0 consecutive lanes -> consecutive cells; 1 pseudo-random cells
__global__ void k_atoms(unsigned* out, int iters, int mode) {
  extern __shared__ unsigned z[];
  const int cells = 12288;
  for (int i = threadIdx.x; i < cells; i += blockDim.x) z[i] = ~0u;
  __syncthreads();
  unsigned x = threadIdx.x * 2654435761u + blockIdx.x;
  unsigned base = (threadIdx.x >> 5) * 97;
  for (int i = 0; i < iters; ++i) {
    unsigned cell;
    if (mode == 0) { cell = (base + (threadIdx.x & 31) + i * 41) % cells; }
    else { x = x * 1664525u + 1013904223u; cell = (x >> 8) % cells; }
    atomicMin(&z[cell], x ^ i);
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = z[5];
}
// (e) global RED.MIN.64 into an L2-resident buffer: mode 0 consecutive cells per warp, 1 random
__global__ void k_redg(u64* zb, unsigned ncells, int iters, int mode) {
  unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  for (int i = 0; i < iters; ++i) {
    unsigned cell;
    x = x * 1664525u + 1013904223u;
    if (mode == 0) cell = ((warp * 977u + i * 7919u) * 32u + lane) % ncells;
    else cell = x % ncells;
    atomicMin(zb + cell, ((u64)(x >> 4) << 32) | i);
  }
}
// (f) test-then-reduce: probe (ld.cg) and RED only if smaller (most probes fail once the buffer is low)
__global__ void k_probe(u64* zb, unsigned ncells, int iters, int mode, u64* sink) {
  unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  u64 acc = 0;
  for (int i = 0; i < iters; ++i) {
    unsigned cell;
    x = x * 1664525u + 1013904223u;
    if (mode == 0) cell = ((warp * 977u + i * 7919u) * 32u + lane) % ncells;
    else cell = x % ncells;
    acc += __ldcg(zb + cell);
  }
  if (acc == 12345) *sink = acc;
}
// (g) match_any + reduce
__global__ void k_match(unsigned* out, int iters) {
  unsigned x = threadIdx.x * 2654435761u;
  unsigned acc = 0;
  for (int i = 0; i < iters; ++i) {
    x = x * 1664525u + 1013904223u;
    unsigned cell = (threadIdx.x & 31) / 2 + (x >> 30);
    unsigned m = __match_any_sync(0xffffffffu, cell);
    acc += m;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_redux(unsigned* out, int iters) {
  unsigned x = threadIdx.x * 2654435761u;
  unsigned acc = 0;
  for (int i = 0; i < iters; ++i) {
    x = x * 1664525u + 1013904223u;
    acc += __reduce_min_sync(0xffffffffu, x);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F> float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("device %s, %d SMs, max clock %d MHz\n", p.name, sms, khz / 1000);
  void* buf; CK(cudaMalloc(&buf, 256u << 20)); CK(cudaMemset(buf, 0xFF, 256u << 20));
  const int threads = 256, ctas = sms * 8, iters = 4096;
  const double winst = (double)ctas * (threads / 32) * iters * 8;   // warp-level chain steps
  float ms;
  ms = time_ms([&] { k_scalar<<<ctas, threads>>>((float*)buf, 1.0001f, 0.5f, iters); });
  printf("scalar FMUL+FADD (rn, unfused): %.3f ms -> %.2f warp-inst/clk/SM at max clock (2 inst per step)\n", ms,
         2 * winst / (ms * 1e-3 * khz * 1e3) / sms);
  ms = time_ms([&] { k_ffma<<<ctas, threads>>>((float*)buf, 1.0001f, 0.5f, iters); });
  printf("scalar FFMA: %.3f ms -> %.2f warp-inst/clk/SM\n", ms, winst / (ms * 1e-3 * khz * 1e3) / sms);
  ms = time_ms([&] { k_packed<<<ctas, threads>>>((u64*)buf, 0x3f8003473f800347ull, 0x3f0000003f000000ull,
                                                 0x3f8000003f800000ull, 0x8000000080000000ull, iters); });
  printf("packed 2x FFMA2 per step (exact mul, exact add on 2 floats): %.3f ms -> %.2f warp-inst/clk/SM (2 inst per step)\n",
         ms, 2 * winst / (ms * 1e-3 * khz * 1e3) / sms);
  for (int mode = 0; mode < 2; ++mode) {
    cudaFuncSetAttribute(k_atoms, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
    const int it = 2048;
    ms = time_ms([&] { k_atoms<<<sms * 4, threads, 49152>>>((unsigned*)buf, it, mode); });
    printf("ATOMS.MIN.u32 %s: %.3f ms -> %.1f clk per warp-inst per SM\n", mode ? "random cells" : "consecutive cells", ms,
           ms * 1e-3 * khz * 1e3 / ((double)4 * (threads / 32) * it));
  }
  const unsigned ncells = 2u << 20;   // 16.8 MB of u64: L2 resident
  for (int mode = 0; mode < 2; ++mode) {
    const int it = 256;
    CK(cudaMemset(buf, 0xFF, (size_t)ncells * 8));
    ms = time_ms([&] { k_redg<<<ctas, threads>>>((u64*)buf, ncells, it, mode); });
    printf("REDG.MIN.u64 %s: %.3f ms -> %.1f clk per warp-inst per SM, %.2f G lane-ops/s\n", mode ? "random cells" : "consecutive cells",
           ms, ms * 1e-3 * khz * 1e3 / ((double)8 * (threads / 32) * it), (double)ctas * threads * it / (ms * 1e-3) / 1e9);
    ms = time_ms([&] { k_probe<<<ctas, threads>>>((u64*)buf, ncells, it, mode, (u64*)buf + ncells); });
    printf("LDG.CG.64 probe %s: %.3f ms -> %.1f clk per warp-inst per SM, %.2f G lane-ops/s\n", mode ? "random cells" : "consecutive cells",
           ms, ms * 1e-3 * khz * 1e3 / ((double)8 * (threads / 32) * it), (double)ctas * threads * it / (ms * 1e-3) / 1e9);
  }
  ms = time_ms([&] { k_match<<<ctas, threads>>>((unsigned*)buf, 1024); });
  printf("MATCH.ANY: %.3f ms -> %.1f clk per warp-inst per SM\n", ms, ms * 1e-3 * khz * 1e3 / ((double)8 * (threads / 32) * 1024));
  ms = time_ms([&] { k_redux<<<ctas, threads>>>((unsigned*)buf, 1024); });
  printf("REDUX.MIN: %.3f ms -> %.1f clk per warp-inst per SM\n", ms, ms * 1e-3 * khz * 1e3 / ((double)8 * (threads / 32) * 1024));
  return 0;
}
