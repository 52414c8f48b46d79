// ubench_pipes.cu — measures the issue rates that bound the receiver kernel on B200 (sm_100a):
// scalar FMUL+FADD (unfusable), packed FFMA2 used as exact mul / exact add, FMNMX, DP hypot, LDS.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench_pipes tools/ubench_pipes.cu
#include <cuda_runtime.h>
#include <cstdio>

#define ITERS 4096
struct RT { float one, negz; };

__global__ void k_scalar(float* out, float a, float b) {   // 8 independent chains: FMUL + FADD per step
  float acc[8]; for (int j = 0; j < 8; ++j) acc[j] = threadIdx.x * 1e-3f + j;
  float x = a + threadIdx.x * 1e-6f;
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = __fadd_rn(acc[j], __fmul_rn(x, b + j));
  }
  float s = 0; for (int j = 0; j < 8; ++j) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_packed(float* out, float a, float b, const __grid_constant__ RT rt) {  // same math, 4 packed chains
  float2 acc[4]; for (int j = 0; j < 4; ++j) acc[j] = make_float2(threadIdx.x * 1e-3f + 2 * j, threadIdx.x * 1e-3f + 2 * j + 1);
  float x = a + threadIdx.x * 1e-6f;
  const float2 one = make_float2(rt.one, rt.one), negz = make_float2(rt.negz, rt.negz), xx = make_float2(x, x);
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 p = __ffma2_rn(xx, make_float2(b + 2 * j, b + 2 * j + 1), negz);
      acc[j] = __ffma2_rn(acc[j], one, p);
    }
  }
  float s = 0; for (int j = 0; j < 4; ++j) s += acc[j].x + acc[j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma(float* out, float a, float b) {   // plain FFMA reference (fused) 8 chains
  float acc[8]; for (int j = 0; j < 8; ++j) acc[j] = threadIdx.x * 1e-3f + j;
  float x = a + threadIdx.x * 1e-6f;
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = __fmaf_rn(x, b + j, acc[j]);
  }
  float s = 0; for (int j = 0; j < 8; ++j) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_agc(float* out, float d0, float bw, float gmin, float gmax) {  // the AGC recurrence: dependent chain
  float g = 1e-4f + threadIdx.x * 1e-9f, d = d0;
  float s = 0;
  for (int i = 0; i < ITERS * 2; ++i) {
    float y = __fmul_rn(d, g);
    g = __fadd_rn(g, __fmul_rn(__fsub_rn(1.0f, fabsf(y)), bw));
    g = fminf(fmaxf(g, gmin), gmax);
    s += y; d = -d;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + g;
}

__global__ void k_hypot(float* out, float a) {
  float s = 0, x = a + threadIdx.x * 1e-3f, y = 0.5f;
  for (int i = 0; i < ITERS / 8; ++i) {
    double p = (double)x, q = (double)y;
    float h = __double2float_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(p, p), __dmul_rn(q, q))));
    s += h; x = h * 0.999f; y += 1e-3f;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 16 * 1024 * 4 * sizeof(float));
  int dev_clock; cudaDeviceGetAttribute(&dev_clock, cudaDevAttrClockRate, 0);
  RT rt{1.0f, -0.0f};
  for (int warps_per_sm : {4, 8, 16, 32, 64}) {
    int threads = 256, blocks = 148 * warps_per_sm * 32 / threads;
    double lane_steps = (double)blocks * threads * ITERS * 8;  // mul+add pairs
    float t1 = timeit([&] { k_scalar<<<blocks, threads>>>(out, 1.0f, 0.5f); });
    float t2 = timeit([&] { k_packed<<<blocks, threads>>>(out, 1.0f, 0.5f, rt); });
    float t3 = timeit([&] { k_ffma<<<blocks, threads>>>(out, 1.0f, 0.5f); });
    float t4 = timeit([&] { k_agc<<<blocks, threads>>>(out, 100.0f, 1.9e-5f, 3e-5f, 5e-3f); });
    float t5 = timeit([&] { k_hypot<<<blocks, threads>>>(out, 1.0f); });
    printf("warps/SM %2d: scalar mul+add %.3f ms (%.1f Gpair/s) | packed FFMA2x2 %.3f ms (%.1f Gpair/s) | FFMA %.3f ms (%.1f G/s) | agc %.3f ms (%.2f ns/sample/warp-chain, %.1f Gsamples/s) | hypot %.3f ms (%.2f G/s)\n",
           warps_per_sm, t1, lane_steps / t1 * 1e-6, t2, lane_steps / t2 * 1e-6, t3, lane_steps / t3 * 1e-6,
           t4, t4 * 1e6 / (ITERS * 2), (double)blocks * threads * ITERS * 2 / t4 * 1e-6,
           t5, (double)blocks * threads * (ITERS / 8) / t5 * 1e-6);
  }
  printf("clock attr %d kHz\n", dev_clock);
  return 0;
}
