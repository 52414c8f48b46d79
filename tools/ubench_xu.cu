// ubench_xu.cu — throughput of the conversion / special-function ("XU") instructions the receiver kernel touches:
// I2F, F2I, F2F(f32<->f64), MUFU.RSQ, DP sqrt, plus FFMA2 and FMNMX dependent-chain latency.  B200 (sm_100a).
#include <cuda_runtime.h>
#include <cstdio>
#define ITERS 2048

__global__ void k_i2f(float* out, int a) {
  int v[8]; float acc = 0;
  for (int j = 0; j < 8; ++j) v[j] = a + threadIdx.x + j;
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc += (float)v[j]; v[j] += i; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_magic(float* out, int a) {   // same work with the magic-number conversion (no XU)
  int v[8]; float acc = 0;
  for (int j = 0; j < 8; ++j) v[j] = (a + threadIdx.x + j) & 0xffff;
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc += __int_as_float(0x4B400000 + v[j]) - 12582912.0f; v[j] = (v[j] + i) & 0xffff; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_f2f(float* out, float a) {
  float v[4]; for (int j = 0; j < 4; ++j) v[j] = a + threadIdx.x * 1e-3f + j;
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { double d = (double)v[j]; d = d * 1.0000001; v[j] = (float)d; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = v[0] + v[1] + v[2] + v[3];
}
__global__ void k_dsqrt(float* out, float a) {
  double v[2]; for (int j = 0; j < 2; ++j) v[j] = a + threadIdx.x * 1e-3 + j;
  for (int i = 0; i < ITERS / 4; ++i) {
#pragma unroll
    for (int j = 0; j < 2; ++j) v[j] = __dsqrt_rn(v[j] + 1.5);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(v[0] + v[1]);
}
__global__ void k_ffma2_chain(float* out, float a, float one, float nz) {  // dependent FFMA2 chain latency
  float2 acc = make_float2(a, a + 1), o2 = make_float2(one, one), x = make_float2(threadIdx.x * 1e-6f, nz);
  for (int i = 0; i < ITERS * 4; ++i) acc = __ffma2_rn(acc, o2, x);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y;
}
__global__ void k_fadd_chain(float* out, float a, float x) {
  float acc = a + threadIdx.x;
  for (int i = 0; i < ITERS * 4; ++i) acc = __fadd_rn(acc, x);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_fmnmx_chain(float* out, float a, float lo, float hi) {
  float acc = a + threadIdx.x * 1e-3f;
  for (int i = 0; i < ITERS * 4; ++i) acc = fminf(fmaxf(__fadd_rn(acc, 1e-3f), lo), hi);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 64 * 32 * sizeof(float) * 2);
  const double clk = 1.965e9;
  for (int w : {1, 4, 8, 16, 32}) {
    int threads = 32 * (w < 4 ? w : 4), blocks = 148 * (w / (threads / 32));
    double warps = (double)blocks * threads / 32;
    float t;
    printf("warps/SM %2d:", w);
    t = timeit([&] { k_i2f<<<blocks, threads>>>(out, 3); });
    printf(" I2F %.2f cyc/warp-instr/SM", t * 1e-3 * clk / (warps / 148 * ITERS * 8));
    t = timeit([&] { k_magic<<<blocks, threads>>>(out, 3); });
    printf(" | magic-i2f %.2f", t * 1e-3 * clk / (warps / 148 * ITERS * 8));
    t = timeit([&] { k_f2f<<<blocks, threads>>>(out, 1.f); });
    printf(" | F2F pair+DMUL %.2f", t * 1e-3 * clk / (warps / 148 * ITERS * 4));
    t = timeit([&] { k_dsqrt<<<blocks, threads>>>(out, 1.f); });
    printf(" | dsqrt %.1f", t * 1e-3 * clk / (warps / 148 * (ITERS / 4) * 2));
    t = timeit([&] { k_ffma2_chain<<<blocks, threads>>>(out, 1.f, 1.f, -0.f); });
    printf(" | FFMA2 chain %.2f cyc/op/warp", t * 1e-3 * clk / (ITERS * 4.0));
    t = timeit([&] { k_fadd_chain<<<blocks, threads>>>(out, 1.f, 1e-3f); });
    printf(" | FADD chain %.2f", t * 1e-3 * clk / (ITERS * 4.0));
    t = timeit([&] { k_fmnmx_chain<<<blocks, threads>>>(out, 1.f, 0.f, 1e9f); });
    printf(" | FADD+2FMNMX chain %.2f\n", t * 1e-3 * clk / (ITERS * 4.0));
  }
  return 0;
}
