// same_synth.cu — synthetic SAME corpus generator (see include/same_synth.h).  Time-parallel: one thread writes
// 8 consecutive samples (one 16-byte store), so the kernel is a plain HBM-write stream.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>

#include "../../include/same_synth.h"

namespace {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0, hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

__global__ void synth_kernel(int16_t* __restrict__ out, uint32_t n_streams, unsigned long long stride,
                             uint32_t n_samples, unsigned long long first_group, double rate,
                             const uint32_t* __restrict__ burst_begin,
                             const same_synth_burst* __restrict__ bursts, const uint8_t* __restrict__ bytes,
                             const uint16_t* __restrict__ cum_marks, const float* __restrict__ foff,
                             const uint32_t* __restrict__ seeds, float amplitude, float sigma) {
  const uint32_t stream = blockIdx.y;
  const uint32_t gl = blockIdx.x * blockDim.x + threadIdx.x;  // group of 8 samples inside this window
  const unsigned long long l0 = (unsigned long long)gl * 8ull;  // first sample of the group, relative to the window
  if (stream >= n_streams || l0 >= n_samples) return;
  const unsigned long long g = first_group + gl;               // group index in the whole stream: noise counter
  const unsigned long long n0 = g * 8ull;                      // absolute sample index
  const uint32_t seed = seeds[stream];
  float z[8];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t r[4];
    const unsigned long long ctr = 2ull * g + (unsigned long long)h;
    philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u, seed, 0u, r);
    float r0 = sqrtf(-2.0f * __logf(u01(r[0]))), r1 = sqrtf(-2.0f * __logf(u01(r[2])));
    float s0, c0, s1, c1;
    __sincosf(6.28318530718f * u01(r[1]), &s0, &c0);
    __sincosf(6.28318530718f * u01(r[3]), &s1, &c1);
    z[4 * h + 0] = r0 * c0; z[4 * h + 1] = r0 * s0; z[4 * h + 2] = r1 * c1; z[4 * h + 3] = r1 * s1;
  }
  const double baud = 520.83, ts = rate / baud;  // samples per symbol (fractional)
  const double fm = 2083.3 + (double)foff[stream], fs = 1562.5 + (double)foff[stream];
  const uint32_t b0 = burst_begin[stream], b1 = burst_begin[stream + 1];
  short v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double n = (double)(n0 + j);
    float sig = 0.0f;
    for (uint32_t b = b0; b < b1; ++b) {
      const same_synth_burst B = bursts[b];
      const double t = n - B.start_sample;  // samples into the burst
      const double nsym = 8.0 * (double)B.n_bytes;
      if (t >= 0.0 && t < nsym * ts) {
        uint32_t k = (uint32_t)(t / ts);
        if (k >= 8u * B.n_bytes) k = 8u * B.n_bytes - 1u;
        const uint32_t byte_i = k >> 3, bit_i = k & 7u;
        const uint32_t byte = bytes[B.byte_offset + byte_i];
        const uint32_t km = cum_marks[B.byte_offset + byte_i] + __popc(byte & ((1u << bit_i) - 1u));  // marks before symbol k
        const bool mark = (byte >> bit_i) & 1u;                                                       // LSb first
        double cycles = (fm * (double)km + fs * (double)(k - km)) / baud + (mark ? fm : fs) * (t - (double)k * ts) / rate;
        cycles -= floor(cycles);
        sig = amplitude * cospif(2.0f * (float)cycles);
        break;
      }
    }
    float x = sig + sigma * z[j];
    int q = __float2int_rn(x);
    q = max(-32768, min(32767, q));
    v[j] = (short)q;
  }
  int16_t* dst = out + (unsigned long long)stream * stride + l0;
  if (l0 + 8 <= n_samples && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
    int4 pk;
    pk.x = (int)((uint32_t)(uint16_t)v[0] | ((uint32_t)(uint16_t)v[1] << 16));
    pk.y = (int)((uint32_t)(uint16_t)v[2] | ((uint32_t)(uint16_t)v[3] << 16));
    pk.z = (int)((uint32_t)(uint16_t)v[4] | ((uint32_t)(uint16_t)v[5] << 16));
    pk.w = (int)((uint32_t)(uint16_t)v[6] | ((uint32_t)(uint16_t)v[7] << 16));
    *reinterpret_cast<int4*>(dst) = pk;
  } else {
    for (int j = 0; j < 8 && l0 + j < n_samples; ++j) dst[j] = v[j];
  }
}

}  // namespace

extern "C" int same_synth_generate(int device, int16_t* d_out, uint32_t n_streams, uint64_t stride, uint64_t first_sample,
                                   uint32_t n_samples, uint32_t rate, const uint32_t* burst_begin, const same_synth_burst* bursts,
                                   uint32_t n_bursts, const uint8_t* bytes, uint64_t n_bytes_total,
                                   const float* freq_offset_hz, const uint32_t* seeds, float amplitude,
                                   float noise_sigma, char* err_text) {
  cudaError_t err = cudaSuccess;
  uint32_t* d_bb = nullptr; same_synth_burst* d_b = nullptr; uint8_t* d_bytes = nullptr; uint16_t* d_cum = nullptr;
  float* d_f = nullptr; uint32_t* d_s = nullptr;
  uint16_t* cum = nullptr;
#define SCK(call) do { err = (call); if (err != cudaSuccess) goto done; } while (0)
  if (first_sample % 8u) {
    if (err_text) snprintf(err_text, 256, "first_sample must be a multiple of 8");
    return (int)cudaErrorInvalidValue;
  }
  SCK(cudaSetDevice(device));
  // marks before each byte, per burst
  cum = new uint16_t[n_bytes_total ? n_bytes_total : 1];
  for (uint32_t b = 0; b < n_bursts; ++b) {
    uint32_t acc = 0;
    for (uint32_t i = 0; i < bursts[b].n_bytes; ++i) {
      cum[bursts[b].byte_offset + i] = (uint16_t)acc;
      acc += (uint32_t)__builtin_popcount(bytes[bursts[b].byte_offset + i]);
    }
  }
  SCK(cudaMalloc(&d_bb, (size_t)(n_streams + 1) * 4));
  SCK(cudaMalloc(&d_b, (size_t)(n_bursts ? n_bursts : 1) * sizeof(same_synth_burst)));
  SCK(cudaMalloc(&d_bytes, n_bytes_total ? n_bytes_total : 1));
  SCK(cudaMalloc(&d_cum, (n_bytes_total ? n_bytes_total : 1) * 2));
  SCK(cudaMalloc(&d_f, (size_t)n_streams * 4));
  SCK(cudaMalloc(&d_s, (size_t)n_streams * 4));
  SCK(cudaMemcpy(d_bb, burst_begin, (size_t)(n_streams + 1) * 4, cudaMemcpyHostToDevice));
  if (n_bursts) SCK(cudaMemcpy(d_b, bursts, (size_t)n_bursts * sizeof(same_synth_burst), cudaMemcpyHostToDevice));
  if (n_bytes_total) {
    SCK(cudaMemcpy(d_bytes, bytes, n_bytes_total, cudaMemcpyHostToDevice));
    SCK(cudaMemcpy(d_cum, cum, n_bytes_total * 2, cudaMemcpyHostToDevice));
  }
  SCK(cudaMemcpy(d_f, freq_offset_hz, (size_t)n_streams * 4, cudaMemcpyHostToDevice));
  SCK(cudaMemcpy(d_s, seeds, (size_t)n_streams * 4, cudaMemcpyHostToDevice));
  {
    const uint32_t groups = (n_samples + 7u) / 8u;
    // grid.y is limited to 65535: generate in slabs of streams
    for (uint32_t s0 = 0; s0 < n_streams; s0 += 32768u) {
      const uint32_t ns = (n_streams - s0 < 32768u) ? (n_streams - s0) : 32768u;
      dim3 grid((groups + 255u) / 256u, ns);
      synth_kernel<<<grid, 256>>>(d_out + (unsigned long long)s0 * stride, ns, stride, n_samples, first_sample / 8u, (double)rate,
                                  d_bb + s0, d_b, d_bytes, d_cum, d_f + s0, d_s + s0, amplitude, noise_sigma);
      SCK(cudaGetLastError());
    }
  }
  SCK(cudaDeviceSynchronize());
done:
#undef SCK
  delete[] cum;
  cudaFree(d_bb); cudaFree(d_b); cudaFree(d_bytes); cudaFree(d_cum); cudaFree(d_f); cudaFree(d_s);
  if (err != cudaSuccess && err_text) snprintf(err_text, 256, "%s", cudaGetErrorString(err));
  return (int)err;
}
