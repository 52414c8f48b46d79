// synth_cpu.hpp — CPU generator of the synthetic SAME corpus for the REFERENCE ARM of
// bench.py (`--impl reference`): the same signal model as the device generator (sameold_b200/csrc/same_synth.cu) —
// AWGN over the whole stream + continuous-phase AFSK bursts at 520.83 Bd, mark/space 2083.3/1562.5 Hz + per-stream
// offset, LSb first (waveform.rs:6-26) — so that the CPU arm decodes the same WORKLOAD without loading any GPU code.
// Not bit-identical to the device generator (libm instead of GPU fast-math); every decoder under test is always fed
// the identical int16 samples.  TEST / BENCH INFRASTRUCTURE ONLY.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace synth_cpu {

struct Burst {
  double start_sample;
  uint32_t byte_offset, n_bytes;
};

inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
inline float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// One stream: out[0..n_samples).  `cum_marks[i]` = number of mark bits before byte i of its burst.
inline void render_stream(int16_t* out, uint64_t n_samples, double rate, const Burst* bursts, uint32_t n_bursts,
                          const uint8_t* bytes, const uint16_t* cum_marks, float foff, uint32_t seed, float amplitude,
                          float sigma) {
  const double baud = 520.83, ts = rate / baud;
  const double fm = 2083.3 + (double)foff, fs = 1562.5 + (double)foff;
  // noise
  for (uint64_t g = 0; g * 4 < n_samples; ++g) {
    uint32_t r[4];
    philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 0u, 0u, seed, 0u, r);
    const float r0 = sqrtf(-2.0f * logf(u01(r[0]))), r1 = sqrtf(-2.0f * logf(u01(r[2])));
    const float a0 = 6.28318530718f * u01(r[1]), a1 = 6.28318530718f * u01(r[3]);
    const float z[4] = {r0 * cosf(a0), r0 * sinf(a0), r1 * cosf(a1), r1 * sinf(a1)};
    for (int j = 0; j < 4 && g * 4 + j < n_samples; ++j) {
      long q = lrintf(sigma * z[j]);
      out[g * 4 + j] = (int16_t)(q < -32768 ? -32768 : q > 32767 ? 32767 : q);
    }
  }
  // bursts (added on top of the rounded noise would double-round: re-derive the noise sample exactly instead)
  for (uint32_t b = 0; b < n_bursts; ++b) {
    const Burst& B = bursts[b];
    const double nsym = 8.0 * (double)B.n_bytes;
    uint64_t i0 = (uint64_t)std::ceil(B.start_sample < 0 ? 0.0 : B.start_sample);
    for (uint64_t n = i0; n < n_samples; ++n) {
      const double t = (double)n - B.start_sample;
      if (t >= nsym * ts) break;
      uint32_t k = (uint32_t)(t / ts);
      if (k >= 8u * B.n_bytes) k = 8u * B.n_bytes - 1u;
      const uint32_t byte_i = k >> 3, bit_i = k & 7u;
      const uint32_t byte = bytes[B.byte_offset + byte_i];
      const uint32_t km = cum_marks[B.byte_offset + byte_i] + (uint32_t)__builtin_popcount(byte & ((1u << bit_i) - 1u));
      const bool mark = (byte >> bit_i) & 1u;
      double cycles = (fm * (double)km + fs * (double)(k - km)) / baud + (mark ? fm : fs) * (t - (double)k * ts) / rate;
      cycles -= std::floor(cycles);
      const float sig = amplitude * cosf(6.28318530718f * (float)cycles);
      uint32_t r[4];
      const uint64_t g = n / 4;
      philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 0u, 0u, seed, 0u, r);
      const int j = (int)(n & 3);
      const float rr = sqrtf(-2.0f * logf(u01(r[j < 2 ? 0 : 2])));
      const float aa = 6.28318530718f * u01(r[j < 2 ? 1 : 3]);
      const float z = (j & 1) ? rr * sinf(aa) : rr * cosf(aa);
      long q = lrintf(sig + sigma * z);
      out[n] = (int16_t)(q < -32768 ? -32768 : q > 32767 ? 32767 : q);
    }
  }
}

}  // namespace synth_cpu
