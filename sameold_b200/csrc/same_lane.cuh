// same_lane.cuh — per-lane (= per-stream) receiver state and the symbol-rate half of the receiver, shared by the
// generic and the fast kernels.  Everything here runs at TED rate (~1042/s per stream) or below; the sample-rate half
// (DC blocker, AGC, demod window) and the matched filter live in same_kernels.cu.
//
// Reference: crates/sameold/src/receiver.rs:343-490 and receiver/{symsync,codesquelch,equalize,framing}.rs.
// Every f32 operation is a single IEEE round-to-nearest operation in the reference's order (explicit _rn intrinsics,
// never contracted), no FTZ.
#pragma once

#include <cuda_runtime.h>

#include "same_params.h"
#include "same_transport.cuh"

namespace same_dev {

#define FMUL(a, b) __fmul_rn((a), (b))
#define FADD(a, b) __fadd_rn((a), (b))
#define FSUB(a, b) __fsub_rn((a), (b))

// Rust f32::clamp: comparisons only, NaN passes through
__device__ __forceinline__ float rclamp(float x, float lo, float hi) {
  if (x < lo) x = lo;
  if (x > hi) x = hi;
  return x;
}
// Rust f32::signum for non-NaN input: +0 -> +1, -0 -> -1  (symsync.rs:320-322, equalize.rs:264-268)
__device__ __forceinline__ float rsignum(float x) { return copysignf(1.0f, x); }

// Complex<f32>::norm() = hypotf(re, im) (demod.rs:163); fixed definition shared with the oracle:
// (float)sqrt((double)re*re + (double)im*im) == glibc hypotf
__device__ __forceinline__ float hypot_fixed(float re, float im) {
  double a = (double)re, b = (double)im;
  return __double2float_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b))));
}

// receiver.rs:352-353 fires at the first clock value c with  r = until - (c as f32);  r <= 0 || |r| < 0.5.
// `cf` is the clock value as an exactly representable float (small integer), so FSUB(until, cf) is the reference's
// subtraction bit for bit.
__device__ __forceinline__ bool fires(float until, float cf) {
  const float r = FSUB(until, cf);
  return r <= 0.0f || fabsf(r) < 0.5f;
}
// Smallest c > clock_now that fires.  The predicate is monotone in c.
// Closed form for 1 <= until < 2^22 (every realistic rate: until is about half a symbol, 21 samples at 22050 Hz):
//   c* = floor(until - 0.5) + 1.
// Proof that this is the reference's first firing clock: until - 0.5 is exact (0.5 is a multiple of ulp(until) <= 0.25
// and the result lies in a binade at or below until's); for an integer 1 <= c <= 2^22 the real difference
// d = until - c is a multiple of ulp(until) with |d| <= until, hence representable, so FSUB(until, c) == d exactly.
// For c <= floor(until - 0.5): d >= 0.5, no fire.  For c = c*: -0.5 <= d < 0.5, i.e. d <= 0 or |d| < 0.5: fires.
// Outside that range (degenerate configurations) the candidate is verified against the predicate and walked.
// Integer<->float conversions use the 2^23 magic-number form (FMA/ALU pipes) instead of the conversion unit.
__device__ __forceinline__ int fire_clock(float until, int clock_now) {
  float t = fminf(fmaxf(until - 0.5f, 0.0f), 4.0e6f);
  float cf = FSUB(__fadd_rd(t, 8388608.0f), 8388608.0f) + 1.0f;   // floor(t) + 1 for 0 <= t < 2^23
  if (!(until >= 1.0f && until < 4.0e6f)) {
    if (!fires(until, cf)) {
      int guard = 0;
      do { cf += 1.0f; } while (!fires(until, cf) && ++guard < 64);
    } else if (cf > 1.0f && fires(until, cf - 1.0f)) {
      int guard = 0;
      do { cf -= 1.0f; } while (cf > 1.0f && fires(until, cf - 1.0f) && ++guard < 64);
    }
  }
  int c = __float_as_int(cf + 12582912.0f) - 0x4B400000;           // exact for integers |c| < 2^22
  if (c <= clock_now) c = clock_now + 1;
  return c;
}

// ----------------------------------------------------------------------------------------------------------------
// Lane state (registers)
// ----------------------------------------------------------------------------------------------------------------
struct Lane {
  // A2 AGC
  float g;
  uint32_t flags;
  // A5 sample clock + timing loop
  int clock;
  float until, pavg, pinst, ted0, ted1, ted2;
  uint32_t tedcnt;
  // A6 squelch
  uint32_t sq_data, sq_pflags, sq_head;
  float sq_power;
  int byteclk;
  unsigned long long symcount;
  // A7 equalizer mode
  uint32_t train_sa, train_cnt;
  // A8 framer, A9 link/transport
  Framer fr;
  uint32_t link_last;
  Transport tr;
  EvCtx ev;
  uint32_t trace_n;
  unsigned long long n0;   // input_sample_counter at the start of this chunk
};

#define LANE_ST(st, L, w) (st)[(size_t)(w) * (L).n_pad]

__device__ __forceinline__ void lane_load(Lane& a, const SameParams& p, const uint32_t* st, uint32_t s) {
  const SameLayout& L = p.layout;
  a.g = __uint_as_float(LANE_ST(st, L, F_AGC_GAIN));
  a.flags = LANE_ST(st, L, F_FLAGS);
  a.clock = (int)LANE_ST(st, L, F_CLOCK);
  a.until = __uint_as_float(LANE_ST(st, L, F_UNTIL));
  a.pavg = __uint_as_float(LANE_ST(st, L, F_PAVG));
  a.pinst = __uint_as_float(LANE_ST(st, L, F_PINST));
  a.ted0 = __uint_as_float(LANE_ST(st, L, F_TED0));
  a.ted1 = __uint_as_float(LANE_ST(st, L, F_TED1));
  a.ted2 = __uint_as_float(LANE_ST(st, L, F_TED2));
  a.tedcnt = LANE_ST(st, L, F_TEDCNT);
  a.sq_data = LANE_ST(st, L, F_SQ_DATA);
  a.sq_power = __uint_as_float(LANE_ST(st, L, F_SQ_POWER));
  a.sq_pflags = LANE_ST(st, L, F_SQ_PFLAGS);
  a.sq_head = LANE_ST(st, L, F_SQ_HEAD);
  a.byteclk = (int)LANE_ST(st, L, F_SQ_BYTECLK);
  a.symcount = ((unsigned long long)LANE_ST(st, L, F_SYMCOUNT_HI) << 32) | LANE_ST(st, L, F_SYMCOUNT_LO);
  a.n0 = ((unsigned long long)LANE_ST(st, L, F_N_HI) << 32) | LANE_ST(st, L, F_N_LO);
  a.train_sa = LANE_ST(st, L, F_EQ_TRAIN_SA);
  a.train_cnt = LANE_ST(st, L, F_EQ_TRAIN_CNT);
  a.fr.st = (a.flags >> FLAG_FR_SHIFT) & 3u;
  a.fr.word = LANE_ST(st, L, F_FR_WORD);
  a.fr.count = LANE_ST(st, L, F_FR_COUNT);
  a.fr.invalid = LANE_ST(st, L, F_FR_INVALID);
  a.fr.msglen = LANE_ST(st, L, F_FR_MSGLEN);
  a.link_last = (a.flags >> FLAG_LINK_SHIFT) & 3u;
  a.tr.hist_n = LANE_ST(st, L, F_HIST_N);
  a.tr.pending = a.flags & FLAG_PENDING;
  a.tr.have_prev = a.flags & FLAG_HAVE_PREV;
  a.tr.have_eom = a.flags & FLAG_FORCE_EOM;
  a.tr.tr_state = (a.flags >> FLAG_TR_SHIFT) & 3u;
  a.tr.next_deadline = ((unsigned long long)LANE_ST(st, L, F_TRNEXT_HI) << 32) | LANE_ST(st, L, F_TRNEXT_LO);
  a.tr.eom_at = ((unsigned long long)LANE_ST(st, L, F_EOM_HI) << 32) | LANE_ST(st, L, F_EOM_LO);
  a.ev.stream = s; a.ev.seq = LANE_ST(st, L, F_SEQ);
  a.trace_n = LANE_ST(st, L, F_TRACE_N);
}

__device__ __forceinline__ void lane_store(const Lane& a, const SameParams& p, uint32_t* st, unsigned long long n1) {
  const SameLayout& L = p.layout;
  uint32_t flags = a.flags & ~((3u << FLAG_FR_SHIFT) | (3u << FLAG_LINK_SHIFT) | (3u << FLAG_TR_SHIFT) | FLAG_PENDING |
                               FLAG_HAVE_PREV | FLAG_FORCE_EOM);
  flags |= (a.fr.st << FLAG_FR_SHIFT) | (a.link_last << FLAG_LINK_SHIFT) | (a.tr.tr_state << FLAG_TR_SHIFT);
  if (a.tr.pending) flags |= FLAG_PENDING;
  if (a.tr.have_prev) flags |= FLAG_HAVE_PREV;
  if (a.tr.have_eom) flags |= FLAG_FORCE_EOM;
  LANE_ST(st, L, F_AGC_GAIN) = __float_as_uint(a.g);
  LANE_ST(st, L, F_FLAGS) = flags;
  LANE_ST(st, L, F_CLOCK) = (uint32_t)a.clock;
  LANE_ST(st, L, F_UNTIL) = __float_as_uint(a.until);
  LANE_ST(st, L, F_PAVG) = __float_as_uint(a.pavg);
  LANE_ST(st, L, F_PINST) = __float_as_uint(a.pinst);
  LANE_ST(st, L, F_TED0) = __float_as_uint(a.ted0);
  LANE_ST(st, L, F_TED1) = __float_as_uint(a.ted1);
  LANE_ST(st, L, F_TED2) = __float_as_uint(a.ted2);
  LANE_ST(st, L, F_TEDCNT) = a.tedcnt;
  LANE_ST(st, L, F_SQ_DATA) = a.sq_data;
  LANE_ST(st, L, F_SQ_POWER) = __float_as_uint(a.sq_power);
  LANE_ST(st, L, F_SQ_PFLAGS) = a.sq_pflags;
  LANE_ST(st, L, F_SQ_HEAD) = a.sq_head;
  LANE_ST(st, L, F_SQ_BYTECLK) = (uint32_t)a.byteclk;
  LANE_ST(st, L, F_SYMCOUNT_LO) = (uint32_t)a.symcount;
  LANE_ST(st, L, F_SYMCOUNT_HI) = (uint32_t)(a.symcount >> 32);
  LANE_ST(st, L, F_N_LO) = (uint32_t)n1;
  LANE_ST(st, L, F_N_HI) = (uint32_t)(n1 >> 32);
  LANE_ST(st, L, F_EQ_TRAIN_SA) = a.train_sa;
  LANE_ST(st, L, F_EQ_TRAIN_CNT) = a.train_cnt;
  LANE_ST(st, L, F_FR_WORD) = a.fr.word;
  LANE_ST(st, L, F_FR_COUNT) = a.fr.count;
  LANE_ST(st, L, F_FR_INVALID) = a.fr.invalid;
  LANE_ST(st, L, F_FR_MSGLEN) = a.fr.msglen;
  LANE_ST(st, L, F_EOM_LO) = (uint32_t)a.tr.eom_at;
  LANE_ST(st, L, F_EOM_HI) = (uint32_t)(a.tr.eom_at >> 32);
  LANE_ST(st, L, F_TRNEXT_LO) = (uint32_t)a.tr.next_deadline;
  LANE_ST(st, L, F_TRNEXT_HI) = (uint32_t)(a.tr.next_deadline >> 32);
  LANE_ST(st, L, F_HIST_N) = a.tr.hist_n;
  LANE_ST(st, L, F_SEQ) = a.ev.seq;
  LANE_ST(st, L, F_TRACE_N) = a.trace_n;
}

// ----------------------------------------------------------------------------------------------------------------
// Equalizer (equalize.rs).  State is read from / written back to the stream's state words around each byte.
// ----------------------------------------------------------------------------------------------------------------
template <int NFF, int NFB>
struct EqRegs {
  float ffc[NFF], fbc[NFB], ffw[NFF], fbw[NFB];  // windows: index 0 oldest
};

template <int NFF, int NFB>
__device__ __forceinline__ void eq_symbol(const SameParams& p, EqRegs<NFF, NFB>& q, int nff, int nfb, float z, float s,
                                          uint32_t& flags, uint32_t& train_sa, uint32_t& train_cnt, bool& bit) {
  // feedforward_wind.push(&[z, s])  equalize.rs:253 (== two push_scalar, see filter.rs:257-273)
#pragma unroll
  for (int i = 0; i < NFF - 1; ++i) if (i < nff - 1) q.ffw[i] = q.ffw[i + 1];
  q.ffw[nff - 1] = z;
#pragma unroll
  for (int i = 0; i < NFF - 1; ++i) if (i < nff - 1) q.ffw[i] = q.ffw[i + 1];
  q.ffw[nff - 1] = s;
  // filters: newest sample pairs with coeff[0]  filter.rs:363-377
  float ff = 0.0f, fb = 0.0f;
#pragma unroll
  for (int i = 0; i < NFF; ++i) if (i < nff) ff = FADD(ff, FMUL(q.ffw[nff - 1 - i], q.ffc[i]));
#pragma unroll
  for (int i = 0; i < NFB; ++i) if (i < nfb) fb = FADD(fb, FMUL(q.fbw[nfb - 1 - i], q.fbc[i]));
  float sym_val = FSUB(ff, fb);
  float sym_est;
  if (flags & FLAG_EQ_TRAINING) {  // equalize.rs:277-300
    sym_est = FSUB(FMUL(2.0f, (float)(train_sa & 1u)), 1.0f);
    train_sa >>= 1;
    train_cnt += 1;
    if (train_cnt >= 32u) flags &= ~FLAG_EQ_TRAINING;
  } else {
    sym_est = rsignum(sym_val);    // equalize.rs:264-276
  }
  float err = FSUB(sym_est, sym_val);
  // evolve: NLMS on both arms  equalize.rs:315-332,354-386  (gain * error * data == (gain*error)*data)
  {
    float ss = 0.0f;
#pragma unroll
    for (int i = 0; i < NFF; ++i) if (i < nff) ss = FADD(ss, FMUL(q.ffw[i], q.ffw[i]));
    float ge = FMUL(__fdiv_rn(p.eq_relax, FADD(p.eq_regul, ss)), err);
#pragma unroll
    for (int i = 0; i < NFF; ++i) if (i < nff) q.ffc[i] = FADD(q.ffc[i], FMUL(ge, q.ffw[nff - 1 - i]));
  }
  {
    float ss = 0.0f;
#pragma unroll
    for (int i = 0; i < NFB; ++i) if (i < nfb) ss = FADD(ss, FMUL(q.fbw[i], q.fbw[i]));
    float ge = FMUL(__fdiv_rn(p.eq_relax, FADD(p.eq_regul, ss)), -err);
#pragma unroll
    for (int i = 0; i < NFB; ++i) if (i < nfb) q.fbc[i] = FADD(q.fbc[i], FMUL(ge, q.fbw[nfb - 1 - i]));
  }
  // feedback_wind.push(&[sym_est, 0.0])  equalize.rs:304
#pragma unroll
  for (int i = 0; i < NFB - 1; ++i) if (i < nfb - 1) q.fbw[i] = q.fbw[i + 1];
  q.fbw[nfb - 1] = sym_est;
#pragma unroll
  for (int i = 0; i < NFB - 1; ++i) if (i < nfb - 1) q.fbw[i] = q.fbw[i + 1];
  q.fbw[nfb - 1] = 0.0f;
  bit = sym_est >= 0.0f;
}

// EXACT: the tap counts equal the template sizes (compile-time constants -> everything stays in registers, and the
// eight symbols are unrolled so that the 16 input samples are read from registers too)
template <int NFF, int NFB, bool EXACT>
__device__ __forceinline__ uint32_t eq_byte_impl(const SameParams& p, uint32_t s, const float* S, uint32_t& flags,
                                                 uint32_t& train_sa, uint32_t& train_cnt) {
  const SameLayout& L = p.layout;
  const int nff = EXACT ? NFF : (int)p.eq_nff, nfb = EXACT ? NFB : (int)p.eq_nfb;
  uint32_t* st = p.state32 + s;
  EqRegs<NFF, NFB> q;
#pragma unroll
  for (int i = 0; i < NFF; ++i) if (i < nff) {
    q.ffc[i] = __uint_as_float(st[(size_t)(L.eq_ffc + i) * L.n_pad]);
    q.ffw[i] = __uint_as_float(st[(size_t)(L.eq_ffw + i) * L.n_pad]);
  }
#pragma unroll
  for (int i = 0; i < NFB; ++i) if (i < nfb) {
    q.fbc[i] = __uint_as_float(st[(size_t)(L.eq_fbc + i) * L.n_pad]);
    q.fbw[i] = __uint_as_float(st[(size_t)(L.eq_fbw + i) * L.n_pad]);
  }
  uint32_t byte = 0;
  if (EXACT) {
    // Same arithmetic as eq_symbol x 8 (equalize.rs:173-186, LSb first), regrouped for instruction-level parallelism:
    // the NLMS step sizes relax / (regul + |window|^2) of all eight symbols depend only on the 16 input samples and on
    // the *squares* of the decisions (always 1.0: a decision is +-1), so they are computed up front, 16 independent
    // chains; only filter -> decision -> error -> tap update remains sequential.
    float X[NFF + 16], X2[NFF + 16];     // feed-forward samples: window (oldest first) then the 16 new ones
    float Y[NFB + 16], Y2[NFB + 16];     // feedback samples: window then (decision, 0) per symbol
#pragma unroll
    for (int i = 0; i < NFF; ++i) X[i] = q.ffw[i];
#pragma unroll
    for (int k = 0; k < 16; ++k) X[NFF + k] = S[k];
#pragma unroll
    for (int k = 0; k < NFF + 16; ++k) X2[k] = FMUL(X[k], X[k]);
#pragma unroll
    for (int i = 0; i < NFB; ++i) { Y[i] = q.fbw[i]; Y2[i] = FMUL(Y[i], Y[i]); }
#pragma unroll
    for (int k = 0; k < 16; ++k) Y2[NFB + k] = (k & 1) ? 0.0f : 1.0f;
    float gff[8], gfb[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      float ss = 0.0f;   // window after pushing (z, s) of symbol b: X[2b+2 .. 2b+1+NFF]   equalize.rs:253, 376-386
#pragma unroll
      for (int i = 0; i < NFF; ++i) ss = FADD(ss, X2[2 * b + 2 + i]);
      gff[b] = __fdiv_rn(p.eq_relax, FADD(p.eq_regul, ss));
      float st2 = 0.0f;  // feedback window before this symbol's decision is pushed: Y[2b .. 2b+NFB-1]
#pragma unroll
      for (int i = 0; i < NFB; ++i) st2 = FADD(st2, Y2[2 * b + i]);
      gfb[b] = __fdiv_rn(p.eq_relax, FADD(p.eq_regul, st2));
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      float ff = 0.0f, fb = 0.0f;        // newest sample pairs with coeff[0]  filter.rs:363-377
#pragma unroll
      for (int i = 0; i < NFF; ++i) ff = FADD(ff, FMUL(X[2 * b + 1 + NFF - i], q.ffc[i]));
#pragma unroll
      for (int i = 0; i < NFB; ++i) fb = FADD(fb, FMUL(Y[2 * b + NFB - 1 - i], q.fbc[i]));
      const float sym_val = FSUB(ff, fb);
      const bool training = (flags & FLAG_EQ_TRAINING) != 0u;                  // equalize.rs:277-300
      const float est_tr = FSUB(FMUL(2.0f, (float)(train_sa & 1u)), 1.0f);
      const float sym_est = training ? est_tr : rsignum(sym_val);              // equalize.rs:264-276
      if (training) {
        train_sa >>= 1;
        train_cnt += 1;
        if (train_cnt >= 32u) flags &= ~FLAG_EQ_TRAINING;
      }
      const float err = FSUB(sym_est, sym_val);
      const float ge_ff = FMUL(gff[b], err), ge_fb = FMUL(gfb[b], -err);       // equalize.rs:315-332, 354-386
#pragma unroll
      for (int i = 0; i < NFF; ++i) q.ffc[i] = FADD(q.ffc[i], FMUL(ge_ff, X[2 * b + 1 + NFF - i]));
#pragma unroll
      for (int i = 0; i < NFB; ++i) q.fbc[i] = FADD(q.fbc[i], FMUL(ge_fb, Y[2 * b + NFB - 1 - i]));
      Y[NFB + 2 * b] = sym_est;                                                // equalize.rs:304
      Y[NFB + 2 * b + 1] = 0.0f;
      byte |= (sym_est >= 0.0f ? 1u : 0u) << b;
    }
#pragma unroll
    for (int i = 0; i < NFF; ++i) q.ffw[i] = X[16 + i];
#pragma unroll
    for (int i = 0; i < NFB; ++i) q.fbw[i] = Y[16 + i];
  } else {
#pragma unroll 1
    for (int b = 0; b < 8; ++b) {
      bool bit;
      eq_symbol<NFF, NFB>(p, q, nff, nfb, S[2 * b], S[2 * b + 1], flags, train_sa, train_cnt, bit);
      byte |= (bit ? 1u : 0u) << b;
    }
  }
#pragma unroll
  for (int i = 0; i < NFF; ++i) if (i < nff) {
    st[(size_t)(L.eq_ffc + i) * L.n_pad] = __float_as_uint(q.ffc[i]);
    st[(size_t)(L.eq_ffw + i) * L.n_pad] = __float_as_uint(q.ffw[i]);
  }
#pragma unroll
  for (int i = 0; i < NFB; ++i) if (i < nfb) {
    st[(size_t)(L.eq_fbc + i) * L.n_pad] = __float_as_uint(q.fbc[i]);
    st[(size_t)(L.eq_fbw + i) * L.n_pad] = __float_as_uint(q.fbw[i]);
  }
  return byte;
}

// rate-generic equalizer (any order up to 16/16): out of line, arrays in local memory
static __device__ __noinline__ uint32_t eq_byte_generic(const SameParams& p, uint32_t s, const float* S, uint32_t& flags,
                                                 uint32_t& train_sa, uint32_t& train_cnt) {
  return eq_byte_impl<SAME_MAX_EQ, SAME_MAX_EQ, false>(p, s, S, flags, train_sa, train_cnt);
}

// Equalizer::reset  equalize.rs:191-196 (mode is kept)
static __device__ __noinline__ void eq_reset(const SameParams& p, uint32_t s) {
  const SameLayout& L = p.layout;
  uint32_t* st = p.state32 + s;
  for (uint32_t i = 0; i < p.eq_nff; ++i) {
    st[(size_t)(L.eq_ffc + i) * L.n_pad] = __float_as_uint(i == 0 ? 1.0f : 0.0f);
    st[(size_t)(L.eq_ffw + i) * L.n_pad] = 0u;
  }
  for (uint32_t i = 0; i < p.eq_nfb; ++i) {
    st[(size_t)(L.eq_fbc + i) * L.n_pad] = __float_as_uint(i == 0 ? 1.0f : 0.0f);
    st[(size_t)(L.eq_fbw + i) * L.n_pad] = 0u;
  }
}

// ----------------------------------------------------------------------------------------------------------------
// A5: one TED instant.  `rem` = until - clock as f32 (receiver.rs:352), `soft` = demodulated sample (demod.rs:163).
// Returns true when the TED emitted a symbol (zero = ted1, sym = ted2).  Updates `until`.
// ----------------------------------------------------------------------------------------------------------------
// `off` = clamp(rem, -0.5, 0.5) (symsync.rs:220) and `offq` = off / samples-per-TED (symsync.rs:225) do not depend on
// the soft symbol; callers may compute them ahead of the matched filter.
__device__ __forceinline__ bool ted_step(Lane& a, const SameParams& p, float soft, float off, float offq) {
  a.ted0 = a.ted1; a.ted1 = a.ted2; a.ted2 = soft;      // symsync.rs:279
  a.tedcnt = (a.tedcnt + 1u) & 1u;                      // symsync.rs:280
  bool have_sym = false;
  if (a.tedcnt == 1u) {
    const float alpha = (a.flags & FLAG_BW_LOCKED) ? p.alpha_l : p.alpha_u;
    const float beta = (a.flags & FLAG_BW_LOCKED) ? p.beta_l : p.beta_u;
    const float terr = FMUL(a.ted1, FSUB(rsignum(a.ted0), rsignum(a.ted2)));   // symsync.rs:311-316
    const float e = rclamp(FSUB(terr, offq), -1.0f, 1.0f);                     // symsync.rs:225
    a.pavg = rclamp(FADD(a.pavg, FMUL(beta, e)), p.pmin, p.pmax);              // symsync.rs:228-229
    a.pinst = FADD(FADD(a.pavg, FMUL(alpha, e)), off);                         // symsync.rs:233
    if (a.pinst < 0.0f) a.pinst = a.pavg;
    have_sym = true;
  } else {
    a.pinst = FADD(a.pinst, off);                                              // symsync.rs:239
  }
  a.until = a.pinst;                                                           // receiver.rs:382
  return have_sym;
}
__device__ __forceinline__ bool ted_step(Lane& a, const SameParams& p, float soft, float rem) {
  const float off = rclamp(rem, -0.5f, 0.5f);           // symsync.rs:220
  return ted_step(a, p, soft, off, __fdiv_rn(off, p.spt));
}

// ----------------------------------------------------------------------------------------------------------------
// A6-A9: one symbol through squelch -> equalizer -> framer -> link/transport events (receiver.rs:407-474, 245-265,
// 291-333).  `n` = input_sample_counter after the sample that produced the symbol.
//
// Split in two so that the expensive byte path can be batched across lanes:
//   symbol_squelch  always runs at once.  Returns SYM_BYTE_READY (bit 0) | adjusted (bit 1) when the squelch emitted
//                   a byte (SquelchState::Ready); otherwise it finishes the symbol itself (symbol_finish).
//   symbol_byte     equalizer + framer for a ready byte, then symbol_finish.  The caller may run it later as long as
//                   the lane consumes no sample in between (nothing the lane can observe changes meanwhile).
// ----------------------------------------------------------------------------------------------------------------
#define SYM_BYTE_READY 1u
#define SYM_ADJUSTED 2u

// SameReceiver::end() + link event + transport for the link state `ls` of this symbol
__device__ __forceinline__ void symbol_finish(Lane& a, const SameParams& p, uint32_t s, StreamBlob* blob, uint32_t ls,
                                              uint32_t burst_len, bool do_end, unsigned long long n) {
  if (do_end) {                                                                // receiver.rs:479-490
    a.flags &= ~(FLAG_AGC_LOCKED | FLAG_SQ_LOCK | FLAG_BW_LOCKED);
    a.byteclk = -1;
    eq_reset(p, s);
    a.ted0 = a.ted1 = a.ted2 = 0.0f; a.tedcnt = 0;                             // symsync.rs:166-170
    a.pavg = p.spt; a.pinst = p.spt;
  }

  // link event  receiver.rs:245-253 (a Burst never equals the previous state: a NoCarrier symbol always separates bursts)
  if (ls != a.link_last || ls == 3u) {
    a.link_last = ls;
    if (ls == 3u)
      emit_event(p, a.ev, SAME_EV_LINK_BURST, 0, n, a.symcount, blob->burst, burst_len, min(burst_len, SAME_BURST_CAP), 0, 0,
                 burst_len > SAME_BURST_CAP ? SAME_EV_FLAG_TRUNCATED : 0u);
    else
      emit_event(p, a.ev, ls, 0, n, a.symcount, nullptr, 0, 0, 0, 0, 0);
  }

  // transport  receiver.rs:291-333
  if (ls == 3u || (ls == 0u && ((a.tr.have_eom && n > a.tr.eom_at) || a.symcount >= a.tr.next_deadline ||
                                (a.tr.hist_n ? 1u : 0u) != a.tr.tr_state))) {
    Transport tr = a.tr;   // copy: the assembler functions are noinline and take it by reference
    uint32_t tk; MsgResult mr; mr.kind = 0; mr.err = 0; mr.len = 0; mr.parity = 0; mr.voting = 0; mr.offset = 0;
    if (ls == 3u) {
      tk = assembler_assemble(p, blob, tr, burst_len, a.symcount, mr);
    } else if (tr.have_eom && n > tr.eom_at) {
      tk = 2; mr.kind = 1; mr.len = 4;                                         // forced EndOfMessage receiver.rs:300-309
    } else if (a.symcount >= tr.next_deadline) {
      tk = assembler_idle(p, blob, tr, a.symcount, mr);
    } else {
      tk = tr.hist_n ? 1u : 0u;                                                // nothing expired: idle() is a no-op
    }
    if (tk == 2u) {
      if (mr.kind == 0u) { tr.have_eom = true; tr.eom_at = n + p.force_eom_samples; }  // receiver.rs:318-325
      else if (mr.kind == 1u) tr.have_eom = false;                                     // receiver.rs:326-328
      // a Message state always differs from the previous transport state (an idle poll separates messages)
      tr.tr_state = 2;
      if (mr.kind == 0u)
        emit_event(p, a.ev, SAME_EV_TR_MSG_SOM, 0, n, a.symcount, blob->pending_text, mr.len, mr.len, mr.parity, mr.voting, 0);
      else if (mr.kind == 1u)
        emit_event(p, a.ev, SAME_EV_TR_MSG_EOM, 0, n, a.symcount, (const uint8_t*)"NNNN", 4, 4, 0, 0, 0);
      else
        emit_event(p, a.ev, SAME_EV_TR_MSG_ERR, mr.err, n, a.symcount, nullptr, 0, 0, 0, 0, 0);
    } else if (tk != tr.tr_state) {
      tr.tr_state = tk;
      emit_event(p, a.ev, tk == 0u ? SAME_EV_TR_IDLE : SAME_EV_TR_ASSEMBLING, 0, n, a.symcount, nullptr, 0, 0, 0, 0, 0);
    }
    a.tr = tr;
  }
}

__device__ __forceinline__ uint32_t symbol_squelch(Lane& a, const SameParams& p, uint32_t s, uint32_t* st,
                                                   StreamBlob* blob, float z, float sy, unsigned long long n) {
  const SameLayout& L = p.layout;
  if (p.trace && a.trace_n < p.trace_cap) {
    same_soft_symbol t; t.input_sample_counter = n; t.zero = z; t.sym = sy;
    p.trace[(size_t)s * p.trace_cap + a.trace_n] = t;
    a.trace_n += 1;
  }
  // squelch  codesquelch.rs:228-304 (sample history: 64-entry ring kept in place in the state words)
  LANE_ST(st, L, L.sqh + (a.sq_head & 63u)) = __float_as_uint(z);
  LANE_ST(st, L, L.sqh + ((a.sq_head + 1u) & 63u)) = __float_as_uint(sy);
  a.sq_head = (a.sq_head + 2u) & 63u;
  a.sq_data = (a.sq_data >> 1) | ((sy >= 0.0f) ? 0x80000000u : 0u);            // codesquelch.rs:421-428
  const uint32_t cerr = __popc(a.sq_data ^ p.sq_sync_word);
  a.sq_power = FADD(a.sq_power, FMUL(FSUB(FMUL(sy, sy), a.sq_power), p.sq_bw));  // codesquelch.rs:483-488
  a.sq_power = fmaxf(a.sq_power, 0.0f);
  a.sq_pflags = (a.sq_pflags >> 1) | ((a.sq_power >= p.sq_close) ? 0x80000000u : 0u);
  a.symcount += 1;

  // Squelch decision (codesquelch.rs:236-304) as straight-line selects: lanes of a warp are in different squelch
  // states all the time during bursts, so separate branches per state would all be executed one after the other.
  const bool full = a.symcount >= 32ull;                                       // sample_history.is_full()
  const bool acquire = full && !(a.flags & FLAG_SQ_LOCK) && cerr <= p.sq_max_err && a.sq_power >= p.sq_open;
  const bool adjusted = acquire && a.byteclk != 0;                             // codesquelch.rs:243-267
  const bool dropped = full && !acquire && a.byteclk >= 0 && !(a.sq_pflags & 1u);   // codesquelch.rs:270-277
  int bc = acquire ? 0 : a.byteclk;
  if (dropped) bc = -1;
  const bool synced = full && bc >= 0;
  if (synced && bc == 0) {
    a.byteclk = 1;                                                             // SquelchState::Ready(adjusted, ..)
    return SYM_BYTE_READY | (adjusted ? SYM_ADJUSTED : 0u);
  }
  a.byteclk = synced ? ((bc + 1) & 7) : bc;
  // Reading -> framer.state() (receiver.rs:419-422); NoCarrier / DroppedCarrier -> framer.end() (receiver.rs:410-418)
  uint32_t burst_len = 0;
  uint32_t ls = framer_state(a.fr);
  if (!synced) ls = framer_end(a.fr, burst_len);
  const bool do_end = dropped;                                                 // SameReceiver::end()  receiver.rs:414-418
  symbol_finish(a, p, s, blob, ls, burst_len, do_end, n);
  return 0u;
}

__device__ __forceinline__ void symbol_byte(Lane& a, const SameParams& p, uint32_t s, uint32_t* st, StreamBlob* blob,
                                            bool adjusted, unsigned long long n) {
  const SameLayout& L = p.layout;
  float S[16];
#pragma unroll
  for (int j = 0; j < 16; ++j)                                                 // oldest 16  codesquelch.rs:288-294
    S[j] = __uint_as_float(LANE_ST(st, L, L.sqh + ((a.sq_head + j) & 63u)));
  if (adjusted) {                                                              // receiver.rs:423-438
    a.flags |= FLAG_AGC_LOCKED | FLAG_BW_LOCKED | FLAG_EQ_TRAINING;
    a.train_sa = p.sq_sync_word; a.train_cnt = 0;
  }
  uint32_t byte;
  if (p.eq_nff == 6u && p.eq_nfb == 4u) {
    byte = eq_byte_impl<6, 4, true>(p, s, S, a.flags, a.train_sa, a.train_cnt);   // inlined: S and the taps stay in registers
  } else {  // by-reference arguments go through short-lived temporaries so that the lane state stays in registers
    uint32_t fl = a.flags, tsa = a.train_sa, tcn = a.train_cnt;
    byte = eq_byte_generic(p, s, S, fl, tsa, tcn);
    a.flags = fl; a.train_sa = tsa; a.train_cnt = tcn;
  }
  uint32_t burst_len = 0;
  bool do_end = false;
  const uint32_t ls = framer_input(p, blob, a.fr, byte, adjusted, burst_len);  // receiver.rs:457-459
  if (ls == 2u) a.flags |= FLAG_SQ_LOCK;                                       // receiver.rs:461-465
  else if (ls == 0u || ls == 3u) do_end = true;                                // receiver.rs:466-469
  symbol_finish(a, p, s, blob, ls, burst_len, do_end, n);
}

// Undeferred form (generic kernel)
__device__ __forceinline__ void symbol_step(Lane& a, const SameParams& p, uint32_t s, uint32_t* st, StreamBlob* blob,
                                            float z, float sy, unsigned long long n) {
  const uint32_t r = symbol_squelch(a, p, s, st, blob, z, sy, n);
  if (r & SYM_BYTE_READY) symbol_byte(a, p, s, st, blob, (r & SYM_ADJUSTED) != 0u, n);
}

}  // namespace same_dev
