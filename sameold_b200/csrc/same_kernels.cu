// same_kernels.cu — the receiver kernel: one SAME receiver per lane, 32 independent streams per warp.
//
// Path (reference crates/sameold/src/receiver.rs:233-490):
//   A0 s16 -> f32                      samedec app.rs:112
//   A1 DC blocker                      dcblock.rs:45-49,104-108
//   A2 AGC                             agc.rs:72-77
//   A3 demod window                    demod.rs:177-179
//   A4 mark/space matched filter       demod.rs:156-164, filter.rs:363-377
//   A5 TED scheduler + timing loop     receiver.rs:347-360, symsync.rs:198-244,278-322
//   A6 code + power squelch            codesquelch.rs:228-304
//   A7 NLMS DFE                        equalize.rs:173-186,249-386
//   A8 framer + link glue              framing.rs:109-197, receiver.rs:407-490
//   A9 link/transport events           receiver.rs:245-265,291-333
//
// Every f32 operation is a single IEEE round-to-nearest operation in the reference's order: explicit __fmul_rn /
// __fadd_rn / __fsub_rn / __fdiv_rn (never contracted), no FTZ, hypot fixed as (float)sqrt((double)re^2 + (double)im^2)
// (== glibc hypotf).  The chain is numerically chaotic: results are bit-identical to the oracle or
// they are different, there is no tolerance.
//
// Loop structure: the warp advances in "rounds".  In each round every lane consumes its own samples up to its next
// timing-error-detector (TED) instant (19..24 samples, trip count divergence only), then all lanes evaluate the
// matched filter + timing loop together (converged), then the lanes whose TED emitted a symbol run the squelch, and
// the few lanes that completed a byte run the equalizer/framer.  Lane sample cursors drift apart; nothing is shared
// between lanes.
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

// Complex<f32>::norm() = hypotf(re, im) (demod.rs:163); fixed definition shared with the oracle
__device__ __forceinline__ float hypot_fixed(float re, float im) {
  double a = (double)re, b = (double)im;
  return __double2float_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b))));
}

// Smallest clock value c >= 1 at which receiver.rs:352-353 fires:  r = until - c as f32;  r <= 0 || |r| < 0.5
__device__ __forceinline__ bool fires(float until, int c) {
  float r = FSUB(until, (float)c);
  return r <= 0.0f || fabsf(r) < 0.5f;
}
__device__ __forceinline__ int fire_clock(float until, int clock_now) {
  float t = until - 0.5f;
  int c = (t < 1.0e6f) ? (int)floorf(t) + 1 : 1000001;
  if (c < 1) c = 1;
  int guard = 0;
  while (!fires(until, c) && guard < 64) { ++c; ++guard; }        // exactness fix-ups (normally 0 iterations)
  while (c > 1 && fires(until, c - 1) && guard < 128) { --c; ++guard; }
  if (c <= clock_now) c = clock_now + 1;  // the predicate is monotone in c: already-passed instants fire on the next sample
  return c;
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

// EXACT: the tap counts equal the template sizes (compile-time constants -> everything stays in registers)
template <int NFF, int NFB, bool EXACT>
__device__ __noinline__ uint32_t eq_byte(const SameParams& p, uint32_t s, const float* S, uint32_t& flags,
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
#pragma unroll 1
  for (int b = 0; b < 8; ++b) {  // equalize.rs:173-186, LSb first
    bool bit;
    eq_symbol<NFF, NFB>(p, q, nff, nfb, S[2 * b], S[2 * b + 1], flags, train_sa, train_cnt, bit);
    byte |= (bit ? 1u : 0u) << b;
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

// Equalizer::reset  equalize.rs:191-196 (mode is kept)
__device__ __noinline__ void eq_reset(const SameParams& p, uint32_t s) {
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
// The receiver kernel.  RW = demod window ring slots (power of two >= ntaps); RD = DC ring slots (power of two >= dc_len)
// ----------------------------------------------------------------------------------------------------------------
template <int RW, int RD>
__global__ void __launch_bounds__(32) same_rx_kernel(const __grid_constant__ SameParams p,
                                                     const __grid_constant__ SameTaps taps,
                                                     const int16_t* __restrict__ samples,
                                                     const unsigned long long* __restrict__ offsets,
                                                     const uint32_t* __restrict__ lengths) {
  extern __shared__ float smem[];
  float* win = smem;             // [RW][32]  AGC output window   (A3)
  float* sqh = win + RW * 32;    // [64][32]  squelch sample history (A6)
  float* dcf = sqh + 64 * 32;    // [RD][32]  DC feed-forward window (A1)
  float* dcb = dcf + RD * 32;    // [RD][32]  DC feedback window

  const SameLayout& L = p.layout;
  const int lane = threadIdx.x;
  const uint32_t s = blockIdx.x * 32u + lane;
  const bool valid = s < p.n_streams;
  const uint32_t sidx = valid ? s : 0u;  // padded lanes read stream 0's words but never write
  uint32_t* st = p.state32 + sidx;
  StreamBlob* blob = p.blobs + sidx;
#define ST(w) st[(size_t)(w) * L.n_pad]

  const uint32_t len = valid ? lengths[s] : 0u;
  if (__all_sync(0xffffffffu, len == 0u)) return;
  const int16_t* src = (samples != nullptr && valid) ? samples + offsets[s] : nullptr;

  // ---- load state ----
  float g = __uint_as_float(ST(F_AGC_GAIN));
  uint32_t flags = ST(F_FLAGS);
  int clock = (int)ST(F_CLOCK);
  float until = __uint_as_float(ST(F_UNTIL));
  float pavg = __uint_as_float(ST(F_PAVG)), pinst = __uint_as_float(ST(F_PINST));
  float ted0 = __uint_as_float(ST(F_TED0)), ted1 = __uint_as_float(ST(F_TED1)), ted2 = __uint_as_float(ST(F_TED2));
  uint32_t tedcnt = ST(F_TEDCNT);
  uint32_t sq_data = ST(F_SQ_DATA);
  float sq_power = __uint_as_float(ST(F_SQ_POWER));
  uint32_t sq_pflags = ST(F_SQ_PFLAGS);
  int byteclk = (int)ST(F_SQ_BYTECLK);
  unsigned long long symcount = ((unsigned long long)ST(F_SYMCOUNT_HI) << 32) | ST(F_SYMCOUNT_LO);
  unsigned long long n0 = ((unsigned long long)ST(F_N_HI) << 32) | ST(F_N_LO);
  uint32_t train_sa = ST(F_EQ_TRAIN_SA), train_cnt = ST(F_EQ_TRAIN_CNT);
  Framer fr;
  fr.st = (flags >> FLAG_FR_SHIFT) & 3u; fr.word = ST(F_FR_WORD); fr.count = ST(F_FR_COUNT);
  fr.invalid = ST(F_FR_INVALID); fr.msglen = ST(F_FR_MSGLEN);
  uint32_t link_last = (flags >> FLAG_LINK_SHIFT) & 3u;
  Transport tr;
  tr.hist_n = ST(F_HIST_N); tr.pending = flags & FLAG_PENDING; tr.have_prev = flags & FLAG_HAVE_PREV;
  tr.have_eom = flags & FLAG_FORCE_EOM; tr.tr_state = (flags >> FLAG_TR_SHIFT) & 3u;
  tr.next_deadline = ((unsigned long long)ST(F_TRNEXT_HI) << 32) | ST(F_TRNEXT_LO);
  tr.eom_at = ((unsigned long long)ST(F_EOM_HI) << 32) | ST(F_EOM_LO);
  EvCtx ev; ev.p = &p; ev.stream = s; ev.seq = ST(F_SEQ);
  uint32_t trace_n = ST(F_TRACE_N);
  float ffsum = __uint_as_float(ST(F_DC_FFSUM)), fbsum = __uint_as_float(ST(F_DC_FBSUM));

  const int dcl = (int)p.dc_len, ntaps = (int)p.ntaps;
  for (int i = 0; i < dcl; ++i) {
    dcf[i * 32 + lane] = __uint_as_float(ST(L.dc_ff + i));
    dcb[i * 32 + lane] = __uint_as_float(ST(L.dc_fb + i));
  }
  for (int i = 0; i < ntaps; ++i) win[i * 32 + lane] = __uint_as_float(ST(L.win + i));
  for (int i = 0; i < 64; ++i) sqh[i * 32 + lane] = __uint_as_float(ST(L.sqh + i));
  int dch = dcl;    // next DC write slot (ring index grows monotonically, masked on use)
  int wh = ntaps;   // next window write slot
  int sh = 0;       // next squelch-history write slot == oldest entry
  __syncwarp();

  const float inv_len = p.dc_inv_len, gate = p.dc_gate;
  const float bw = p.agc_bw, gmin = p.agc_min, gmax = p.agc_max;

  uint32_t pos = 0;
  int cfire = fire_clock(until, clock);

  while (__any_sync(0xffffffffu, pos < len)) {
    // ---------------- segment: this lane's samples up to its next TED instant (A0-A3) ----------------
    int nseg = 0;
    if (pos < len) nseg = min(cfire - clock, (int)min(len - pos, 1u << 20));
    const int maxseg = __reduce_max_sync(0xffffffffu, nseg);
    for (int k = 0; k < maxseg; ++k) {
      if (k < nseg) {
        float x = src ? (float)src[pos + k] : 0.0f;
        // DC blocker  dcblock.rs:45-49,104-108
        const int ia = ((dch - dcl) & (RD - 1)) * 32 + lane;
        const int iw = (dch & (RD - 1)) * 32 + lane;
        const int ifr = ((dch - dcl + 1) & (RD - 1)) * 32 + lane;
        float aged = dcf[ia];
        dcf[iw] = x;
        ffsum = FADD(ffsum, FSUB(x, aged));
        float ma0 = FMUL(ffsum, inv_len);
        float sig = dcf[ifr];
        float aged2 = dcb[ia];
        dcb[iw] = ma0;
        fbsum = FADD(fbsum, FSUB(ma0, aged2));
        float ma1 = FMUL(fbsum, inv_len);
        dch += 1;
        float d = FSUB(sig, FMUL(gate, ma1));
        // AGC  agc.rs:72-77
        float y = FMUL(d, g);
        float u = (flags & FLAG_AGC_LOCKED) ? 0.0f : 1.0f;
        g = FADD(g, FMUL(FMUL(u, FSUB(1.0f, fabsf(y))), bw));
        g = rclamp(g, gmin, gmax);
        // demod window  demod.rs:177-179
        win[(wh & (RW - 1)) * 32 + lane] = y;
        wh += 1;
      }
    }
    pos += (uint32_t)nseg;
    clock += nseg;
    const bool fire = (nseg > 0) && (clock == cfire);
    if (!__any_sync(0xffffffffu, fire)) continue;

    // ---------------- TED instant (A4, A5) ----------------
    float soft = 0.0f, rem = 0.0f;
    {
      float mr = 0.0f, mi = 0.0f, sr = 0.0f, si = 0.0f;
      for (int i = 0; i < ntaps; ++i) {  // newest sample pairs with tap 0, sequential accumulation  filter.rs:363-377
        float v = win[((wh - 1 - i) & (RW - 1)) * 32 + lane];
        mr = FADD(mr, FMUL(v, taps.mark_re[i]));
        mi = FADD(mi, FMUL(v, taps.mark_im[i]));
        sr = FADD(sr, FMUL(v, taps.space_re[i]));
        si = FADD(si, FMUL(v, taps.space_im[i]));
      }
      soft = rclamp(FSUB(hypot_fixed(mr, mi), hypot_fixed(sr, si)), -1.0f, 1.0f);  // demod.rs:163
    }
    bool have_sym = false;
    if (fire) {
      rem = FSUB(until, (float)clock);  // receiver.rs:352
      clock = 0;
      ted0 = ted1; ted1 = ted2; ted2 = soft;          // symsync.rs:279
      tedcnt = (tedcnt + 1u) & 1u;                    // symsync.rs:280
      float off = rclamp(rem, -0.5f, 0.5f);           // symsync.rs:220
      if (tedcnt == 1u) {
        const float alpha = (flags & FLAG_BW_LOCKED) ? p.alpha_l : p.alpha_u;
        const float beta = (flags & FLAG_BW_LOCKED) ? p.beta_l : p.beta_u;
        float terr = FMUL(ted1, FSUB(rsignum(ted0), rsignum(ted2)));           // symsync.rs:311-316
        float e = rclamp(FSUB(terr, __fdiv_rn(off, p.spt)), -1.0f, 1.0f);      // symsync.rs:225
        pavg = rclamp(FADD(pavg, FMUL(beta, e)), p.pmin, p.pmax);              // symsync.rs:228-229
        pinst = FADD(FADD(pavg, FMUL(alpha, e)), off);                         // symsync.rs:233
        if (pinst < 0.0f) pinst = pavg;
        have_sym = true;
      } else {
        pinst = FADD(pinst, off);                                              // symsync.rs:239
      }
      until = pinst;                                                           // receiver.rs:382
      cfire = fire_clock(until, 0);
    }
    if (!__any_sync(0xffffffffu, have_sym)) continue;

    // ---------------- symbol (A6-A9) ----------------
    if (have_sym) {
      const unsigned long long n = n0 + pos;  // input_sample_counter after this sample
      const float z = ted1, sy = ted2;
      if (p.trace && trace_n < p.trace_cap) {
        same_soft_symbol t; t.input_sample_counter = n; t.zero = z; t.sym = sy;
        p.trace[(size_t)s * p.trace_cap + trace_n] = t;
        trace_n += 1;
      }
      // squelch  codesquelch.rs:228-304
      sqh[(sh & 63) * 32 + lane] = z;
      sqh[((sh + 1) & 63) * 32 + lane] = sy;
      sh += 2;
      sq_data = (sq_data >> 1) | ((sy >= 0.0f) ? 0x80000000u : 0u);            // codesquelch.rs:421-428
      const uint32_t cerr = __popc(sq_data ^ p.sq_sync_word);
      sq_power = FADD(sq_power, FMUL(FSUB(FMUL(sy, sy), sq_power), p.sq_bw));  // codesquelch.rs:483-488
      sq_power = fmaxf(sq_power, 0.0f);
      sq_pflags = (sq_pflags >> 1) | ((sq_power >= p.sq_close) ? 0x80000000u : 0u);
      symcount += 1;

      uint32_t ls;                 // link state kind returned for this symbol
      uint32_t burst_len = 0;      // valid when ls == 3
      bool do_end = false;         // SameReceiver::end()  receiver.rs:479-490
      if (symcount < 32ull) {
        ls = framer_end(fr, burst_len);                                        // NoCarrier: receiver.rs:410-413
      } else {
        bool adjusted = false, dropped = false;
        if (!(flags & FLAG_SQ_LOCK) && cerr <= p.sq_max_err && sq_power >= p.sq_open) {
          adjusted = (byteclk != 0);                                           // codesquelch.rs:243-267
          byteclk = 0;
        } else if (byteclk >= 0 && !(sq_pflags & 1u)) {
          dropped = true;                                                      // codesquelch.rs:270-277
        }
        if (dropped) {
          byteclk = -1; do_end = true;
          ls = framer_end(fr, burst_len);                                      // receiver.rs:414-418
        } else if (byteclk < 0) {
          ls = framer_end(fr, burst_len);                                      // NoCarrier
        } else if (byteclk != 0) {
          byteclk = (byteclk + 1) & 7;
          ls = framer_state(fr);                                               // Reading: receiver.rs:419-422
        } else {
          byteclk = 1;
          float S[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) S[j] = sqh[((sh + j) & 63) * 32 + lane];  // oldest 16  codesquelch.rs:288-294
          if (adjusted) {                                                      // receiver.rs:423-438
            flags |= FLAG_AGC_LOCKED | FLAG_BW_LOCKED | FLAG_EQ_TRAINING;
            train_sa = p.sq_sync_word; train_cnt = 0;
          }
          uint32_t byte;
          if (p.eq_nff == 6u && p.eq_nfb == 4u) byte = eq_byte<6, 4, true>(p, s, S, flags, train_sa, train_cnt);
          else byte = eq_byte<SAME_MAX_EQ, SAME_MAX_EQ, false>(p, s, S, flags, train_sa, train_cnt);
          ls = framer_input(p, blob, fr, byte, adjusted, burst_len);           // receiver.rs:457-459
          if (ls == 2u) flags |= FLAG_SQ_LOCK;                                 // receiver.rs:461-465
          else if (ls == 0u || ls == 3u) do_end = true;                        // receiver.rs:466-469
        }
      }
      if (do_end) {
        flags &= ~(FLAG_AGC_LOCKED | FLAG_SQ_LOCK | FLAG_BW_LOCKED);
        byteclk = -1;
        eq_reset(p, s);
        ted0 = ted1 = ted2 = 0.0f; tedcnt = 0;                                 // symsync.rs:166-170
        pavg = p.spt; pinst = p.spt;
      }

      // link event  receiver.rs:245-253 (a Burst never equals the previous state: a NoCarrier symbol always separates bursts)
      if (ls != link_last || ls == 3u) {
        link_last = ls;
        if (ls == 3u)
          emit_event(ev, SAME_EV_LINK_BURST, 0, n, symcount, blob->burst, burst_len, min(burst_len, SAME_BURST_CAP), 0, 0,
                     burst_len > SAME_BURST_CAP ? SAME_EV_FLAG_TRUNCATED : 0u);
        else
          emit_event(ev, ls, 0, n, symcount, nullptr, 0, 0, 0, 0, 0);
      }

      // transport  receiver.rs:291-333
      if (ls == 3u || ls == 0u) {
        uint32_t tk; MsgResult mr; mr.kind = 0; mr.err = 0; mr.len = 0; mr.parity = 0; mr.voting = 0; mr.offset = 0;
        bool forced = false;
        if (ls == 3u) {
          tk = assembler_assemble(p, blob, tr, burst_len, symcount, mr);
        } else if (tr.have_eom && n > tr.eom_at) {
          tk = 2; mr.kind = 1; mr.len = 4; forced = true;                      // forced EndOfMessage receiver.rs:300-309
        } else if (symcount >= tr.next_deadline) {
          tk = assembler_idle(p, blob, tr, symcount, mr);
        } else {
          tk = tr.hist_n ? 1u : 0u;                                            // nothing expired: idle() is a no-op
        }
        if (tk == 2u) {
          if (mr.kind == 0u) { tr.have_eom = true; tr.eom_at = n + p.force_eom_samples; }  // receiver.rs:318-325
          else if (mr.kind == 1u) tr.have_eom = false;                                     // receiver.rs:326-328
          // a Message state always differs from the previous transport state (an idle poll separates messages)
          tr.tr_state = 2;
          if (mr.kind == 0u)
            emit_event(ev, SAME_EV_TR_MSG_SOM, 0, n, symcount, blob->pending_text, mr.len, mr.len, mr.parity, mr.voting, 0);
          else if (mr.kind == 1u)
            emit_event(ev, SAME_EV_TR_MSG_EOM, 0, n, symcount, (const uint8_t*)"NNNN", 4, 4, 0, 0, forced ? 0u : 0u);
          else
            emit_event(ev, SAME_EV_TR_MSG_ERR, mr.err, n, symcount, nullptr, 0, 0, 0, 0, 0);
        } else if (tk != tr.tr_state) {
          tr.tr_state = tk;
          emit_event(ev, tk == 0u ? SAME_EV_TR_IDLE : SAME_EV_TR_ASSEMBLING, 0, n, symcount, nullptr, 0, 0, 0, 0, 0);
        }
      }
    }
  }

  if (!valid || len == 0u) return;

  // ---- store state ----
  const unsigned long long n1 = n0 + len;
  flags &= ~((3u << FLAG_FR_SHIFT) | (3u << FLAG_LINK_SHIFT) | (3u << FLAG_TR_SHIFT) | FLAG_PENDING | FLAG_HAVE_PREV | FLAG_FORCE_EOM);
  flags |= (fr.st << FLAG_FR_SHIFT) | (link_last << FLAG_LINK_SHIFT) | (tr.tr_state << FLAG_TR_SHIFT);
  if (tr.pending) flags |= FLAG_PENDING;
  if (tr.have_prev) flags |= FLAG_HAVE_PREV;
  if (tr.have_eom) flags |= FLAG_FORCE_EOM;
  ST(F_AGC_GAIN) = __float_as_uint(g);
  ST(F_FLAGS) = flags;
  ST(F_CLOCK) = (uint32_t)clock;
  ST(F_UNTIL) = __float_as_uint(until);
  ST(F_PAVG) = __float_as_uint(pavg); ST(F_PINST) = __float_as_uint(pinst);
  ST(F_TED0) = __float_as_uint(ted0); ST(F_TED1) = __float_as_uint(ted1); ST(F_TED2) = __float_as_uint(ted2);
  ST(F_TEDCNT) = tedcnt;
  ST(F_SQ_DATA) = sq_data; ST(F_SQ_POWER) = __float_as_uint(sq_power); ST(F_SQ_PFLAGS) = sq_pflags;
  ST(F_SQ_BYTECLK) = (uint32_t)byteclk;
  ST(F_SYMCOUNT_LO) = (uint32_t)symcount; ST(F_SYMCOUNT_HI) = (uint32_t)(symcount >> 32);
  ST(F_N_LO) = (uint32_t)n1; ST(F_N_HI) = (uint32_t)(n1 >> 32);
  ST(F_EQ_TRAIN_SA) = train_sa; ST(F_EQ_TRAIN_CNT) = train_cnt;
  ST(F_FR_WORD) = fr.word; ST(F_FR_COUNT) = fr.count; ST(F_FR_INVALID) = fr.invalid; ST(F_FR_MSGLEN) = fr.msglen;
  ST(F_EOM_LO) = (uint32_t)tr.eom_at; ST(F_EOM_HI) = (uint32_t)(tr.eom_at >> 32);
  ST(F_TRNEXT_LO) = (uint32_t)tr.next_deadline; ST(F_TRNEXT_HI) = (uint32_t)(tr.next_deadline >> 32);
  ST(F_HIST_N) = tr.hist_n;
  ST(F_SEQ) = ev.seq;
  ST(F_TRACE_N) = trace_n;
  ST(F_DC_FFSUM) = __float_as_uint(ffsum); ST(F_DC_FBSUM) = __float_as_uint(fbsum);
  for (int i = 0; i < dcl; ++i) {  // canonical order: oldest first
    ST(L.dc_ff + i) = __float_as_uint(dcf[((dch - dcl + i) & (RD - 1)) * 32 + lane]);
    ST(L.dc_fb + i) = __float_as_uint(dcb[((dch - dcl + i) & (RD - 1)) * 32 + lane]);
  }
  for (int i = 0; i < ntaps; ++i) ST(L.win + i) = __float_as_uint(win[((wh - ntaps + i) & (RW - 1)) * 32 + lane]);
  for (int i = 0; i < 64; ++i) ST(L.sqh + i) = __float_as_uint(sqh[((sh + i) & 63) * 32 + lane]);
#undef ST
}

// Constructor state (receiver.rs:502-560) or SameReceiver::reset (receiver.rs:182-198) for the selected streams.
// `ids` == nullptr: all streams.  `after_reset`: AGC gain 1.0 (agc.rs:61) instead of min(1, min_gain) (agc.rs:55).
__global__ void same_init_kernel(const __grid_constant__ SameParams p, const uint32_t* __restrict__ ids, uint32_t n,
                                 int after_reset) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = ids ? ids[i] : i;
  if (s >= p.n_streams) return;
  const SameLayout& L = p.layout;
  uint32_t* st = p.state32 + s;
  for (uint32_t w = 0; w < L.n_words; ++w) st[(size_t)w * L.n_pad] = 0u;
  st[(size_t)F_AGC_GAIN * L.n_pad] = __float_as_uint(after_reset ? 1.0f : p.agc_gain0);
  st[(size_t)F_UNTIL * L.n_pad] = __float_as_uint(p.spt);
  st[(size_t)F_PAVG * L.n_pad] = __float_as_uint(p.spt);
  st[(size_t)F_PINST * L.n_pad] = __float_as_uint(p.spt);
  st[(size_t)F_SQ_BYTECLK * L.n_pad] = (uint32_t)(-1);
  st[(size_t)F_TRNEXT_LO * L.n_pad] = 0xffffffffu;
  st[(size_t)F_TRNEXT_HI * L.n_pad] = 0xffffffffu;
  st[(size_t)L.eq_ffc * L.n_pad] = __float_as_uint(1.0f);  // identity taps  equalize.rs:131-132
  st[(size_t)L.eq_fbc * L.n_pad] = __float_as_uint(1.0f);
}

}  // namespace same_dev

// ----------------------------------------------------------------------------------------------------------------
// Launchers (called from same_engine.cu)
// ----------------------------------------------------------------------------------------------------------------
extern "C" cudaError_t same_launch_rx(const SameParams* p, const SameTaps* taps, const int16_t* d_samples,
                                      const unsigned long long* d_offsets, const uint32_t* d_lengths,
                                      cudaStream_t stream) {
  const uint32_t blocks = (p->n_streams + 31u) / 32u;
  if (p->ntaps <= 64 && p->dc_len <= 16) {
    const size_t smem = (size_t)(64 + 64 + 2 * 16) * 32 * sizeof(float);
    same_dev::same_rx_kernel<64, 16><<<blocks, 32, smem, stream>>>(*p, *taps, d_samples, d_offsets, d_lengths);
  } else {
    const size_t smem = (size_t)(128 + 64 + 2 * 64) * 32 * sizeof(float);
    same_dev::same_rx_kernel<128, 64><<<blocks, 32, smem, stream>>>(*p, *taps, d_samples, d_offsets, d_lengths);
  }
  return cudaGetLastError();
}

extern "C" cudaError_t same_launch_init(const SameParams* p, const uint32_t* d_ids, uint32_t n, int after_reset,
                                        cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  same_dev::same_init_kernel<<<(n + 127u) / 128u, 128, 0, stream>>>(*p, d_ids, n, after_reset);
  return cudaGetLastError();
}
