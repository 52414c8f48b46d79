// same_oracle.hpp — CPU ORACLE for the sameold receiver path.  TEST INFRASTRUCTURE ONLY.
//
// This is a C++17 restatement of the reference's algorithm (cbs228/sameold 0.6.0, Rust),
// one object per stream, every f32 operation in the reference's order.  It exists to CHECK
// the CUDA engine; nothing under sameold_b200/ may include, link or call it.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// Parity pinning: the restatement is pinned by the reference's own golden files
// (sample/*.22050.s16le.{bin,txt}, copies under tests/golden/) and by the known-answer
// values of the reference's in-crate unit tests (mirrored in oracle/selftest.cpp).
// Intermediate soft symbols / event sample indices are NOT pinned by any reference test
// bit-for-bit (they depend on the platform libm behind Rust's f32::hypot); this file fixes
// hypot as (float)sqrt((double)re*re + (double)im*im), which equals glibc hypotf.
// The Rust reference cannot be built here (no cargo/rustc, no network, crates not vendored).
//
// Build with:  g++ -std=c++17 -O2 -ffp-contract=off -fno-fast-math   (never -ffast-math, no FTZ/DAZ)
//
// Citations are to /root/reference/crates/... (file:line).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <functional>
#include <optional>
#include <string>
#include <vector>

namespace same_oracle {

// ---------------------------------------------------------------------------------------
// Rust f32 semantics helpers
// ---------------------------------------------------------------------------------------

// f32::clamp: NaN passes through, comparisons only (core::f32::clamp)
static inline float clampf(float x, float lo, float hi) {
  if (x < lo) x = lo;
  if (x > hi) x = hi;
  return x;
}

// f32::signum: +0 -> +1, -0 -> -1, NaN -> NaN
static inline float signumf(float x) {
  if (std::isnan(x)) return x;
  return std::copysign(1.0f, x);
}

// Complex::norm() == re.hypot(im) -> libm hypotf.  Fixed definition (see header).
static inline float hypot_fixed(float re, float im) {
  return (float)std::sqrt((double)re * (double)re + (double)im * (double)im);
}

// `x as usize` for f32: truncates toward zero, saturates, NaN -> 0
static inline size_t f32_as_usize(float x) {
  if (!(x > 0.0f)) return 0;
  if (x >= 18446744073709551616.0f) return SIZE_MAX;
  return (size_t)x;
}

// ---------------------------------------------------------------------------------------
// waveform.rs:6-64 — constants and matched filter taps
// ---------------------------------------------------------------------------------------
constexpr float FSK_MARK_HZ = 2083.3f;                 // waveform.rs:6
constexpr float FSK_SPACE_HZ = 1562.5f;                // waveform.rs:9
constexpr float BAUD_HZ = 520.83f;                     // waveform.rs:12
constexpr uint8_t PREAMBLE = 0xab;                     // waveform.rs:19
constexpr uint32_t PREAMBLE_SYNC_WORD = 0xabababab;    // waveform.rs:26

static inline float samples_per_symbol(uint32_t fs) { return (float)fs / BAUD_HZ; }  // waveform.rs:29-31

struct Cf32 { float re, im; };

// waveform.rs:54-64.  Complex::new(0, th).exp() == from_polar(exp(0)=1, th) = (1*cos th, 1*sin th);
// .conj(); 2.0f32 * c (component-wise); / points as f32 (component-wise).
static inline std::vector<Cf32> cisoid_matched_filter(size_t points, float freq_fs) {
  std::vector<Cf32> out(points);
  for (size_t iter = 0; iter < points; ++iter) {
    float th = 2.0f * 3.14159265358979323846f * freq_fs * (float)(points - 1 - iter);
    float r = expf(0.0f);
    float re = r * cosf(th);
    float im = -(r * sinf(th));
    out[iter].re = (2.0f * re) / (float)points;
    out[iter].im = (2.0f * im) / (float)points;
  }
  return out;
}

// waveform.rs:39-44
static inline void matched_filter(uint32_t fs, std::vector<Cf32>& mark, std::vector<Cf32>& space) {
  size_t ntaps = f32_as_usize(floorf(samples_per_symbol(fs)));
  mark = cisoid_matched_filter(ntaps, FSK_MARK_HZ / (float)fs);
  space = cisoid_matched_filter(ntaps, FSK_SPACE_HZ / (float)fs);
}

// ---------------------------------------------------------------------------------------
// filter.rs:218-323 — Window; filter.rs:363-377 — multiply_accumulate
// ---------------------------------------------------------------------------------------
template <typename T>
struct Window {
  std::deque<T> q;
  explicit Window(size_t len = 0) : q(len, T(0)) {}
  void reset() { for (auto& s : q) s = T(0); }
  size_t len() const { return q.size(); }
  // filter.rs:284-288
  T push_scalar(T v) {
    T out = T(0);
    if (!q.empty()) { out = q.front(); q.pop_front(); }
    q.push_back(v);
    return out;
  }
  // filter.rs:257-273
  void push(const T* in, size_t n) {
    if (n > q.size()) { in += n - q.size(); n = q.size(); }
    for (size_t i = 0; i < n; ++i) q.pop_front();
    for (size_t i = 0; i < n; ++i) q.push_back(in[i]);
  }
  T back() const { return q.back(); }
  T front() const { return q.front(); }
};

// filter.rs:363-377: newest history sample pairs with coeff[0]; sequential `out += hi * co`
static inline float mac_real(const std::deque<float>& hist, const std::vector<float>& coeff) {
  float out = 0.0f;
  size_t n = hist.size() < coeff.size() ? hist.size() : coeff.size();
  for (size_t i = 0; i < n; ++i) out += hist[hist.size() - 1 - i] * coeff[i];
  return out;
}
// f32 * Complex<f32> = (re*v, im*v) component-wise; Complex += component-wise (num-complex 0.4.6)
static inline Cf32 mac_cplx(const std::deque<float>& hist, const std::vector<Cf32>& coeff) {
  Cf32 out{0.0f, 0.0f};
  size_t n = hist.size() < coeff.size() ? hist.size() : coeff.size();
  for (size_t i = 0; i < n; ++i) {
    float v = hist[hist.size() - 1 - i];
    out.re += v * coeff[i].re;
    out.im += v * coeff[i].im;
  }
  return out;
}

// ---------------------------------------------------------------------------------------
// dcblock.rs
// ---------------------------------------------------------------------------------------
struct MovingAverage {  // dcblock.rs:62-109
  Window<float> window;
  float inv_len;
  float moving_sum = 0.0f;
  explicit MovingAverage(size_t len) : window(len), inv_len(1.0f / (float)len) {}
  void reset() { window.reset(); moving_sum = 0.0f; }
  size_t len() const { return window.len(); }
  // dcblock.rs:104-108
  void filter(float input, float& avg, float& delayed) {
    float aged = window.push_scalar(input);
    moving_sum += input - aged;
    avg = moving_sum * inv_len;
    delayed = window.front();
  }
};

struct DCBlocker {  // dcblock.rs:18-50
  MovingAverage ff, fb;
  explicit DCBlocker(size_t len) : ff(len), fb(len) {}
  void reset() { ff.reset(); fb.reset(); }
  float filter(float input) {  // dcblock.rs:45-49
    float ma0, sig, ma1, unused;
    ff.filter(input, ma0, sig);
    fb.filter(ma0, ma1, unused);
    return sig - ((ff.len() > 1) ? 1.0f : 0.0f) * ma1;
  }
};

// ---------------------------------------------------------------------------------------
// agc.rs
// ---------------------------------------------------------------------------------------
struct Agc {
  float bandwidth, min_gain, max_gain;
  bool locked = false;
  float gain;
  Agc(float bw, float gmin, float gmax)  // agc.rs:49-57
      : bandwidth(clampf(bw, 0.0f, 1.0f)), min_gain(gmin), max_gain(gmax), gain(fminf(1.0f, gmin)) {}
  void reset() { gain = 1.0f; locked = false; }  // agc.rs:60-63
  float input(float in) {                        // agc.rs:72-77
    float out = in * gain;
    gain += (locked ? 0.0f : 1.0f) * (1.0f - fabsf(out)) * bandwidth;
    gain = clampf(gain, min_gain, max_gain);
    return out;
  }
  void lock(bool l) { locked = l; }
};

// ---------------------------------------------------------------------------------------
// demod.rs
// ---------------------------------------------------------------------------------------
struct FskDemod {
  Window<float> window_input;
  std::vector<Cf32> coeff_mark, coeff_space;
  FskDemod(const std::vector<Cf32>& mark, const std::vector<Cf32>& space)
      : window_input(mark.size()), coeff_mark(mark), coeff_space(space) {}
  static FskDemod from_same(uint32_t fs) {  // demod.rs:129-132
    std::vector<Cf32> m, s;
    matched_filter(fs, m, s);
    return FskDemod(m, s);
  }
  size_t ntaps() const { return coeff_mark.size(); }
  void push_scalar(float v) { window_input.push_scalar(v); }
  void push(const float* v, size_t n) { window_input.push(v, n); }
  float demod() const {  // demod.rs:156-164
    Cf32 mark = mac_cplx(window_input.q, coeff_mark);
    Cf32 space = mac_cplx(window_input.q, coeff_space);
    return clampf(hypot_fixed(mark.re, mark.im) - hypot_fixed(space.re, space.im), -1.0f, 1.0f);
  }
  void reset() { window_input.reset(); }
};

// ---------------------------------------------------------------------------------------
// symsync.rs
// ---------------------------------------------------------------------------------------
struct SymbolEstimate { float data[2]; float err; };

static inline void compute_loop_alphabeta(float bw, float& alpha, float& beta) {  // symsync.rs:329-337
  float omega = 2.0f * 3.14159265358979323846f * bw;
  float k0 = 2.0f;
  float k1 = expf(-omega);
  float sh = sinhf(omega);
  alpha = k0 * k1 * sh;
  beta = k0 * (1.0f - k1 * (sh + 1.0f));
}

struct ZeroCrossingTed {  // symsync.rs:247-300
  float history[3] = {0, 0, 0};
  uint32_t sample_counter = 0;
  void reset() { history[0] = history[1] = history[2] = 0.0f; sample_counter = 0; }
  bool input(float sample, SymbolEstimate& out) {  // symsync.rs:278-287
    history[0] = history[1]; history[1] = history[2]; history[2] = sample;  // Wrapping ArrayDeque<_,3>
    sample_counter = (sample_counter + 1) % 2;
    if (sample_counter == 1) {
      float err = history[1] * (signumf(history[0]) - signumf(history[2]));  // symsync.rs:311-322
      out.data[0] = history[1]; out.data[1] = history[2]; out.err = err;
      return true;
    }
    return false;
  }
};

struct TimingLoop {  // symsync.rs:100-245
  float samples_per_ted, period_min, period_max, loop_alpha, loop_beta, period_avg, period_inst;
  ZeroCrossingTed ted;
  TimingLoop(float sps, float loop_bw, float max_dev) {  // symsync.rs:142-163
    compute_loop_alphabeta(loop_bw, loop_alpha, loop_beta);
    samples_per_ted = sps / 2.0f;
    float dev = sps * clampf(max_dev, 0.0f, 0.5f);
    period_avg = samples_per_ted;
    period_inst = samples_per_ted;
    period_min = period_avg - dev;
    period_max = period_avg + dev;
  }
  void reset() { ted.reset(); period_avg = samples_per_ted; period_inst = samples_per_ted; }  // :166-170
  void set_loop_bandwidth(float bw) { compute_loop_alphabeta(bw, loop_alpha, loop_beta); }     // :176-180
  // symsync.rs:198-201
  float input(float sample, float offset, bool& have, SymbolEstimate& sym) {
    have = ted.input(sample, sym);
    return advance_loop(offset, have ? &sym : nullptr);
  }
  float advance_loop(float offset, const SymbolEstimate* sym) {  // symsync.rs:219-244
    offset = clampf(offset, -0.5f, 0.5f);
    if (sym) {
      float err = clampf(sym->err - offset / samples_per_ted, -1.0f, 1.0f);
      period_avg += loop_beta * err;
      period_avg = clampf(period_avg, period_min, period_max);
      period_inst = period_avg + loop_alpha * err + offset;
      if (period_inst < 0.0f) period_inst = period_avg;
    } else {
      period_inst += offset;
    }
    return period_inst;
  }
};

// ---------------------------------------------------------------------------------------
// codesquelch.rs
// ---------------------------------------------------------------------------------------
enum class SquelchKind { NoCarrier, DroppedCarrier, Reading, Ready };
struct SquelchOut { float samples[16]; uint64_t symbol_counter; float power; };
struct SquelchState { SquelchKind kind; bool resync; SquelchOut out; };

struct CodeAndPowerSquelch {
  uint32_t max_errors;
  float power_open, power_close;
  uint32_t sync_to, data = 0;               // CodeCorrelator :404-435
  float pt_bandwidth, pt_power = 0.0f;      // PowerTracker :453-489
  std::deque<float> sample_history;         // ArrayDeque<f32,64,Wrapping>
  std::deque<bool> power_history;           // ArrayDeque<bool,32,Wrapping>
  uint64_t symbol_counter = 0;
  int sample_clock = -1;                    // Option<u8>: -1 = None
  bool sync_lock = false;

  CodeAndPowerSquelch(uint32_t sync, uint32_t maxerr, float popen, float pclose, float bw)  // :189-210
      : max_errors(maxerr), power_open(popen), power_close(fminf(pclose, popen)), sync_to(sync),
        pt_bandwidth(clampf(bw, 0.0f, 1.0f)) {}

  void end() { sync_lock = false; sample_clock = -1; }  // :336-339
  void reset() {                                         // :320-327
    end(); data = 0; sample_history.clear(); pt_power = 0.0f; power_history.clear(); symbol_counter = 0;
  }
  void lock(bool l) { sync_lock = l; }
  bool is_sync() const { return sample_clock >= 0; }
  uint64_t symbol_count() const { return symbol_counter; }
  float power() const { return pt_power; }

  SquelchState input(const float in[2]) {  // :228-304
    SquelchState st{};
    for (int i = 0; i < 2; ++i) {
      if (sample_history.size() == 64) sample_history.pop_front();
      sample_history.push_back(in[i]);
    }
    // CodeCorrelator::search :421-428
    uint32_t bit = (in[1] >= 0.0f) ? 1u : 0u;
    data = data >> 1;
    data |= bit << 31;
    uint32_t err = (uint32_t)__builtin_popcount(sync_to ^ data);  // :441-445
    // PowerTracker::track :483-488
    float p2 = in[1] * in[1];
    pt_power += (p2 - pt_power) * pt_bandwidth;
    pt_power = fmaxf(pt_power, 0.0f);
    float pwr = pt_power;
    if (power_history.size() == 32) power_history.pop_front();
    power_history.push_back(pwr >= power_close);
    symbol_counter += 1;

    if (sample_history.size() < 64) { st.kind = SquelchKind::NoCarrier; return st; }

    bool adjusted = false;
    if (!sync_lock && err <= max_errors && pwr >= power_open) {
      if (sample_clock < 0) { adjusted = true; sample_clock = 0; }
      else if (sample_clock == 0) { sample_clock = 0; }
      else { adjusted = true; sample_clock = 0; }
    } else if (is_sync() && !power_history.front()) {
      end();
      st.kind = SquelchKind::DroppedCarrier;
      return st;
    }

    if (sample_clock < 0) { st.kind = SquelchKind::NoCarrier; return st; }
    if (sample_clock == 0) {
      sample_clock = 1;
      for (int i = 0; i < 16; ++i) st.out.samples[i] = sample_history[i];
      st.out.symbol_counter = symbol_counter;
      st.out.power = pwr;
      st.kind = SquelchKind::Ready;
      st.resync = adjusted;
      return st;
    }
    sample_clock = (sample_clock + 1) % 8;
    st.kind = SquelchKind::Reading;
    return st;
  }
};

// ---------------------------------------------------------------------------------------
// equalize.rs
// ---------------------------------------------------------------------------------------
struct Equalizer {
  enum Mode { Disabled, EnabledFeedback, EnabledTraining };
  float relaxation, regularization;
  bool have_train; uint32_t train_to;
  std::vector<float> ff_coeff, fb_coeff;
  Window<float> ff_wind, fb_wind;
  Mode mode = EnabledFeedback;
  uint32_t train_sa = 0, train_count = 0;

  Equalizer(size_t nff, size_t nfb, float relax, float regul, bool have_t, uint32_t t)  // :124-151
      : relaxation(relax), regularization(regul), have_train(have_t), train_to(t),
        ff_coeff(nff, 0.0f), fb_coeff(nfb, 0.0f), ff_wind(nff), fb_wind(nfb) {
    ff_coeff[0] = 1.0f; fb_coeff[0] = 1.0f;
  }
  void reset() {  // :191-196 (mode preserved)
    for (auto& c : ff_coeff) c = 0.0f; ff_coeff[0] = 1.0f;
    for (auto& c : fb_coeff) c = 0.0f; fb_coeff[0] = 1.0f;
    ff_wind.reset(); fb_wind.reset();
  }
  void enable(bool e) { mode = e ? EnabledFeedback : Disabled; }
  bool train() {  // :216-220
    if (!have_train) return false;
    mode = EnabledTraining; train_sa = train_to; train_count = 0;
    return true;
  }
  bool is_training() const { return mode == EnabledTraining; }

  static float nlms_gain(float relax, float regul, const std::deque<float>& w) {  // :376-386 (oldest first)
    float sumsq = 0.0f;
    for (float v : w) sumsq += v * v;
    return relax / (regul + sumsq);
  }
  static void nlms_update(float relax, float regul, float error, const std::deque<float>& w,
                          std::vector<float>& filt) {  // :354-364
    float gain = nlms_gain(relax, regul, w);
    size_t n = filt.size() < w.size() ? filt.size() : w.size();
    for (size_t i = 0; i < n; ++i) filt[i] += gain * error * w[w.size() - 1 - i];
  }
  void evolve(float error) {  // :315-332
    nlms_update(relaxation, regularization, error, ff_wind.q, ff_coeff);
    nlms_update(relaxation, regularization, -error, fb_wind.q, fb_coeff);
  }
  // :249-308
  bool estimate_symbol(const float in[2], float& err_out) {
    ff_wind.push(in, 2);
    float ff = mac_real(ff_wind.q, ff_coeff);
    float fb = mac_real(fb_wind.q, fb_coeff);
    float sym_val = ff - fb;
    float sym_est, err;
    switch (mode) {
      case Disabled:
        sym_est = signumf(sym_val); err = 0.0f; break;
      case EnabledFeedback:
        sym_est = signumf(sym_val); err = sym_est - sym_val; evolve(err); break;
      default: {
        sym_est = (2.0f * (float)(train_sa & 1u)) - 1.0f;
        train_sa >>= 1;
        err = sym_est - sym_val;
        evolve(err);
        train_count += 1;
        if (train_count >= 32) mode = EnabledFeedback;
        break;
      }
    }
    float fbpush[2] = {sym_est, 0.0f};
    fb_wind.push(fbpush, 2);
    err_out = err;
    return sym_est >= 0.0f;
  }
  uint8_t input(const float samples[16], float& last_err) {  // :173-186
    uint8_t byte = 0; last_err = 0.0f;
    for (int b = 0; b < 8; ++b) {
      float e; bool bit = estimate_symbol(samples + 2 * b, e);
      last_err = e;
      byte |= (uint8_t)((bit ? 1 : 0) << b);
    }
    return byte;
  }
};

// ---------------------------------------------------------------------------------------
// combiner.rs:105-137
// ---------------------------------------------------------------------------------------
static inline bool is_allowed_byte(uint8_t c) {
  return c == '-' || (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') ||
         c == '/' || c == '?' || c == '(' || c == ')' || c == '[' || c == ']' || c == '.' ||
         c == '_' || c == ',' || c == '+' || c == ' ';
}

// ---------------------------------------------------------------------------------------
// output.rs — link state
// ---------------------------------------------------------------------------------------
enum class LinkKind : uint32_t { NoCarrier = 0, Searching = 1, Reading = 2, Burst = 3 };
struct LinkState {
  LinkKind kind = LinkKind::NoCarrier;
  std::vector<uint8_t> burst;
  bool operator==(const LinkState& o) const { return kind == o.kind && burst == o.burst; }
  bool operator!=(const LinkState& o) const { return !(*this == o); }
};

// ---------------------------------------------------------------------------------------
// framing.rs
// ---------------------------------------------------------------------------------------
static inline uint32_t message_prefix_errors(uint32_t inp) {  // framing.rs:235-243
  const uint32_t S = 0x5A435A43u;  // "ZCZC"
  const uint32_t E = 0x4E4E4E4Eu;  // "NNNN"
  uint32_t a = (uint32_t)__builtin_popcount(inp ^ S), b = (uint32_t)__builtin_popcount(inp ^ E);
  return a < b ? a : b;
}

struct Framer {
  enum St { Idle, PrefixSearch, DataRead };
  St st = Idle;
  uint32_t search_word = 0, count = 0;
  std::vector<uint8_t> msg; uint32_t invalid = 0;
  uint32_t max_prefix_bit_errors, max_invalid_bytes;
  static constexpr uint32_t PREFIX_SEARCH_LEN = 21;  // framing.rs:201
  Framer(uint32_t mpe, uint32_t mib) : max_prefix_bit_errors(mpe), max_invalid_bytes(mib) {}
  void reset() { st = Idle; msg.clear(); }
  LinkState state() const {  // framing.rs:191-197
    LinkState l;
    l.kind = st == Idle ? LinkKind::NoCarrier : (st == PrefixSearch ? LinkKind::Searching : LinkKind::Reading);
    return l;
  }
  LinkState end() {  // framing.rs:174-186
    LinkState l;
    if (st == DataRead) { l.kind = LinkKind::Burst; l.burst = std::move(msg); msg.clear(); }
    else l.kind = LinkKind::NoCarrier;
    st = Idle;
    return l;
  }
  LinkState input(uint8_t data, uint64_t symbol_count, bool restart) {  // framing.rs:109-164
    if (restart) {
      LinkState out = end();
      st = PrefixSearch; search_word = 0; count = 0;
      (void)input(data, symbol_count, false);
      if (out.kind == LinkKind::Burst) return out;
      LinkState l; l.kind = LinkKind::Searching; return l;
    }
    switch (st) {
      case Idle: { LinkState l; return l; }
      case PrefixSearch: {
        search_word = (search_word << 8) | (uint32_t)data;
        count += 1;
        if (message_prefix_errors(search_word) <= max_prefix_bit_errors) {
          msg.clear();
          msg.push_back((uint8_t)(search_word >> 24)); msg.push_back((uint8_t)(search_word >> 16));
          msg.push_back((uint8_t)(search_word >> 8));  msg.push_back((uint8_t)(search_word));
          invalid = 0;
          st = DataRead;
        } else if (count > PREFIX_SEARCH_LEN) {
          st = Idle;
        }
        return state();
      }
      default: {
        invalid += is_allowed_byte(data) ? 0u : 1u;
        if (invalid > max_invalid_bytes) return end();
        msg.push_back(data);
        return state();
      }
    }
  }
};

// ---------------------------------------------------------------------------------------
// sameplace message.rs:718-736, 801-828 — the part of Message parsing the receiver needs
// ---------------------------------------------------------------------------------------
enum class DecodeErr : uint32_t { None = 0, UnrecognizedPrefix = 1, NotAscii = 2, Malformed = 3 };

struct Message {
  bool is_som = false;        // StartOfMessage(header) vs EndOfMessage
  std::string text;           // header text (truncated at hdr_length) or "NNNN"
  size_t offset_time = 0;
  size_t parity_error_count = 0, voting_byte_count = 0;
  bool operator==(const Message& o) const {
    return is_som == o.is_som && text == o.text && offset_time == o.offset_time &&
           parity_error_count == o.parity_error_count && voting_byte_count == o.voting_byte_count;
  }
  const std::string& as_str() const { return text; }
};

// message.rs:813-828: ^ZCZC-[[:alpha:]]{3}-[[:alpha:]]{3}(-[0-9]{6})+(\+[0-9]{4}-[0-9]{7}-.{3,8}-)
// (leftmost-first: each greedy repetition prefers the longest, then backtracks)
static inline bool check_header(const std::string& h, size_t& off_time, size_t& hdr_len) {
  auto alpha = [](char c) { return (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'); };
  auto digit = [](char c) { return c >= '0' && c <= '9'; };
  size_t n = h.size(), p = 0;
  if (n < 5 || h.compare(0, 5, "ZCZC-") != 0) return false;
  p = 5;
  for (int k = 0; k < 3; ++k, ++p) if (p >= n || !alpha(h[p])) return false;
  if (p >= n || h[p] != '-') return false; ++p;
  for (int k = 0; k < 3; ++k, ++p) if (p >= n || !alpha(h[p])) return false;
  // greedy (-dddddd)+ : collect all possible end positions, try longest first
  std::vector<size_t> ends;
  size_t q = p;
  while (true) {
    if (q + 7 > n || h[q] != '-') break;
    bool ok = true;
    for (int k = 1; k <= 6; ++k) if (!digit(h[q + k])) { ok = false; break; }
    if (!ok) break;
    q += 7; ends.push_back(q);
  }
  for (size_t ei = ends.size(); ei-- > 0;) {
    size_t s = ends[ei];  // position of '+'
    size_t r = s;
    if (r >= n || h[r] != '+') continue; ++r;
    bool ok = true;
    for (int k = 0; k < 4; ++k, ++r) if (r >= n || !digit(h[r])) { ok = false; break; }
    if (!ok) continue;
    if (r >= n || h[r] != '-') continue; ++r;
    for (int k = 0; k < 7; ++k, ++r) if (r >= n || !digit(h[r])) { ok = false; break; }
    if (!ok) continue;
    if (r >= n || h[r] != '-') continue; ++r;
    for (int len = 8; len >= 3; --len) {  // .{3,8}- greedy; '.' excludes '\n'
      if (r + (size_t)len >= n) continue;
      bool nl = false;
      for (int k = 0; k < len; ++k) if (h[r + k] == '\n') { nl = true; break; }
      if (nl) continue;
      if (h[r + len] == '-') { off_time = s; hdr_len = r + len + 1; return true; }
    }
  }
  return false;
}

// Message::try_from((&[u8], &[u8], &[u8])) message.rs:718-736 with MessageHeader::new_with_error_info :235-254
static inline DecodeErr message_try_from(const uint8_t* bytes, size_t n, const uint8_t* errs, size_t nerr,
                                         const uint8_t* bursts, size_t nb, Message& out) {
  for (size_t i = 0; i < n; ++i) if (bytes[i] & 0x80) return DecodeErr::NotAscii;  // from_utf8 failure (inputs are 7-bit)
  std::string s((const char*)bytes, n);
  if (s.compare(0, 5, "ZCZC-") == 0 && s.size() >= 5) {
    size_t off, len;
    if (!check_header(s, off, len)) return DecodeErr::Malformed;
    out.is_som = true;
    out.text = s.substr(0, len);
    out.offset_time = off;
    size_t pe = 0, vc = 0;
    for (size_t i = 0; i < nerr && i < len; ++i) pe += errs[i];
    for (size_t i = 0; i < nb && i < len; ++i) vc += (bursts[i] >= 3) ? 1 : 0;
    out.parity_error_count = pe; out.voting_byte_count = vc;
    return DecodeErr::None;
  } else if (s.size() >= 2 && s.compare(0, 2, "NN") == 0) {
    out = Message{}; out.is_som = false; out.text = "NNNN";
    return DecodeErr::None;
  }
  return DecodeErr::UnrecognizedPrefix;
}

struct MessageResult {
  bool ok = false; Message msg; DecodeErr err = DecodeErr::None;
  bool operator==(const MessageResult& o) const {
    return ok == o.ok && (ok ? (msg == o.msg) : (err == o.err));
  }
};

// ---------------------------------------------------------------------------------------
// combiner.rs:32-271
// ---------------------------------------------------------------------------------------
constexpr size_t MAX_MESSAGE_LENGTH = 268;  // assembler.rs:70

static inline void bit_vote_detect(uint8_t b0, uint8_t b1, uint8_t& out, uint32_t& nerr) {  // combiner.rs:216-222
  uint8_t x = b0 ^ b1;
  out = (uint8_t)(b0 & (uint8_t)~(uint8_t)(0xff * (x != 0 ? 1 : 0)));
  nerr = (uint32_t)__builtin_popcount(x);
}
static inline void bit_vote_correct(uint8_t b0, uint8_t b1, uint8_t b2, uint8_t& out, uint32_t& nerr) {  // :234-249
  uint8_t p0 = (uint8_t)~(b0 ^ b1), p1 = (uint8_t)~(b1 ^ b2), p2 = (uint8_t)~(b0 ^ b2);
  out = (uint8_t)((b0 & p0) | (b2 & p1) | (b2 & p2));
  nerr = 8u - (uint32_t)__builtin_popcount((uint8_t)(p0 & p1 & p2));
}

struct Estimate { std::vector<uint8_t> bytes, nbursts, errs; };

static inline Estimate estimate_message(const std::vector<const std::vector<uint8_t>*>& bursts) {  // combiner.rs:154-203
  Estimate e;
  size_t nb = bursts.size() < 3 ? bursts.size() : 3;
  size_t pos[3] = {0, 0, 0};
  while (e.bytes.size() < MAX_MESSAGE_LENGTH) {
    uint8_t cur[3]; size_t nc = 0;
    for (size_t k = 0; k < nb; ++k)
      if (pos[k] < bursts[k]->size()) cur[nc++] = (*bursts[k])[pos[k]++];
    bool msb = false;
    for (size_t k = 0; k < nc; ++k) { msb |= (cur[k] & 0x80) != 0; cur[k] &= 0x7f; }
    uint8_t est = 0; uint32_t be = 0;
    if (nc == 0) break;
    else if (nc == 1) { est = cur[0]; be = 0; }
    else if (nc == 2) bit_vote_detect(cur[0], cur[1], est, be);
    else bit_vote_correct(cur[0], cur[1], cur[2], est, be);
    if (!is_allowed_byte(est)) break;
    e.bytes.push_back(est);
    e.nbursts.push_back((uint8_t)nc);
    e.errs.push_back((uint8_t)(be + (msb ? 1 : 0)));
  }
  return e;
}

// combiner.rs:32-80.  Returns false for `None`.
static inline bool combine(const std::vector<const std::vector<uint8_t>*>& bursts, MessageResult& res) {
  Estimate e = estimate_message(bursts);
  if (e.bytes.empty()) return false;
  size_t good = 0;  // truncate_bytes_with_reference(.., 2) combiner.rs:262-271
  for (size_t i = 0; i < e.bytes.size() && i < e.nbursts.size(); ++i) { if (e.nbursts[i] < 2) break; ++good; }
  Message m;
  DecodeErr err = message_try_from(e.bytes.data(), good, e.errs.data(), e.errs.size(), e.nbursts.data(),
                                   e.nbursts.size(), m);
  if (err == DecodeErr::None) { res.ok = true; res.msg = m; res.err = DecodeErr::None; return true; }
  if (e.bytes.size() >= 2 && e.bytes[0] == 'N' && e.bytes[1] == 'N') {  // Fast EOM :251-258
    res.ok = true; res.msg = Message{}; res.msg.text = "NNNN"; res.err = DecodeErr::None; return true;
  }
  if (good == 0) return false;
  res.ok = false; res.err = err; res.msg = Message{};
  return true;
}

// ---------------------------------------------------------------------------------------
// assembler.rs
// ---------------------------------------------------------------------------------------
enum class TransportKind : uint32_t { Idle = 0, Assembling = 1, Message = 2 };
struct TransportState {
  TransportKind kind = TransportKind::Idle;
  MessageResult res;
  bool operator==(const TransportState& o) const {
    return kind == o.kind && (kind != TransportKind::Message || res == o.res);
  }
  bool operator!=(const TransportState& o) const { return !(*this == o); }
};

static inline uint64_t max_interburst_symbols() {  // assembler.rs:85, evaluated in f32
  float v = (1.05f * BAUD_HZ) + 17.0f * 8.0f;
  return (uint64_t)v;
}
static inline uint64_t max_history_duration() {  // assembler.rs:92-93
  return 2 * (max_interburst_symbols() + 8 * (uint64_t)MAX_MESSAGE_LENGTH);
}

struct Assembler {
  struct TimedBurst { std::vector<uint8_t> data; uint64_t deadline; };
  std::deque<TimedBurst> history;
  bool pending = false; MessageResult pending_res; uint64_t pending_deadline = 0;
  bool have_prev = false; Message prev; uint64_t prev_deadline = 0;

  void reset() { history.clear(); pending = false; have_prev = false; }

  void prune_history(uint64_t now) {  // assembler.rs:357-363
    for (auto it = history.begin(); it != history.end();) {
      if (it->deadline <= now) it = history.erase(it); else ++it;
    }
    while (history.size() > 2) history.pop_front();
  }
  // PendingResult::accept assembler.rs:294-331
  bool accept(const MessageResult& msg, uint64_t now) {
    uint64_t dl = (msg.ok && !msg.msg.is_som) ? now : now + max_interburst_symbols();
    if (pending) {
      bool replace;
      if (!pending_res.ok) replace = true;
      else if (!pending_res.msg.is_som && msg.ok && msg.msg.is_som) replace = true;
      else if (pending_res.msg.is_som && msg.ok && msg.msg.is_som)
        replace = msg.msg.voting_byte_count >= pending_res.msg.voting_byte_count;
      else replace = false;
      if (replace) { pending_res = msg; pending_deadline = dl; return true; }
      return false;
    }
    pending = true; pending_res = msg; pending_deadline = dl;
    return true;
  }
  TransportState idle(uint64_t now) {  // assembler.rs:205-234
    prune_history(now);
    TransportState ts;
    if (pending && pending_deadline <= now) {  // poll :339-348
      MessageResult r = pending_res; pending = false;
      if (r.ok) { have_prev = true; prev = r.msg; prev_deadline = now + max_history_duration(); }
      ts.kind = TransportKind::Message; ts.res = r;
      return ts;
    }
    ts.kind = history.empty() ? TransportKind::Idle : TransportKind::Assembling;
    return ts;
  }
  TransportState assemble(const std::vector<uint8_t>& burst, uint64_t now) {  // assembler.rs:154-184
    if (burst.empty()) return idle(now);
    prune_history(now);
    if (have_prev && prev_deadline <= now) have_prev = false;  // prune_previous :366-371
    TimedBurst tb;
    size_t n = burst.size() < MAX_MESSAGE_LENGTH ? burst.size() : MAX_MESSAGE_LENGTH;
    tb.data.assign(burst.begin(), burst.begin() + n);
    tb.deadline = now + max_history_duration();
    history.push_back(std::move(tb));
    std::vector<const std::vector<uint8_t>*> bs;
    for (auto& h : history) bs.push_back(&h.data);
    MessageResult r;
    if (combine(bs, r)) {
      bool keep = true;  // deduplicate :245-265
      if (r.ok && have_prev && prev.as_str() == r.msg.as_str()) keep = false;
      if (keep) accept(r, now);
    }
    return idle(now);
  }
};

// ---------------------------------------------------------------------------------------
// builder.rs — configuration (defaults :50-67, :369-376) with the clamping setters
// ---------------------------------------------------------------------------------------
struct Config {
  uint32_t input_rate = 22050;
  float dc_blocker_len = 0.38f;
  float agc_bandwidth = 0.01f;
  float agc_gain_min = 0.0f, agc_gain_max = 1.0e6f;
  float timing_bw_unlocked = 0.125f, timing_bw_locked = 0.05f;
  float timing_max_deviation = 0.01f;
  float squelch_power_open = 0.10f, squelch_power_close = 0.05f;
  float squelch_bandwidth = 0.125f;
  uint32_t preamble_max_errors = 2;
  uint32_t eq_enabled = 1;
  uint32_t eq_nff = 6, eq_nfb = 4;
  float eq_relaxation = 0.05f, eq_regularization = 1.0e-6f;
  uint32_t frame_prefix_max_errors = 2, frame_max_invalid_bytes = 5;

  // crates/samedec/src/main.rs:29-37 with cli.rs defaults
  static Config samedec(uint32_t rate = 22050) {
    Config c; c.input_rate = rate;
    c.agc_gain_min = 1.0f / 32767.0f; c.agc_gain_max = 1.0f / 200.0f;
    return c;
  }
};

// ---------------------------------------------------------------------------------------
// receiver.rs — SameReceiver
// ---------------------------------------------------------------------------------------
struct Event {
  bool is_transport = false;
  LinkState link;
  TransportState transport;
  uint64_t input_sample_counter = 0;
  uint64_t symbol_count = 0;  // squelch.symbol_count() when the event was queued (debug aid; not in the reference event)
};

struct SoftSym { uint64_t sample; float zero, sym; };

struct SameReceiver {
  DCBlocker dc_block;
  Agc agc;
  FskDemod demod;
  TimingLoop symsync;
  CodeAndPowerSquelch squelch;
  Equalizer equalizer;
  Framer framer;
  Assembler assembler;
  float timing_bw_unlocked, timing_bw_locked;
  uint32_t input_rate;
  uint64_t input_sample_counter = 0;
  LinkState link_state;
  TransportState transport_state;
  std::deque<Event> event_queue;
  uint32_t ted_sample_clock = 0;
  float samples_until_next_ted;
  bool have_force_eom = false; uint64_t force_eom_at_sample = 0;
  std::vector<SoftSym>* trace = nullptr;

  static constexpr uint64_t MAX_MESSAGE_DURATION_SECS = 135;  // receiver.rs:496

  explicit SameReceiver(const Config& c)  // receiver.rs:502-560
      : dc_block(f32_as_usize(c.dc_blocker_len * samples_per_symbol(c.input_rate))),
        agc(c.agc_bandwidth * samples_per_symbol(c.input_rate) / (float)c.input_rate, c.agc_gain_min, c.agc_gain_max),
        demod(FskDemod::from_same(c.input_rate)),
        symsync(samples_per_symbol(c.input_rate), c.timing_bw_unlocked, c.timing_max_deviation),
        squelch(PREAMBLE_SYNC_WORD, c.preamble_max_errors, c.squelch_power_open, c.squelch_power_close,
                c.squelch_bandwidth),
        equalizer(c.eq_enabled ? c.eq_nff : 1, c.eq_enabled ? c.eq_nfb : 1,  // disabled_equalizer() :585-590
                  c.eq_enabled ? c.eq_relaxation : 0.0f, c.eq_regularization, true, PREAMBLE_SYNC_WORD),
        framer(c.frame_prefix_max_errors, c.frame_max_invalid_bytes),
        timing_bw_unlocked(c.timing_bw_unlocked), timing_bw_locked(c.timing_bw_locked),
        input_rate(c.input_rate), samples_until_next_ted(symsync.samples_per_ted) {}

  void reset() {  // receiver.rs:182-198
    dc_block.reset(); agc.reset(); demod.reset(); symsync.reset(); squelch.reset(); equalizer.reset();
    framer.reset(); assembler.reset();
    input_sample_counter = 0; link_state = LinkState{}; transport_state = TransportState{};
    event_queue.clear(); ted_sample_clock = 0; samples_until_next_ted = symsync.samples_per_ted;
    have_force_eom = false;
  }

  void end() {  // receiver.rs:479-490
    agc.lock(false); squelch.end(); equalizer.reset();
    symsync.set_loop_bandwidth(timing_bw_unlocked); symsync.reset();
  }

  LinkState process_linklayer_symbol(const SymbolEstimate& symbol) {  // receiver.rs:407-474
    SquelchState sq = squelch.input(symbol.data);
    bool is_resync;
    switch (sq.kind) {
      case SquelchKind::NoCarrier: return framer.end();
      case SquelchKind::DroppedCarrier: end(); return framer.end();
      case SquelchKind::Reading: return framer.state();
      default:
        if (sq.resync) {
          agc.lock(true);
          symsync.set_loop_bandwidth(timing_bw_locked);
          equalizer.train();
          is_resync = true;
        } else is_resync = false;
    }
    float adaptive_err;
    uint8_t byte_est = equalizer.input(sq.out.samples, adaptive_err);
    LinkState ls = framer.input(byte_est, sq.out.symbol_counter, is_resync);
    if (ls.kind == LinkKind::Reading) squelch.lock(true);
    else if (ls.kind == LinkKind::NoCarrier || ls.kind == LinkKind::Burst) end();
    return ls;
  }

  // receiver.rs:343-395; returns true if a link state was produced
  bool process_linklayer_high_rate(float input, LinkState& out) {
    float sa = agc.input(dc_block.filter(input));
    demod.push_scalar(sa);
    ted_sample_clock += 1;
    input_sample_counter += 1;
    float clock_remaining_sa = samples_until_next_ted - (float)ted_sample_clock;
    if (clock_remaining_sa <= 0.0f || fabsf(clock_remaining_sa) < 0.5f) {
      ted_sample_clock = 0;
      float sa_low = demod.demod();
      bool have; SymbolEstimate sym;
      samples_until_next_ted = symsync.input(sa_low, clock_remaining_sa, have, sym);
      if (!have) return false;
      if (trace) trace->push_back(SoftSym{input_sample_counter, sym.data[0], sym.data[1]});
      out = process_linklayer_symbol(sym);
      return true;
    }
    return false;
  }

  bool process_transportlayer(const LinkState& ls, TransportState& out) {  // receiver.rs:291-333
    if (ls.kind == LinkKind::Burst) {
      out = assembler.assemble(ls.burst, squelch.symbol_count());
    } else if (ls.kind == LinkKind::NoCarrier && have_force_eom && input_sample_counter > force_eom_at_sample) {
      out = TransportState{}; out.kind = TransportKind::Message; out.res.ok = true; out.res.msg.text = "NNNN";
    } else if (ls.kind == LinkKind::NoCarrier) {
      out = assembler.idle(squelch.symbol_count());
    } else return false;
    if (out.kind == TransportKind::Message && out.res.ok) {
      if (out.res.msg.is_som) {
        have_force_eom = true;
        force_eom_at_sample = input_sample_counter + MAX_MESSAGE_DURATION_SECS * (uint64_t)input_rate;
      } else have_force_eom = false;
    }
    return true;
  }

  // One sample through receiver.rs:243-270 (without the early return: all queued events are appended to `sink`)
  void process_sample(float sample, std::vector<Event>& sink) {
    LinkState ls;
    if (process_linklayer_high_rate(sample, ls)) {
      if (ls != link_state) {
        link_state = ls;
        Event e; e.is_transport = false; e.link = link_state; e.input_sample_counter = input_sample_counter;
        e.symbol_count = squelch.symbol_count();
        sink.push_back(std::move(e));
      }
      TransportState ts;
      if (process_transportlayer(ls, ts) && ts != transport_state) {
        transport_state = ts;
        Event e; e.is_transport = true; e.transport = transport_state; e.input_sample_counter = input_sample_counter;
        e.symbol_count = squelch.symbol_count();
        sink.push_back(std::move(e));
      }
    }
  }
};

// crates/samedec/src/app.rs:71-74,103-119 + receiver.rs:216-224: after EOF, flush() = up to 4 s of zeros,
// abandoned at the first message; repeated until a full 4 s of zeros yields no message.
static inline void samedec_eof_flush(SameReceiver& rx, std::vector<Event>& sink) {
  const uint64_t nflush = (uint64_t)rx.input_rate * 4;
  while (true) {
    bool got = false;
    for (uint64_t i = 0; i < nflush && !got; ++i) {
      size_t before = sink.size();
      rx.process_sample(0.0f, sink);
      for (size_t k = before; k < sink.size(); ++k)
        if (sink[k].is_transport && sink[k].transport.kind == TransportKind::Message && sink[k].transport.res.ok)
          got = true;
    }
    if (!got) break;
  }
}

}  // namespace same_oracle
