// same_oracle_capi.cpp — C ABI over the CPU oracle (TEST INFRASTRUCTURE ONLY; see same_oracle.hpp).
// Loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
#include "same_oracle.hpp"
#include "synth_cpu.hpp"

#include <atomic>
#include <chrono>
#include <thread>

using namespace same_oracle;

extern "C" {

// Mirrors SameReceiverBuilder (builder.rs:50-67, 369-376); same field order as include/same_engine.h:same_config.
typedef struct oracle_config {
  uint32_t input_rate;
  float dc_blocker_len;
  float agc_bandwidth;
  float agc_gain_min, agc_gain_max;
  float timing_bw_unlocked, timing_bw_locked;
  float timing_max_deviation;
  float squelch_power_open, squelch_power_close;
  float squelch_bandwidth;
  uint32_t preamble_max_errors;
  uint32_t eq_enabled;
  uint32_t eq_nff, eq_nfb;
  float eq_relaxation, eq_regularization;
  uint32_t frame_prefix_max_errors, frame_max_invalid_bytes;
} oracle_config;

enum {
  OR_EV_LINK_NOCARRIER = 0, OR_EV_LINK_SEARCHING = 1, OR_EV_LINK_READING = 2, OR_EV_LINK_BURST = 3,
  OR_EV_TR_IDLE = 16, OR_EV_TR_ASSEMBLING = 17, OR_EV_TR_MSG_SOM = 18, OR_EV_TR_MSG_EOM = 19, OR_EV_TR_MSG_ERR = 20
};

typedef struct oracle_event {
  uint32_t kind;
  uint32_t err;
  uint64_t input_sample_counter;
  uint64_t symbol_count;
  uint32_t data_len;
  uint32_t parity_errors;
  uint32_t voting_bytes;
  uint32_t reserved;
} oracle_event;

typedef struct oracle_soft { uint64_t sample; float zero, sym; } oracle_soft;

struct OracleHandle {
  SameReceiver rx;
  std::vector<Event> events;
  std::vector<SoftSym> trace;
  explicit OracleHandle(const Config& c) : rx(c) {}
};

static Config to_config(const oracle_config* c) {
  Config k;
  k.input_rate = c->input_rate; k.dc_blocker_len = c->dc_blocker_len; k.agc_bandwidth = c->agc_bandwidth;
  k.agc_gain_min = c->agc_gain_min; k.agc_gain_max = c->agc_gain_max;
  k.timing_bw_unlocked = c->timing_bw_unlocked; k.timing_bw_locked = c->timing_bw_locked;
  k.timing_max_deviation = c->timing_max_deviation;
  k.squelch_power_open = c->squelch_power_open; k.squelch_power_close = c->squelch_power_close;
  k.squelch_bandwidth = c->squelch_bandwidth; k.preamble_max_errors = c->preamble_max_errors;
  k.eq_enabled = c->eq_enabled; k.eq_nff = c->eq_nff; k.eq_nfb = c->eq_nfb;
  k.eq_relaxation = c->eq_relaxation; k.eq_regularization = c->eq_enabled ? c->eq_regularization : 1.0e-6f;
  k.frame_prefix_max_errors = c->frame_prefix_max_errors; k.frame_max_invalid_bytes = c->frame_max_invalid_bytes;
  return k;
}

void oracle_default_config(oracle_config* c, uint32_t rate, int samedec) {
  Config k = samedec ? Config::samedec(rate) : Config();
  k.input_rate = rate;
  c->input_rate = k.input_rate; c->dc_blocker_len = k.dc_blocker_len; c->agc_bandwidth = k.agc_bandwidth;
  c->agc_gain_min = k.agc_gain_min; c->agc_gain_max = k.agc_gain_max;
  c->timing_bw_unlocked = k.timing_bw_unlocked; c->timing_bw_locked = k.timing_bw_locked;
  c->timing_max_deviation = k.timing_max_deviation;
  c->squelch_power_open = k.squelch_power_open; c->squelch_power_close = k.squelch_power_close;
  c->squelch_bandwidth = k.squelch_bandwidth; c->preamble_max_errors = k.preamble_max_errors;
  c->eq_enabled = k.eq_enabled; c->eq_nff = k.eq_nff; c->eq_nfb = k.eq_nfb;
  c->eq_relaxation = k.eq_relaxation; c->eq_regularization = k.eq_regularization;
  c->frame_prefix_max_errors = k.frame_prefix_max_errors; c->frame_max_invalid_bytes = k.frame_max_invalid_bytes;
}

void* oracle_create(const oracle_config* c) {
  Config k = to_config(c);
  if (f32_as_usize(k.dc_blocker_len * samples_per_symbol(k.input_rate)) == 0) return nullptr;  // MovingAverage::new asserts len>0
  if (f32_as_usize(floorf(samples_per_symbol(k.input_rate))) == 0) return nullptr;
  return new OracleHandle(k);
}
void oracle_destroy(void* h) { delete (OracleHandle*)h; }
void oracle_reset(void* h) { auto* o = (OracleHandle*)h; o->rx.reset(); o->events.clear(); o->trace.clear(); }
void oracle_enable_trace(void* h, int on) { auto* o = (OracleHandle*)h; o->rx.trace = on ? &o->trace : nullptr; }

void oracle_process_s16(void* h, const int16_t* s, size_t n) {
  auto* o = (OracleHandle*)h;
  for (size_t i = 0; i < n; ++i) o->rx.process_sample((float)s[i], o->events);  // `sa as f32` app.rs:112
}
void oracle_process_f32(void* h, const float* s, size_t n) {
  auto* o = (OracleHandle*)h;
  for (size_t i = 0; i < n; ++i) o->rx.process_sample(s[i], o->events);
}
void oracle_process_zeros(void* h, size_t n) {
  auto* o = (OracleHandle*)h;
  for (size_t i = 0; i < n; ++i) o->rx.process_sample(0.0f, o->events);
}
void oracle_flush_samedec(void* h) { auto* o = (OracleHandle*)h; samedec_eof_flush(o->rx, o->events); }

uint64_t oracle_input_sample_counter(void* h) { return ((OracleHandle*)h)->rx.input_sample_counter; }
size_t oracle_num_events(void* h) { return ((OracleHandle*)h)->events.size(); }

static const std::string& ev_text(const Event& e, std::string& tmp) {
  if (!e.is_transport) { tmp.assign((const char*)e.link.burst.data(), e.link.burst.size()); return tmp; }
  if (e.transport.kind == TransportKind::Message && e.transport.res.ok) return e.transport.res.msg.text;
  tmp.clear(); return tmp;
}

// Returns 0 on success; copies up to `cap` payload bytes (burst bytes or message text) into `data`.
int oracle_get_event(void* h, size_t i, oracle_event* out, uint8_t* data, size_t cap) {
  auto* o = (OracleHandle*)h;
  if (i >= o->events.size()) return -1;
  const Event& e = o->events[i];
  memset(out, 0, sizeof(*out));
  out->input_sample_counter = e.input_sample_counter;
  out->symbol_count = e.symbol_count;
  if (!e.is_transport) {
    out->kind = (uint32_t)e.link.kind;
  } else if (e.transport.kind == TransportKind::Idle) out->kind = OR_EV_TR_IDLE;
  else if (e.transport.kind == TransportKind::Assembling) out->kind = OR_EV_TR_ASSEMBLING;
  else if (!e.transport.res.ok) { out->kind = OR_EV_TR_MSG_ERR; out->err = (uint32_t)e.transport.res.err; }
  else if (e.transport.res.msg.is_som) {
    out->kind = OR_EV_TR_MSG_SOM;
    out->parity_errors = (uint32_t)e.transport.res.msg.parity_error_count;
    out->voting_bytes = (uint32_t)e.transport.res.msg.voting_byte_count;
  } else out->kind = OR_EV_TR_MSG_EOM;
  std::string tmp;
  const std::string& t = ev_text(e, tmp);
  out->data_len = (uint32_t)t.size();
  if (data && cap) memcpy(data, t.data(), t.size() < cap ? t.size() : cap);
  return 0;
}

size_t oracle_trace_len(void* h) { return ((OracleHandle*)h)->trace.size(); }
void oracle_get_trace(void* h, oracle_soft* out, size_t cap) {
  auto* o = (OracleHandle*)h;
  size_t n = o->trace.size() < cap ? o->trace.size() : cap;
  for (size_t i = 0; i < n; ++i) { out[i].sample = o->trace[i].sample; out[i].zero = o->trace[i].zero; out[i].sym = o->trace[i].sym; }
}

// Derived constants (receiver.rs:502-560) so tests can compare with the engine's host-side derivation.
typedef struct oracle_derived {
  float sps, agc_bw, agc_gain0, samples_per_ted, period_min, period_max;
  float alpha_unlocked, beta_unlocked, alpha_locked, beta_locked;
  uint32_t dc_len, ntaps;
} oracle_derived;

void oracle_get_derived(const oracle_config* c, oracle_derived* d, float* mark_re_im, float* space_re_im, size_t cap_taps) {
  Config k = to_config(c);
  SameReceiver rx(k);
  d->sps = samples_per_symbol(k.input_rate);
  d->agc_bw = rx.agc.bandwidth; d->agc_gain0 = rx.agc.gain;
  d->samples_per_ted = rx.symsync.samples_per_ted; d->period_min = rx.symsync.period_min; d->period_max = rx.symsync.period_max;
  d->alpha_unlocked = rx.symsync.loop_alpha; d->beta_unlocked = rx.symsync.loop_beta;
  compute_loop_alphabeta(k.timing_bw_locked, d->alpha_locked, d->beta_locked);
  d->dc_len = (uint32_t)rx.dc_block.ff.len(); d->ntaps = (uint32_t)rx.demod.ntaps();
  for (size_t i = 0; i < rx.demod.ntaps() && i < cap_taps; ++i) {
    if (mark_re_im) { mark_re_im[2 * i] = rx.demod.coeff_mark[i].re; mark_re_im[2 * i + 1] = rx.demod.coeff_mark[i].im; }
    if (space_re_im) { space_re_im[2 * i] = rx.demod.coeff_space[i].re; space_re_im[2 * i + 1] = rx.demod.coeff_space[i].im; }
  }
}

// CPU baseline: decode `n_streams` independent streams (contiguous, `stride` samples apart, `len` samples each)
// with `n_threads` host threads, one receiver per stream.  Returns wall seconds of the decode loop;
// writes the number of link Burst events and Message events per stream if the arrays are given.
double oracle_decode_batch(const oracle_config* c, const int16_t* samples, size_t n_streams, size_t stride, size_t len,
                           int n_threads, uint32_t* n_bursts, uint32_t* n_msgs) {
  Config k = to_config(c);
  if (n_threads < 1) n_threads = 1;
  std::atomic<size_t> next{0};
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t) {
    th.emplace_back([&]() {
      std::vector<Event> ev;
      while (true) {
        size_t s = next.fetch_add(1);
        if (s >= n_streams) break;
        SameReceiver rx(k);
        ev.clear();
        const int16_t* p = samples + s * stride;
        for (size_t i = 0; i < len; ++i) rx.process_sample((float)p[i], ev);
        uint32_t nb = 0, nm = 0;
        for (auto& e : ev) {
          if (!e.is_transport && e.link.kind == LinkKind::Burst) nb++;
          if (e.is_transport && e.transport.kind == TransportKind::Message && e.transport.res.ok) nm++;
        }
        if (n_bursts) n_bursts[s] = nb;
        if (n_msgs) n_msgs[s] = nm;
      }
    });
  }
  for (auto& x : th) x.join();
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// ---- batch decode with the events handed back (full-size parity tests: every stream of config 3 / sampled streams of
// config 4 against the engine).  Records use the engine's same_event layout (include/same_engine.h) so that the test
// compares whole arrays: stream, seq (per-stream occurrence index), input_sample_counter, symbol_count, kind, err,
// data_offset, data_len, parity_errors, voting_bytes, flags; burst payloads are capped at 1024 bytes with flag 1 set,
// as the engine's burst buffer does (data_len keeps the true length).
typedef struct oracle_batch_event {
  uint32_t stream, seq;
  uint64_t input_sample_counter, symbol_count;
  uint32_t kind, err, data_offset, data_len;
  uint16_t parity_errors, voting_bytes;
  uint32_t flags;
} oracle_batch_event;

struct OracleBatch {
  std::vector<oracle_batch_event> ev;
  std::vector<uint8_t> payload;
  double seconds = 0.0;
};

// `lengths` may be NULL (every stream has `len` samples).  `flush` != 0 applies samedec's EOF flush rule per stream.
void* oracle_decode_batch_events(const oracle_config* c, const int16_t* samples, size_t n_streams, size_t stride, size_t len,
                                 const uint32_t* lengths, int n_threads, int flush) {
  Config k = to_config(c);
  if (n_threads < 1) n_threads = 1;
  std::vector<std::vector<oracle_batch_event>> evs(n_streams);
  std::vector<std::vector<uint8_t>> pays(n_streams);
  std::atomic<size_t> next{0};
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t) {
    th.emplace_back([&]() {
      std::vector<Event> ev;
      while (true) {
        size_t s = next.fetch_add(1);
        if (s >= n_streams) break;
        SameReceiver rx(k);
        ev.clear();
        const int16_t* p = samples + s * stride;
        const size_t n = lengths ? lengths[s] : len;
        for (size_t i = 0; i < n; ++i) rx.process_sample((float)p[i], ev);
        if (flush) samedec_eof_flush(rx, ev);
        uint32_t seq = 0;
        for (auto& e : ev) {
          oracle_batch_event r;
          memset(&r, 0, sizeof r);
          r.stream = (uint32_t)s; r.seq = seq++;
          r.input_sample_counter = e.input_sample_counter; r.symbol_count = e.symbol_count;
          if (!e.is_transport) r.kind = (uint32_t)e.link.kind;
          else if (e.transport.kind == TransportKind::Idle) r.kind = OR_EV_TR_IDLE;
          else if (e.transport.kind == TransportKind::Assembling) r.kind = OR_EV_TR_ASSEMBLING;
          else if (!e.transport.res.ok) { r.kind = OR_EV_TR_MSG_ERR; r.err = (uint32_t)e.transport.res.err; }
          else if (e.transport.res.msg.is_som) {
            r.kind = OR_EV_TR_MSG_SOM;
            r.parity_errors = (uint16_t)e.transport.res.msg.parity_error_count;
            r.voting_bytes = (uint16_t)e.transport.res.msg.voting_byte_count;
          } else r.kind = OR_EV_TR_MSG_EOM;
          std::string tmp;
          const std::string& txt = ev_text(e, tmp);
          r.data_len = (uint32_t)txt.size();
          size_t keep = txt.size();
          if (r.kind == OR_EV_LINK_BURST && keep > 1024) { keep = 1024; r.flags = 1; }
          r.data_offset = (uint32_t)pays[s].size();
          pays[s].insert(pays[s].end(), txt.begin(), txt.begin() + keep);
          evs[s].push_back(r);
        }
      }
    });
  }
  for (auto& x : th) x.join();
  auto t1 = std::chrono::steady_clock::now();
  auto* out = new OracleBatch();
  out->seconds = std::chrono::duration<double>(t1 - t0).count();
  for (size_t s = 0; s < n_streams; ++s) {
    const uint32_t base = (uint32_t)out->payload.size();
    for (auto r : evs[s]) { r.data_offset += base; out->ev.push_back(r); }
    out->payload.insert(out->payload.end(), pays[s].begin(), pays[s].end());
  }
  return out;
}
size_t oracle_batch_num_events(void* h) { return ((OracleBatch*)h)->ev.size(); }
size_t oracle_batch_payload_bytes(void* h) { return ((OracleBatch*)h)->payload.size(); }
double oracle_batch_seconds(void* h) { return ((OracleBatch*)h)->seconds; }
void oracle_batch_copy(void* h, oracle_batch_event* events, uint8_t* payload) {
  auto* b = (OracleBatch*)h;
  if (events && !b->ev.empty()) memcpy(events, b->ev.data(), b->ev.size() * sizeof(oracle_batch_event));
  if (payload && !b->payload.empty()) memcpy(payload, b->payload.data(), b->payload.size());
}
void oracle_batch_free(void* h) { delete (OracleBatch*)h; }

// Model of the long-stream path's speculative AGC (sameold_b200/csrc/same_long.cu, kernel B), on the oracle's own
// DCBlocker and Agc: the unlocked AGC recurrence over `n` samples in blocks of `block` samples, every block after the
// first warm-started `warm` samples earlier from `guess_gain`.  Counts the blocks whose warm-started gain differs
// BITWISE from the sequential trajectory at the block start (those the engine would recompute) and measures, per
// block, how many samples the two trajectories needed to coalesce.  Pins the property the path relies on.
void oracle_agc_block_model(const oracle_config* c, const int16_t* samples, size_t n, size_t block, size_t warm,
                            float guess_gain, uint32_t* n_blocks, uint32_t* n_mismatch, uint32_t* worst_coalesce) {
  Config k = to_config(c);
  SameReceiver rx(k);
  std::vector<float> d(n), g(n + 1);
  for (size_t i = 0; i < n; ++i) d[i] = rx.dc_block.filter((float)samples[i]);
  Agc agc = rx.agc;
  g[0] = agc.gain;
  for (size_t i = 0; i < n; ++i) { (void)agc.input(d[i]); g[i + 1] = agc.gain; }   // g[i] = gain before sample i
  uint32_t nb = 0, bad = 0, worst = 0;
  for (size_t sk = block; sk < n; sk += block) {
    Agc a = rx.agc;
    a.gain = guess_gain;
    const size_t w0 = sk - warm;
    uint32_t coalesced_after = (uint32_t)warm + 1u;
    for (size_t i = w0; i < sk; ++i) {
      if (coalesced_after > warm && a.gain == g[i]) coalesced_after = (uint32_t)(i - w0);
      (void)a.input(d[i]);
    }
    nb++;
    if (a.gain != g[sk]) bad++;
    else if (coalesced_after <= warm && coalesced_after > worst) worst = coalesced_after;
  }
  *n_blocks = nb; *n_mismatch = bad; *worst_coalesce = worst;
}

// CPU corpus generator for bench.py's reference arm (see synth_cpu.hpp): out[s * stride + n], n < n_samples, for
// n_streams streams on n_threads host threads.  burst_begin has n_streams + 1 entries (CSR into `bursts`).
void oracle_synth_generate(int16_t* out, size_t n_streams, size_t stride, size_t n_samples, uint32_t rate,
                           const uint32_t* burst_begin, const synth_cpu::Burst* bursts, const uint8_t* bytes,
                           size_t n_bytes_total, const float* freq_offset_hz, const uint32_t* seeds, float amplitude,
                           float noise_sigma, int n_threads) {
  std::vector<uint16_t> cum(n_bytes_total ? n_bytes_total : 1);
  for (uint32_t b = 0; b < burst_begin[n_streams]; ++b) {
    uint32_t acc = 0;
    for (uint32_t i = 0; i < bursts[b].n_bytes; ++i) {
      cum[bursts[b].byte_offset + i] = (uint16_t)acc;
      acc += (uint32_t)__builtin_popcount(bytes[bursts[b].byte_offset + i]);
    }
  }
  if (n_threads < 1) n_threads = 1;
  std::atomic<size_t> next{0};
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t) {
    th.emplace_back([&]() {
      while (true) {
        const size_t s = next.fetch_add(1);
        if (s >= n_streams) break;
        synth_cpu::render_stream(out + s * stride, n_samples, (double)rate, bursts + burst_begin[s],
                                 burst_begin[s + 1] - burst_begin[s], bytes, cum.data(), freq_offset_hz[s], seeds[s],
                                 amplitude, noise_sigma);
      }
    });
  }
  for (auto& x : th) x.join();
}

}  // extern "C"
