// selftest.cpp — pins the CPU oracle against the known-answer values of the reference's own unit tests.
// TEST INFRASTRUCTURE ONLY.  Each block cites the reference test it mirrors (crates/sameold/src/receiver/...).
// Run by tests/test_oracle_kat.py; prints "OK <n>" and exits 0 when every check passes.
#include "same_oracle.hpp"

#include <cstdio>
#include <cstdlib>

using namespace same_oracle;

static int g_checks = 0, g_fail = 0;
#define CHECK(cond)                                                              \
  do {                                                                           \
    ++g_checks;                                                                  \
    if (!(cond)) { ++g_fail; fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
  } while (0)
static bool approx(float a, float b, float eps = 1.0e-6f) { return fabsf(a - b) < eps; }  // assert_approx_eq default

// ---- test fixtures restated from waveform.rs:73-155 (cfg(test) helpers) ----
static std::vector<float> modulate_afsk(const std::vector<float>& syms, uint32_t fs, size_t& symlen_out) {  // :73-104
  const float TWOPI = 2.0f * 3.14159265358979323846f;
  float mark = TWOPI * FSK_MARK_HZ / (float)fs, space = TWOPI * FSK_SPACE_HZ / (float)fs;
  size_t symlen = f32_as_usize(floorf(samples_per_symbol(fs)));
  if (symlen % 2) symlen += 1;
  std::vector<float> out(syms.size() * symlen, 0.0f);
  float phase = 0.0f;
  for (size_t i = 0; i < out.size(); ++i) {
    bool sym = syms[i / symlen] >= 0.0f;
    phase += sym ? mark : space;
    if (phase > TWOPI) phase = -TWOPI + phase;
    out[i] = cosf(phase);
  }
  symlen_out = symlen;
  return out;
}
static std::vector<float> bytes_to_symbols(const std::vector<uint8_t>& b) {  // :112-128
  std::vector<float> v;
  for (uint8_t byte : b) for (int i = 0; i < 8; ++i) v.push_back(((byte >> i) & 1) ? 1.0f : -1.0f);
  return v;
}
static std::vector<float> bytes_to_samples(const std::vector<uint8_t>& b, size_t nsps) {  // :137-155
  if (nsps < 1) nsps = 1;
  std::vector<float> v;
  for (uint8_t byte : b) for (int i = 0; i < 8; ++i) {
    for (size_t k = 0; k + 1 < nsps; ++k) v.push_back(0.0f);
    v.push_back(((byte >> i) & 1) ? 1.0f : -1.0f);
  }
  return v;
}

static const char* TEST_MESSAGE =
    "ZCZC-EAS-DMO-372088-091724-919623-645687-745748-175234-039940-955869-091611-304171-931612-334828-179485-"
    "569615-809223-830187-611340-014693-472885-084645-977764-466883-406863-390018-701741-058097-752790-311648-"
    "820127-255900-581947+0000-0001122-NOCALL00-";

static void test_dcblock() {  // dcblock.rs:118-173
  { MovingAverage m(1); float a, d; m.filter(1.0f, a, d); CHECK(d == 1.0f && approx(a, 1.0f));
    m.filter(-10.0f, a, d); CHECK(d == -10.0f && approx(a, -10.0f)); }
  { MovingAverage m(2); float a, d; m.filter(1.0f, a, d); CHECK(d == 0.0f && approx(a, 0.5f));
    m.filter(2.0f, a, d); CHECK(d == 1.0f && approx(a, 1.5f)); }
  { const float in[] = {1, 2, -1, 3, 8}, ex[] = {0.25f, 0.75f, 0.5f, 1.25f, 3.0f};
    MovingAverage m(4); float a, d = 0;
    for (int i = 0; i < 5; ++i) { m.filter(in[i], a, d); CHECK(approx(a, ex[i])); }
    CHECK(d == 2.0f); }
  { DCBlocker u(1); CHECK(u.filter(100.0f) == 100.0f); CHECK(u.filter(-200.0f) == -200.0f); }
  { DCBlocker u(31); float clk = 1.0f, h0 = 0, h1 = 0;
    for (int i = 0; i < 256; ++i) { h0 = h1; h1 = u.filter(100.0f + clk); clk = -clk; }
    CHECK(approx(h0, 1.0f, 1e-2f)); CHECK(approx(h1, -1.0f, 1e-2f)); }
}

static void test_agc() {  // agc.rs:105-125
  Agc agc(0.05f, 0.0f, 1.0e6f); float v = 0;
  for (int i = 0; i < 256; ++i) v = agc.input(-2.0f);
  CHECK(approx(agc.gain, 0.5f)); CHECK(approx(v, -1.0f));
  agc.reset(); agc.lock(true);
  for (int i = 0; i < 16; ++i) v = agc.input(-2.0f);
  CHECK(agc.gain == 1.0f); CHECK(approx(v, -2.0f));
}

static void test_filter() {  // filter.rs:388-462
  { std::deque<float> h; std::vector<float> c; CHECK(mac_real(h, c) == 0.0f); }
  { std::deque<float> h{20.0f, 1.0f}; std::vector<float> c{1.0f}; CHECK(mac_real(h, c) == 1.0f); }
  { std::deque<float> h{1.0f}; std::vector<float> c{1.0f, 20.0f}; CHECK(mac_real(h, c) == 1.0f); }
  { std::deque<float> h{20.0f, 20.0f}; std::vector<float> c{1.0f, -1.0f}; CHECK(approx(mac_real(h, c), 0.0f)); }
  { Window<float> w(4);
    float a1[] = {1.0f}; w.push(a1, 1); CHECK(w.q[3] == 1.0f && w.q[2] == 0.0f);
    w.push(a1, 0); CHECK(w.q[3] == 1.0f);
    float a2[] = {2.0f}; w.push(a2, 1); CHECK(w.q[2] == 1.0f && w.q[3] == 2.0f);
    float a6[] = {-1, -2, 1, 2, 3, 4}; w.push(a6, 6);
    CHECK(w.q[0] == 1 && w.q[1] == 2 && w.q[2] == 3 && w.q[3] == 4 && w.back() == 4 && w.front() == 1 && w.len() == 4);
    CHECK(w.push_scalar(10.0f) == 1.0f); CHECK(w.q[0] == 2 && w.q[3] == 10);
    float a4[] = {5, 4, 3, 2}; w.push(a4, 4); CHECK(w.q[0] == 5 && w.q[3] == 2);
    w.reset(); CHECK(w.len() == 4 && w.q[0] == 0 && w.q[3] == 0); }
}

static void test_waveform() {  // waveform.rs:162-173, 175-187
  const float ER[] = {-0.719973f, -0.208581f, 0.374184f, 0.828910f, 1.000000f};
  const float EI[] = {-0.694002f, -0.978005f, -0.927355f, -0.559382f, -0.000000f};
  auto out = cisoid_matched_filter(5, 0.0944807256f);
  float gain = 2.0f / 5.0f;
  for (int i = 0; i < 5; ++i) {
    float dr = out[i].re - gain * ER[i], di = out[i].im - gain * EI[i];
    CHECK(hypotf(dr, di) < 1e-4f);
  }
  const float ES[] = {1, 1, -1, 1, -1, 1, -1, 1, 1, -1, -1, -1, -1, 1, -1, -1};
  auto s = bytes_to_symbols({0xAB, 0x21});
  CHECK(s.size() == 16);
  for (int i = 0; i < 16; ++i) CHECK(s[i] == ES[i]);
  CHECK(PREAMBLE_SYNC_WORD == 0xabababab);
}

static void test_demod() {  // demod.rs:197-227
  const std::vector<float> syms{1, -1, 1, -1, -1};
  size_t sps; auto mod = modulate_afsk(syms, 11025, sps);
  size_t delay = sps / 2;
  mod.insert(mod.end(), delay, 0.0f);
  FskDemod d = FskDemod::from_same(11025);
  for (size_t i = 0, c = 0; i < mod.size(); i += delay, ++c) {
    size_t n = std::min(delay, mod.size() - i);
    d.push(mod.data() + i, n);
    float sym = d.demod();
    if (c % 2 == 0) continue;
    size_t bit = (c - 1) / 2;
    if (syms[bit] >= 0.0f) CHECK(sym >= 0.95f); else CHECK(sym <= 0.95f);
  }
}

static SymbolEstimate timing_test(TimingLoop& t, const std::vector<float>& inp, size_t start) {  // symsync.rs:452-485
  float offset = 0.0f; size_t sa = start; SymbolEstimate last{{0, 0}, 0};
  t.reset();
  for (int i = 0; i < 128; ++i) {
    bool have; SymbolEstimate s;
    float skip = t.input(inp[sa], offset, have, s);
    float whole = roundf(skip);
    offset = skip - whole;
    sa += (size_t)whole; sa %= inp.size();
    if (have) last = s;
  }
  return last;
}

static void test_symsync() {  // symsync.rs:357-563
  auto zc = [](float a, float b, float c) { return b * (signumf(a) - signumf(c)); };
  CHECK(approx(zc(1, 0, -1), 0)); CHECK(approx(zc(-1, 0, 1), 0)); CHECK(approx(zc(1, 1, 1), 0));
  CHECK(approx(zc(-1, -1, -1), 0)); CHECK(approx(zc(0.8f, 0.2f, -0.8f), 0.4f)); CHECK(approx(zc(0.8f, -0.2f, -0.8f), -0.4f));
  float a, b;
  compute_loop_alphabeta(0.0f, a, b); CHECK(approx(a, 0) && approx(b, 0));
  compute_loop_alphabeta(0.5f, a, b); CHECK(approx(a, 0.99813f, 1e-4f) && approx(b, 0.91544f, 1e-4f));
  compute_loop_alphabeta(1.0f, a, b); CHECK(approx(a, 1.0f, 1e-4f) && approx(b, 0.99627f, 1e-4f));
  { ZeroCrossingTed ted; SymbolEstimate s;
    CHECK(ted.input(0.8f, s)); CHECK(!ted.input(0.2f, s));
    CHECK(ted.input(-0.8f, s)); CHECK(s.data[1] == -0.8f && approx(s.err, 0.4f));
    CHECK(!ted.input(0.2f, s)); CHECK(ted.input(0.8f, s)); CHECK(s.data[1] == 0.8f && approx(s.err, -0.4f)); }
  { TimingLoop t(32.0f, 0.25f, 0.125f);
    CHECK(approx(t.period_inst, 16.0f)); CHECK(approx(t.period_max, 20.0f));
    CHECK(approx(t.advance_loop(0.0f, nullptr), 16.0f)); CHECK(approx(t.advance_loop(0.5f, nullptr), 16.5f));
    CHECK(approx(t.advance_loop(-0.5f, nullptr), 16.0f)); CHECK(approx(t.advance_loop(-0.5f, nullptr), 15.5f));
    t.reset(); CHECK(approx(t.period_inst, 16.0f));
    SymbolEstimate s{{0.0f, 1.0f}, 0.0f}; CHECK(approx(t.advance_loop(0.0f, &s), 16.0f));
    SymbolEstimate e{{0.0f, 0.95f}, 0.5f / 16.0f}; CHECK(approx(t.advance_loop(0.5f, &e), 16.5f));
    SymbolEstimate l{{0.0f, 0.95f}, -0.5f / 16.0f}; CHECK(approx(t.advance_loop(-0.5f, &l), 15.5f)); }
  { std::vector<float> inp(64);
    for (int n = 0; n < 64; ++n) inp[n] = sinf((2.0f * 3.14159265358979323846f) * (float)n / 64.0f);
    CHECK(approx(inp[0], 0)); CHECK(approx(inp[16], 1)); CHECK(approx(inp[48], -1));
    struct C { float bw; size_t start; } cases[] = {{0.25f, 16}, {0.25f, 15}, {0.25f, 0}, {0.20f, 16}, {0.05f, 3}};
    for (auto& c : cases) {
      TimingLoop t(32.0f, c.bw, 0.125f);
      SymbolEstimate s = timing_test(t, inp, c.start);
      CHECK(fabsf(s.data[1]) > 0.99f); CHECK(s.err < 1e-4f);
    } }
}

static void test_codesquelch() {  // codesquelch.rs:499-667
  auto nbe = [](uint32_t a, uint32_t b) { return (uint32_t)__builtin_popcount(a ^ b); };
  CHECK(nbe(PREAMBLE_SYNC_WORD, PREAMBLE_SYNC_WORD) == 0); CHECK(nbe(PREAMBLE_SYNC_WORD, PREAMBLE_SYNC_WORD | 0x40u) == 1);
  CHECK(nbe(PREAMBLE_SYNC_WORD, 0xa9ababab) == 1);
  { // test_codecorr (through the squelch's correlator fields)
    auto syms = bytes_to_symbols({0xAB, 0xAB, 0xAB, 0xAB, 0x21});
    CodeAndPowerSquelch q(PREAMBLE_SYNC_WORD, 0, 0, 0, 0.1f);
    auto search = [&](float s) { float in[2] = {0, s}; q.input(in); return nbe(q.sync_to, q.data); };
    for (size_t i = 0; i < syms.size(); ++i) { uint32_t e = search(syms[i]); if (i == 31) CHECK(e == 0); else CHECK(e > 0); }
    syms[19] = -syms[19];
    for (size_t i = 0; i < syms.size(); ++i) { uint32_t e = search(syms[i]); if (i == 31) CHECK(e == 1); else CHECK(e >= 1); } }
  { // test_power_tracker
    CodeAndPowerSquelch q(PREAMBLE_SYNC_WORD, 0, 0, 0, 1.0f);
    float in[2] = {0, 1.0f}; q.input(in); q.pt_bandwidth = 0.5f;
    float in2[2] = {0, -0.5f}; q.input(in2); CHECK(approx(q.pt_power, 0.625f));
    q.pt_power = 1.0f; for (int i = 0; i < 16; ++i) q.input(in); CHECK(approx(q.pt_power, 1.0f)); }
  { // test_simple_sync
    auto ins = bytes_to_samples({0xAB, 0xAB, 0xAB, 0xAB, 0x21}, 2);
    CodeAndPowerSquelch q(PREAMBLE_SYNC_WORD, 0, 0.0f, 0.0f, 0.1f);
    CHECK(!q.is_sync());
    size_t align = 0;
    for (size_t chunk = 0; chunk * 2 < ins.size(); ++chunk) {
      SquelchState s = q.input(&ins[chunk * 2]);
      if (s.kind == SquelchKind::Ready) {
        CHECK(!s.resync || chunk == 31);
        if (chunk == 31) CHECK(q.data == 0xabababab);
        for (int i = 0; i < 16; ++i) CHECK(s.out.samples[i] == ins[align + i]);
        align += 16;
        CHECK(s.out.symbol_counter - 1 == chunk);
      }
    }
    CHECK(q.is_sync()); CHECK(align == 32); q.end(); CHECK(!q.is_sync()); }
  { // test_sync_with_error
    auto ins = bytes_to_samples({0xF0, 0x0B, 0xA9, 0xAB, 0xAB, 0xAB, 0x21}, 2);
    CodeAndPowerSquelch q(PREAMBLE_SYNC_WORD, 1, 0.0f, 0.0f, 0.1f);
    size_t align = 32;
    for (size_t chunk = 0; chunk * 2 < ins.size(); ++chunk) {
      SquelchState s = q.input(&ins[chunk * 2]);
      if (s.kind == SquelchKind::Ready) {
        CHECK(!s.resync || chunk == 47);
        for (int i = 0; i < 16; ++i) CHECK(s.out.samples[i] == ins[align + i]);
        align += 16;
      }
    }
    CHECK(q.is_sync()); }
  { // test_sync_with_lots_of_errors
    auto ins = bytes_to_samples({0xAB, 0x0B, 0xA9, 0xAB, 0xAB, 0xAA, 0x21}, 2);
    CodeAndPowerSquelch q(PREAMBLE_SYNC_WORD, 3, 0.8f, 0.1f, 0.1f);
    bool early = false, later = false; size_t align = 32;
    for (size_t chunk = 0; chunk * 2 < ins.size(); ++chunk) {
      SquelchState s = q.input(&ins[chunk * 2]);
      if (s.kind == SquelchKind::Ready) {
        if (chunk == 47) { for (int i = 0; i < 16; ++i) CHECK(s.out.samples[i] == ins[align + i]); align += 16; later = true; }
        else early = true;
      }
    }
    CHECK(q.is_sync() && later && early); }
  { // test_power_detection
    auto ins = bytes_to_samples({0xF0, 0x0B, 0xA9, 0xAB, 0xAB, 0xAB, 0x21}, 2);
    CodeAndPowerSquelch q(PREAMBLE_SYNC_WORD, 1, 0.9f, 0.5f, 0.1f);
    for (size_t chunk = 0; chunk * 2 < ins.size(); ++chunk) q.input(&ins[chunk * 2]);
    CHECK(q.is_sync());
    bool r = false, d = false, n = false; float z[2] = {0, 0};
    for (int i = 0; i < 40; ++i) {
      SquelchState s = q.input(z);
      if (s.kind == SquelchKind::Reading) r = true;
      if (s.kind == SquelchKind::DroppedCarrier) d = true;
      if (s.kind == SquelchKind::NoCarrier) n = true;
    }
    CHECK(r && d && n && !q.is_sync()); }
}

static void test_equalize() {  // equalize.rs:412-593
  { const float in[] = {0.0f, 0.5f, 0.0f, -0.5f};
    Equalizer e(8, 4, 0.2f, 1.0e-5f, false, 0); e.enable(false);
    float err; bool b0 = e.estimate_symbol(in, err); CHECK(b0 && approx(err, 0));
    bool b1 = e.estimate_symbol(in + 2, err); CHECK(!b1 && approx(err, 0));
    e.enable(true);
    for (int i = 0; i < 32; ++i) e.estimate_symbol(in + 2 * (i % 2), err);
    CHECK(fabsf(err) < 1.0e-5f); }
  { // test_nlms_evolve (Proakis B)
    const float in[] = {0, 1, 0, -1}; std::vector<float> ch{0.407f, 0.815f, 0.407f};
    Window<float> cw(3), iw(3); std::vector<float> inv{1, 0, 0}; float err = 0;
    for (int i = 0; i < 128; ++i) {
      float s = in[i % 4]; cw.push(&s, 1); float chs = mac_real(cw.q, ch);
      iw.push(&chs, 1); float est = mac_real(iw.q, inv); err = s - est;
      Equalizer::nlms_update(0.10f, 1.0e-6f, err, iw.q, inv);
    }
    CHECK(fabsf(err) < 1e-2f); }
  { // test_estimate_symbol_proakis
    std::vector<float> ch{0.8f, -0.2f}; const float in[] = {0, 1, 0, -1};
    Window<float> cw(2); Equalizer u(8, 4, 0.2f, 1.0e-5f, false, 0); float err = 0; bool bit = false;
    auto run = [&](const float* s2) { float c[2]; for (int k = 0; k < 2; ++k) { cw.push(&s2[k], 1); c[k] = mac_real(cw.q, ch); } bit = u.estimate_symbol(c, err); };
    for (int i = 0; i < 32; ++i) run(in + 2 * (i % 2));
    CHECK(fabsf(err) < 1e-4f);
    for (int i = 0; i < 2; ++i) { run(in + 2 * i); CHECK((in[2 * i + 1] >= 0.0f) == bit); CHECK(fabsf(err) < 1e-4f); } }
  { // test_estimate_symbol_training
    Equalizer u(8, 4, 0.2f, 1.0e-5f, true, PREAMBLE_SYNC_WORD);
    CHECK(u.train()); CHECK(u.mode == Equalizer::EnabledTraining && u.train_sa == PREAMBLE_SYNC_WORD && u.train_count == 0);
    auto sig = bytes_to_samples({0x54, 0x54}, 2); float err;
    for (size_t i = 0; i < sig.size(); i += 2) u.estimate_symbol(&sig[i], err);
    CHECK(u.mode == Equalizer::EnabledTraining && u.train_count == 16);
    for (size_t i = 0; i < sig.size(); i += 2) u.estimate_symbol(&sig[i], err);
    CHECK(u.mode == Equalizer::EnabledFeedback);
    float neg[2] = {0.0f, -1.0f}; CHECK(u.estimate_symbol(neg, err));  // evolved a bit-flipping DFE
    u.reset(); CHECK(u.train());
    auto sig2 = bytes_to_samples({0xAB, 0xAB, 0xAB, 0xAB}, 2);
    for (size_t i = 0; i < sig2.size(); i += 2) u.estimate_symbol(&sig2[i], err);
    CHECK(u.mode == Equalizer::EnabledFeedback); CHECK(!u.estimate_symbol(neg, err)); }
  { auto sig = bytes_to_samples({0xAB, 0xBA}, 2); Equalizer u(8, 4, 0.2f, 1.0e-5f, false, 0); float err;
    CHECK(u.input(&sig[0], err) == 0xAB); CHECK(u.input(&sig[16], err) == 0xBA); }
  { Equalizer u(6, 4, 0.05f, 1e-6f, false, 0); CHECK(!u.train()); }  // NoTrainingSequenceErr
}

static bool active(const LinkState& l) { return l.kind == LinkKind::Searching || l.kind == LinkKind::Reading; }

static void test_framing() {  // framing.rs:259-349
  CHECK(message_prefix_errors(0x5A435A43u) == 0); CHECK(message_prefix_errors(0x4E4E4E4Eu) == 0);
  CHECK(message_prefix_errors(0xABABABABu) == 18); CHECK(message_prefix_errors(0x5A435A45u) == 2);
  { Framer f(1, 10); bool gave_up = false;
    for (uint32_t i = 0; i < 32; ++i) {
      LinkState l = f.input(PREAMBLE, 0, i == 0);
      if (l.kind == LinkKind::NoCarrier) { CHECK(i >= Framer::PREFIX_SEARCH_LEN); gave_up = true; }
      else CHECK(l.kind == LinkKind::Searching);
    }
    CHECK(gave_up);
    f.input(PREAMBLE, 0, true); f.input(PREAMBLE, 0, true);
    LinkState last; for (char c : std::string("ZCZC")) { last = f.input((uint8_t)c, 0, false); CHECK(active(last)); }
    CHECK(last.kind == LinkKind::Reading); CHECK(f.st == Framer::DataRead && f.msg.size() == 4 && f.invalid == 0);
    LinkState e = f.end(); CHECK(e.kind == LinkKind::Burst && std::string(e.burst.begin(), e.burst.end()) == "ZCZC");
    CHECK(f.end().kind == LinkKind::NoCarrier); }
  { const std::string M = "gArbAZgEZCZC-ORG-EEE-012345-567890+0000-0001122-NOCALL00-GARBAGE"; const uint32_t PI = 10;
    Framer f(2, PI); bool found = false;
    f.input(PREAMBLE, 0, true);
    for (char c : M) CHECK(active(f.input((uint8_t)c, 0, false)));
    for (uint32_t j = 0; j < PI + 1; ++j) {
      LinkState o = f.input(PREAMBLE, 0, false);
      if (j >= PI) {
        CHECK(o.kind == LinkKind::Burst);
        std::string s(o.burst.begin(), o.burst.end());
        CHECK(s.rfind("ZCZC-ORG-EEE-012345-567890+0000-0001122-NOCALL00-", 0) == 0); found = true;
      } else CHECK(active(o));
    }
    CHECK(found); }
}

static std::vector<uint8_t> B(const char* s) { return std::vector<uint8_t>(s, s + strlen(s)); }

static void test_combiner() {  // combiner.rs:279-441
  uint8_t o; uint32_t n;
  struct D { uint8_t a, b, o; uint32_t n; } det[] = {{0xab, 0xab, 0xab, 0}, {0xff, 0xff, 0xff, 0}, {0, 0, 0, 0}, {0, 1, 0, 1},
                                                    {2, 1, 0, 2}, {0xff, 0xf0, 0, 4}, {0x0f, 0xf0, 0, 8}, {0xff, 0, 0, 8}};
  for (auto& d : det) { bit_vote_detect(d.a, d.b, o, n); CHECK(o == d.o && n == d.n); }
  struct C { uint8_t a, b, c, o; uint32_t n; } cor[] = {{0xab, 0xab, 0xab, 0xab, 0}, {0xff, 0xff, 0xff, 0xff, 0}, {0, 0, 0, 0, 0},
      {0xaa, 0xab, 0xab, 0xab, 1}, {0xa0, 0xa0, 0xaf, 0xa0, 4}, {0x0f, 0xf0, 0xff, 0xff, 8}, {0x00, 0xf0, 0xff, 0xf0, 8},
      {0xaa, 0x55, 0xff, 0xff, 8}, {0xaa, 0x55, 0xa5, 0xa5, 8}};
  for (auto& c : cor) { bit_vote_correct(c.a, c.b, c.c, o, n); CHECK(o == c.o && n == c.n); }
  auto est = [](std::vector<std::vector<uint8_t>> bs) { std::vector<const std::vector<uint8_t>*> p; for (auto& b : bs) p.push_back(&b); return estimate_message(p); };
  auto str = [](const std::vector<uint8_t>& v) { return std::string(v.begin(), v.end()); };
  { auto e = est({B("")}); CHECK(e.bytes.empty() && e.nbursts.empty() && e.errs.empty()); }
  { auto e = est({B("@@"), B("")}); CHECK(e.bytes.empty()); }
  { auto e = est({B("HIHI"), B("HI")}); CHECK(str(e.bytes) == "HIHI"); CHECK((e.nbursts == std::vector<uint8_t>{2, 2, 1, 1})); CHECK((e.errs == std::vector<uint8_t>{0, 0, 0, 0})); }
  { auto e = est({B("TEST"), B("TESZ"), B("")}); CHECK(str(e.bytes) == "TES"); CHECK((e.nbursts == std::vector<uint8_t>{2, 2, 2})); }
  { auto e = est({B("NNNN"), B("NNNN"), B("ZCZC-")}); CHECK(str(e.bytes) == "NNNN-"); CHECK((e.nbursts == std::vector<uint8_t>{3, 3, 3, 3, 1})); CHECK((e.errs == std::vector<uint8_t>{2, 3, 2, 3, 0})); }
  { auto e = est({B("NNNN"), B("NNNNB"), B("ZC")}); CHECK(str(e.bytes) == "NNNNB"); CHECK((e.nbursts == std::vector<uint8_t>{3, 3, 2, 2, 1})); CHECK((e.errs == std::vector<uint8_t>{2, 3, 0, 0, 0})); }
  { auto e = est({{0xce, 'N'}, B("NN")}); CHECK(str(e.bytes) == "NN"); CHECK((e.errs == std::vector<uint8_t>{1, 0})); }
  { auto e = est({{0xce, 'N'}, B("NN"), {'N', 0xce}}); CHECK(str(e.bytes) == "NN"); CHECK((e.nbursts == std::vector<uint8_t>{3, 3})); CHECK((e.errs == std::vector<uint8_t>{1, 1})); }
  auto comb = [](std::vector<std::vector<uint8_t>> bs, MessageResult& r) { std::vector<const std::vector<uint8_t>*> p; for (auto& b : bs) p.push_back(&b); return combine(p, r); };
  const char* MSG = "ZCZC-EAS-DMO-999000+0015-0011122-NOCALL00-"; const char* COR = "ZKZK-EAS-DMO-999000+0015-0011122-NOCALL00-";
  MessageResult r;
  CHECK(!comb({B(MSG)}, r));
  CHECK(comb({B("NNZZ")}, r) && r.ok && !r.msg.is_som);
  { std::vector<uint8_t> part = B(MSG); part.resize(16); CHECK(comb({B(MSG), part}, r) && !r.ok && r.err == DecodeErr::Malformed); }
  CHECK(comb({B("NOPE"), B("NOPE")}, r) && !r.ok && r.err == DecodeErr::UnrecognizedPrefix);
  CHECK(comb({B(MSG), B(MSG)}, r) && r.ok && r.msg.text == MSG && r.msg.voting_byte_count == 0);
  CHECK(comb({B(MSG), B(MSG), B(COR)}, r) && r.ok && r.msg.text == MSG && r.msg.voting_byte_count == strlen(MSG) && r.msg.parity_error_count == 2);
  CHECK(comb({B("NNZZ"), B(MSG), B(MSG)}, r) && r.ok && r.msg.text == MSG && r.msg.voting_byte_count == 4);
}

static void test_message() {  // sameplace message.rs:911-929
  size_t off, len;
  CHECK(!check_header("ZCZC-ORG-EEE-+0000-0001122-NOCALL00-", off, len));
  CHECK(check_header("ZCZC-ORG-EEE-012345+0000-0001122-NOCALL00-", off, len) && off == 19 && len == 42);
  CHECK(check_header("ZCZC-ORG-EEE-012345-567890+0000-0001122-NOCALL00-garbage", off, len) && off == 26 && len == 49);
  CHECK(check_header("ZCZC-PEP-NPT-000000+0030-2771820-TEST    -", off, len) && len == 42);
}

static void test_assembler() {  // assembler.rs:418-779
  const uint64_t ONE = (uint64_t)BAUD_HZ, BT = (uint64_t)(1.31f * BAUD_HZ), AT = (uint64_t)(1.2f * BAUD_HZ);
  CHECK(max_interburst_symbols() == 682); CHECK(max_history_duration() == 5652);
  const auto EOM = B("NNNN"), GOOD = B("ZCZC-EAS-DMO-999000+0015-0011122-NOCALL00-"),
             ERRS = B("ZCZK-EAS-DMF-999!00+0015-0011122-NOCALL00-KXYZ"), LONGEST = B(TEST_MESSAGE), NONE = B("");
  struct Step { uint64_t delay; const std::vector<uint8_t>* d; };
  auto sim = [](uint64_t& t, const Step& s) { t += 8 * s.d->size() + s.delay; if (!s.d->empty()) t += 16 * 8; return t; };
  { Assembler a; a.history.push_back({{}, 1}); a.history.push_back({{}, 2144}); a.history.push_back({{}, 3000});
    a.prune_history(0); CHECK(a.history.size() == 2);
    a.prune_history(2143); CHECK(a.history.size() == 2 && a.history[0].deadline == 2144 && a.history[1].deadline == 3000);
    a.prune_history(2999); CHECK(a.history.size() == 1 && a.history[0].deadline == 3000);
    a.prune_history(6000); CHECK(a.history.empty()); }
  { // test_pending_result
    std::vector<uint8_t> noerr(GOOD.size(), 0), v2(GOOD.size(), 2), v3(GOOD.size(), 3);
    MessageResult nov, vot; nov.ok = vot.ok = true;
    CHECK(message_try_from(GOOD.data(), GOOD.size(), noerr.data(), noerr.size(), v2.data(), v2.size(), nov.msg) == DecodeErr::None);
    CHECK(message_try_from(GOOD.data(), GOOD.size(), noerr.data(), noerr.size(), v3.data(), v3.size(), vot.msg) == DecodeErr::None);
    MessageResult e1, e2, eom; e1.err = DecodeErr::NotAscii; e2.err = DecodeErr::UnrecognizedPrefix; eom.ok = true; eom.msg.text = "NNNN";
    Assembler u; auto poll = [&](uint64_t now, MessageResult& out) { if (u.pending && u.pending_deadline <= now) { out = u.pending_res; u.pending = false; return true; } return false; };
    MessageResult o;
    CHECK(u.accept(e1, 0)); CHECK(u.accept(e2, 0)); CHECK(!poll(0, o));
    CHECK(u.accept(eom, 0)); CHECK(!u.accept(eom, 0)); CHECK(poll(0, o) && o == eom); CHECK(!u.pending);
    CHECK(u.accept(nov, 0)); CHECK(u.accept(nov, 0)); CHECK(!poll(0, o));
    CHECK(u.accept(vot, 5650)); CHECK(!u.accept(eom, 5650)); CHECK(!u.accept(e1, 5650)); CHECK(!u.accept(nov, 5650));
    CHECK(!poll(5650, o)); CHECK(poll(2 * 5650, o) && o == vot); }
  auto isEOM = [](const TransportState& t) { return t.kind == TransportKind::Message && t.res.ok && !t.res.msg.is_som; };
  auto isSOM = [](const TransportState& t) { return t.kind == TransportKind::Message && t.res.ok && t.res.msg.is_som; };
  { // test_assembler_deduplicate
    Step st[] = {{999 * ONE, &NONE}, {0, &EOM}, {ONE, &EOM}, {ONE, &EOM}, {12 * ONE, &EOM}};
    Assembler a; uint64_t t = 0; int i = 0;
    for (auto& s : st) { uint64_t tm = sim(t, s); TransportState o = a.assemble(*s.d, tm);
      switch (i++) { case 0: CHECK(o.kind == TransportKind::Idle && !a.pending); break; case 1: CHECK(isEOM(o) && !a.pending); break;
        case 2: case 3: CHECK(o.kind == TransportKind::Assembling && !a.pending); break; case 4: CHECK(isEOM(o) && !a.pending); break; } } }
  { // test_assembler_normal_operation
    Step st[] = {{0, &GOOD}, {ONE, &NONE}, {0, &GOOD}, {ONE, &NONE}, {0, &ERRS}, {BT, &NONE}, {15 * ONE, &EOM}, {ONE, &EOM}, {ONE, &EOM}};
    Assembler a; uint64_t t = 0; int i = 0;
    for (auto& s : st) { uint64_t tm = sim(t, s); TransportState o = a.assemble(*s.d, tm);
      switch (i++) { case 0: case 1: CHECK(o.kind == TransportKind::Assembling && !a.pending); break;
        case 2: case 3: case 4: CHECK(o.kind == TransportKind::Assembling && a.pending); break;
        case 5: CHECK(isSOM(o) && !a.pending && o.res.msg.voting_byte_count == GOOD.size()); break;
        case 6: CHECK(isEOM(o) && !a.pending); break; case 7: case 8: CHECK(o.kind == TransportKind::Assembling && !a.pending); break; } } }
  { // test_assembler_very_long_message
    Step st[] = {{0, &LONGEST}, {AT, &NONE}, {0, &LONGEST}, {AT, &NONE}, {0, &LONGEST}, {BT, &NONE}};
    Assembler a; uint64_t t = 0; int i = 0;
    for (auto& s : st) { uint64_t tm = sim(t, s); TransportState o = a.assemble(*s.d, tm);
      switch (i++) { case 0: case 1: CHECK(o.kind == TransportKind::Assembling && !a.pending); break;
        case 2: case 3: case 4: CHECK(o.kind == TransportKind::Assembling && a.pending); break;
        case 5: CHECK(isSOM(o) && !a.pending && o.res.msg.voting_byte_count == LONGEST.size() && o.res.msg.text == TEST_MESSAGE); break; } } }
  { // test_assembler_very_long_message_missing_middle
    Step st[] = {{0, &LONGEST}, {AT, &NONE}, {268 * 8, &NONE}, {AT, &NONE}, {0, &LONGEST}, {BT, &NONE}};
    Assembler a; uint64_t t = 0; int i = 0;
    for (auto& s : st) { uint64_t tm = sim(t, s); TransportState o = a.assemble(*s.d, tm);
      if (i == 4) CHECK(o.kind == TransportKind::Assembling && a.pending);
      else if (i == 5) CHECK(isSOM(o) && !a.pending && o.res.msg.voting_byte_count == 0 && o.res.msg.text == TEST_MESSAGE);
      else CHECK(o.kind == TransportKind::Assembling && !a.pending);
      ++i; } }
  { // test_assembler_quickly_with_missing
    Step st[] = {{0, &EOM}, {ONE, &EOM}, {ONE, &GOOD}, {(uint64_t)(1.1f * (float)ONE), &GOOD}, {BT, &NONE}, {ONE, &EOM}, {ONE, &EOM}};
    Assembler a; uint64_t t = 0; int i = 0;
    for (auto& s : st) { uint64_t tm = sim(t, s); TransportState o = a.assemble(*s.d, tm);
      switch (i++) { case 0: CHECK(isEOM(o) && !a.pending); break; case 1: case 2: CHECK(o.kind == TransportKind::Assembling && !a.pending); break;
        case 3: CHECK(o.kind == TransportKind::Assembling && a.pending); break;
        case 4: CHECK(isSOM(o) && !a.pending && o.res.msg.voting_byte_count == 4); break;
        case 5: CHECK(o.kind == TransportKind::Assembling && !a.pending); break; case 6: CHECK(isEOM(o) && !a.pending); break; } } }
}

static std::vector<float> make_test_burst(const std::string& payload, size_t num_bursts) {  // receiver.rs:611-639
  std::vector<uint8_t> msg(16, PREAMBLE); msg.insert(msg.end(), payload.begin(), payload.end());
  auto low = bytes_to_samples(msg, 1); size_t sps; auto high = modulate_afsk(low, 22050, sps);
  std::vector<float> burst(high.size()); for (size_t i = 0; i < high.size(); ++i) burst[i] = high[i] * 16384.0f;
  std::vector<float> out = burst;
  for (size_t i = 1; i < num_bursts; ++i) { out.insert(out.end(), 22050, 0.0f); out.insert(out.end(), burst.begin(), burst.end()); }
  out.insert(out.end(), 2 * 22050, 0.0f);
  return out;
}

static void test_receiver() {  // receiver.rs:641-705
  { auto afsk = make_test_burst(TEST_MESSAGE, 1);
    Config c; c.timing_max_deviation = 0.01f; SameReceiver rx(c); std::vector<Event> ev;
    for (float s : afsk) rx.process_sample(s, ev);
    CHECK(ev.size() == 5);
    if (ev.size() == 5) {
      CHECK(!ev[0].is_transport && ev[0].link.kind == LinkKind::Searching);
      CHECK(!ev[1].is_transport && ev[1].link.kind == LinkKind::Reading);
      CHECK(!ev[2].is_transport && ev[2].link.kind == LinkKind::Burst);
      std::string b(ev[2].link.burst.begin(), ev[2].link.burst.end()); CHECK(b.rfind(TEST_MESSAGE, 0) == 0);
      CHECK(ev[3].is_transport && ev[3].transport.kind == TransportKind::Assembling);
      CHECK(!ev[4].is_transport && ev[4].link.kind == LinkKind::NoCarrier);
    } }
  { auto afsk = make_test_burst(TEST_MESSAGE, 3);
    Config c; c.timing_max_deviation = 0.01f; SameReceiver rx(c); std::vector<Event> ev;
    bool got = false; size_t i = 0;
    for (; i < afsk.size() && !got; ++i) { size_t b = ev.size(); rx.process_sample(afsk[i], ev);
      for (size_t k = b; k < ev.size(); ++k) if (ev[k].is_transport && ev[k].transport.kind == TransportKind::Message && ev[k].transport.res.ok) {
        got = true; CHECK(ev[k].transport.res.msg.text == TEST_MESSAGE); } }
    CHECK(got); CHECK(rx.have_force_eom);
    rx.input_sample_counter = rx.force_eom_at_sample - 3 * (uint64_t)rx.input_rate;
    // flush(): first message within 4 s of zeros
    bool eom = false; ev.clear();
    for (uint32_t z = 0; z < rx.input_rate * 4 && !eom; ++z) { rx.process_sample(0.0f, ev);
      for (auto& e : ev) if (e.is_transport && e.transport.kind == TransportKind::Message) { eom = e.transport.res.ok && !e.transport.res.msg.is_som; }
      if (!eom) ev.clear(); }
    CHECK(eom); }
}

int main() {
  test_dcblock(); test_agc(); test_filter(); test_waveform(); test_demod(); test_symsync(); test_codesquelch();
  test_equalize(); test_framing(); test_combiner(); test_message(); test_assembler(); test_receiver();
  if (g_fail) { fprintf(stderr, "%d of %d checks FAILED\n", g_fail, g_checks); return 1; }
  printf("OK %d\n", g_checks);
  return 0;
}
