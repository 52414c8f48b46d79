// C++ host-layer test: reads like the reference's own receiver tests (crates/sameold/src/receiver.rs:641-705) and its
// sample/test.sh, against the C++ mirror of the API (include/same_receiver.hpp) over libsame_b200.so.
// usage: test_receiver <long_message.bin> <npt.bin> <two_and_two.bin>     (raw s16le files)
// Without a GPU it must fail loudly with SAME_ERR_NO_DEVICE (exit code 3), never fall back.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#include "same_receiver.hpp"

static std::vector<int16_t> read_s16(const char* path) {
  std::ifstream f(path, std::ios::binary);
  std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  std::vector<int16_t> out(raw.size() / 2);
  std::memcpy(out.data(), raw.data(), out.size() * 2);
  return out;
}
#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main(int argc, char** argv) {
  if (argc != 4) { std::fprintf(stderr, "usage: %s long.bin npt.bin two.bin\n", argv[0]); return 2; }
  try {
    auto builder = same::SameReceiverBuilder::samedec(22050);
    CHECK(builder.input_rate() == 22050);
    CHECK(builder.with_timing_max_deviation(0.9f).timing_max_deviation() == 0.5f);   // setter clamps (builder.rs:162)
    builder.with_timing_max_deviation(0.01f);

    // batched: the three sample/ recordings as one ragged batch == what samedec prints for each
    std::vector<std::vector<int16_t>> recs = {read_s16(argv[1]), read_s16(argv[2]), read_s16(argv[3])};
    auto rx = builder.build_batch(3);
    auto lines = rx.decode_samedec(recs);
    CHECK(lines[0].size() == 1 && lines[0][0].rfind("ZCZC-EAS-DMO-372088-091724-919623", 0) == 0 && lines[0][0].size() == 252);
    CHECK(lines[1].size() == 1 && lines[1][0] == "ZCZC-PEP-NPT-000000+0030-2771820-TEST    -");
    CHECK(lines[2].size() == 2 && lines[2][0] == "NNNN" &&
          lines[2][1] == "ZCZC-WXR-SVR-012079-013019-013027-013075-013185-013173+0130-0462024-N0C4LL  -");

    // single receiver: event order of the first burst (receiver.rs:651-671): Searching, Reading, Burst, Assembling, NoCarrier
    auto one = builder.build();
    auto evs = one.iter_events(std::vector<int16_t>(recs[1].begin(), recs[1].begin() + 30000));
    CHECK(evs.size() == 5);
    const uint32_t want[5] = {SAME_EV_LINK_SEARCHING, SAME_EV_LINK_READING, SAME_EV_LINK_BURST, SAME_EV_TR_ASSEMBLING, SAME_EV_LINK_NOCARRIER};
    for (int i = 0; i < 5; ++i) CHECK(evs[i].kind == want[i]);
    CHECK(evs[2].burst() && std::string(evs[2].data.begin(), evs[2].data.end()).rfind("ZCZC-PEP-NPT-000000+0030-2771820-TEST    -", 0) == 0);
    // rest of the recording: the header is released 682 symbols after the third burst, still inside the file
    auto msgs = one.iter_messages(std::vector<int16_t>(recs[1].begin() + 30000, recs[1].end()));
    CHECK(msgs.size() == 1 && msgs[0].is_start && msgs[0].text == lines[1][0] && msgs[0].voting_byte_count == 42 &&
          msgs[0].parity_error_count == 0);
    CHECK(!one.flush());
    one.reset();
    CHECK(one.input_sample_counter() == 0);
    // long_message is cut close: iter_messages sees nothing, flush() releases the header and stops there
    CHECK(one.iter_messages(recs[0]).empty());
    auto m = one.flush();
    CHECK(m && m->is_start && m->text == lines[0][0] && m->voting_byte_count == 252 && m->parity_error_count == 0);
    CHECK(one.input_sample_counter() > recs[0].size() && one.input_sample_counter() < recs[0].size() + 4 * 22050);
    CHECK(!one.flush());
    // the reference's own item type: f32 PCM normalised to [-1, 1) (lib.rs:78-79), library-default builder (AGC limits
    // [0, 1e6], builder.rs:55) -- the same header comes out
    {
      auto fb = same::SameReceiverBuilder(22050);
      std::vector<float> f(recs[1].size());
      for (size_t i = 0; i < f.size(); ++i) f[i] = (float)recs[1][i] / 32768.0f;
      auto fr = fb.build();
      auto fm = fr.iter_messages(f);
      CHECK(fm.size() == 1 && fm[0].is_start && fm[0].text == lines[1][0]);
      CHECK(fr.input_sample_counter() == f.size());
    }
    std::printf("CPP_OK\n");
    return 0;
  } catch (const same::EngineError& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return e.code == SAME_ERR_NO_DEVICE ? 3 : 4;
  }
}
