// samedec_oracle.cpp — CLI over the CPU oracle mirroring `samedec --rate R --file F` stdout (TEST INFRASTRUCTURE ONLY).
// Prints one line per decoded message (crates/samedec/src/app.rs:137-139); with -v also every link/transport event.
#include "same_oracle.hpp"

#include <cstdio>
#include <cstdlib>

using namespace same_oracle;

int main(int argc, char** argv) {
  uint32_t rate = 22050; const char* file = nullptr; bool verbose = false; bool library_defaults = false;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--rate" && i + 1 < argc) rate = (uint32_t)atoi(argv[++i]);
    else if (a == "--file" && i + 1 < argc) file = argv[++i];
    else if (a == "-v") verbose = true;
    else if (a == "--library-defaults") library_defaults = true;
  }
  if (!file) { fprintf(stderr, "usage: %s [--rate R] --file F.s16le.bin [-v]\n", argv[0]); return 2; }
  FILE* f = fopen(file, "rb");
  if (!f) { perror(file); return 1; }
  std::vector<int16_t> s;
  int16_t buf[4096]; size_t n;
  while ((n = fread(buf, 2, 4096, f)) > 0) s.insert(s.end(), buf, buf + n);
  fclose(f);
  Config c = library_defaults ? Config() : Config::samedec(rate);
  c.input_rate = rate;
  SameReceiver rx(c);
  std::vector<Event> ev;
  for (int16_t v : s) rx.process_sample((float)v, ev);
  samedec_eof_flush(rx, ev);
  for (auto& e : ev) {
    if (verbose) {
      if (!e.is_transport) {
        static const char* nm[] = {"NoCarrier", "Searching", "Reading", "Burst"};
        fprintf(stderr, "[%10llu sym %6llu] link %s", (unsigned long long)e.input_sample_counter,
                (unsigned long long)e.symbol_count, nm[(int)e.link.kind]);
        if (e.link.kind == LinkKind::Burst) {
          fprintf(stderr, " (%zu B) \"", e.link.burst.size());
          for (uint8_t b : e.link.burst) fputc((b >= 32 && b < 127) ? b : '.', stderr);
          fputc('"', stderr);
        }
        fputc('\n', stderr);
      } else {
        static const char* nm[] = {"Idle", "Assembling", "Message"};
        fprintf(stderr, "[%10llu sym %6llu] transport %s", (unsigned long long)e.input_sample_counter,
                (unsigned long long)e.symbol_count, nm[(int)e.transport.kind]);
        if (e.transport.kind == TransportKind::Message) {
          if (e.transport.res.ok)
            fprintf(stderr, " ok voting=%zu parity=%zu \"%s\"", e.transport.res.msg.voting_byte_count,
                    e.transport.res.msg.parity_error_count, e.transport.res.msg.text.c_str());
          else fprintf(stderr, " err=%u", (unsigned)e.transport.res.err);
        }
        fputc('\n', stderr);
      }
    }
    if (e.is_transport && e.transport.kind == TransportKind::Message && e.transport.res.ok)
      printf("%s\n", e.transport.res.msg.text.c_str());
  }
  return 0;
}
