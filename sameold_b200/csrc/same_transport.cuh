// same_transport.cuh — event-rate device code: event/payload arenas, the byte Framer and the transport layer
// (Assembler + combiner + header validation).  Runs at burst/message rate (a handful of times per burst), one lane
// per stream, on the byte-oriented StreamBlob in global memory.  Integer/byte work only.
//
// Restates (does not copy) the behaviour of:
//   crates/sameold/src/receiver/framing.rs:109-197,235-243        Framer
//   crates/sameold/src/receiver/combiner.rs:32-80,105-137,154-271 combine / estimate_message / bit votes
//   crates/sameold/src/receiver/assembler.rs:154-234,245-371      Assembler / PendingResult / pruning
//   crates/sameplace/src/message.rs:718-736,813-828               Message::try_from / check_header
#pragma once

#include "same_params.h"

namespace same_dev {

// ----------------------------------------------------------------------------------------------------------------
// Event arena
// ----------------------------------------------------------------------------------------------------------------
struct EvCtx {
  uint32_t stream;
  uint32_t seq;   // per-stream sequence (state)
};

// Word-wise copy of `nbytes` (rounded up to whole words: every byte buffer in StreamBlob and the payload arena is
// 4-byte aligned and sized in whole words).  Loads of a group are issued before its stores so that their latencies
// overlap — this code runs on one lane while 31 others wait.
__device__ __forceinline__ void copy_words(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t nbytes) {
  const uint32_t nw = (nbytes + 3u) >> 2;
  uint32_t* __restrict__ d = reinterpret_cast<uint32_t*>(dst);
  const uint32_t* __restrict__ q = reinterpret_cast<const uint32_t*>(src);
  uint32_t i = 0;
  for (; i + 8 <= nw; i += 8) {
    uint32_t t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = q[i + k];
#pragma unroll
    for (int k = 0; k < 8; ++k) d[i + k] = t[k];
  }
  for (; i < nw; ++i) d[i] = q[i];
}

// Everything is passed by value: a by-reference argument to a noinline function would force the caller's lane state
// out of registers into local memory for the whole kernel.
static __device__ __noinline__ void emit_event_impl(const SameParams& p, uint32_t stream, uint32_t seq, uint32_t kind,
                                             uint32_t err, unsigned long long n, unsigned long long symcount,
                                             const uint8_t* data, uint32_t data_len, uint32_t copy_len, uint32_t parity,
                                             uint32_t voting, uint32_t flags) {
  unsigned int idx = atomicAdd(&p.counters[0], 1u);
  uint32_t off = 0;
  if (copy_len) {
    const uint32_t padded = (copy_len + 3u) & ~3u;
    off = atomicAdd(&p.counters[1], padded);   // offsets stay multiples of 4
    if ((unsigned long long)off + padded <= p.payload_cap) {
      if ((reinterpret_cast<uintptr_t>(data) & 3u) == 0) copy_words(p.payload + off, data, copy_len);
      else for (uint32_t i = 0; i < copy_len; ++i) p.payload[off + i] = data[i];
    } else {
      // payload arena full: the event is still delivered, without bytes, and says so (never an offset past the arena)
      atomicAdd(&p.counters[2], 1u);
      off = 0; data_len = 0; flags |= SAME_EV_FLAG_PAYLOAD_LOST;
    }
  }
  if (idx < p.events_cap) {
    same_event e;
    e.stream = stream; e.seq = seq; e.input_sample_counter = n; e.symbol_count = symcount;
    e.kind = kind; e.err = err; e.data_offset = off; e.data_len = data_len;
    e.parity_errors = (uint16_t)parity; e.voting_bytes = (uint16_t)voting; e.flags = flags;
    p.events[idx] = e;
  }
}

__device__ __forceinline__ void emit_event(const SameParams& p, EvCtx& c, uint32_t kind, uint32_t err,
                                           unsigned long long n, unsigned long long symcount, const uint8_t* data,
                                           uint32_t data_len, uint32_t copy_len, uint32_t parity, uint32_t voting,
                                           uint32_t flags) {
  emit_event_impl(p, c.stream, c.seq, kind, err, n, symcount, data, data_len, copy_len, parity, voting, flags);
  c.seq += 1;
}

// ----------------------------------------------------------------------------------------------------------------
// combiner.rs:105-137
// ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_allowed_byte(uint32_t c) {
  return c == '-' || (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '/' ||
         c == '?' || c == '(' || c == ')' || c == '[' || c == ']' || c == '.' || c == '_' || c == ',' || c == '+' ||
         c == ' ';
}

// ----------------------------------------------------------------------------------------------------------------
// Framer (framing.rs).  State lives in registers (fr_state/word/count/invalid/msglen); bytes in blob->burst.
// Link kinds: 0 NoCarrier, 1 Searching, 2 Reading, 3 Burst.
// ----------------------------------------------------------------------------------------------------------------
struct Framer {
  uint32_t st;       // 0 Idle, 1 PrefixSearch, 2 DataRead
  uint32_t word, count, invalid, msglen;
};

__device__ __forceinline__ uint32_t framer_state(const Framer& f) { return f.st; }  // framing.rs:191-197 (same numbering)

// framing.rs:174-186.  Returns 3 (Burst, bytes = blob->burst[0..msglen)) or 0; `burst_len` receives the length.
__device__ __forceinline__ uint32_t framer_end(Framer& f, uint32_t& burst_len) {
  uint32_t out = 0;
  if (f.st == 2) { out = 3; burst_len = f.msglen; }
  f.st = 0;
  return out;
}

__device__ __forceinline__ uint32_t prefix_errors(uint32_t w) {  // framing.rs:235-243
  return min(__popc(w ^ 0x5A435A43u), __popc(w ^ 0x4E4E4E4Eu));
}

// framing.rs:109-164 without the restart branch
__device__ __forceinline__ uint32_t framer_input_norestart(const SameParams& p, StreamBlob* blob, Framer& f,
                                                           uint32_t data, uint32_t& burst_len) {
  if (f.st == 0) return 0;
  if (f.st == 1) {
    f.word = (f.word << 8) | data;
    f.count += 1;
    if (prefix_errors(f.word) <= p.fr_max_prefix_err) {
      blob->burst[0] = (uint8_t)(f.word >> 24); blob->burst[1] = (uint8_t)(f.word >> 16);
      blob->burst[2] = (uint8_t)(f.word >> 8);  blob->burst[3] = (uint8_t)f.word;
      f.msglen = 4; f.invalid = 0; f.st = 2;
    } else if (f.count > 21u) {  // PREFIX_SEARCH_LEN framing.rs:201
      f.st = 0;
    }
    return f.st;
  }
  f.invalid += is_allowed_byte(data) ? 0u : 1u;
  if (f.invalid > p.fr_max_invalid) return framer_end(f, burst_len);
  if (f.msglen < SAME_BURST_CAP) blob->burst[f.msglen] = (uint8_t)data;
  f.msglen += 1;
  return 2;
}

// framing.rs:109-123: restart ends the frame in progress (possibly emitting it), then starts a prefix search and
// feeds the byte.  NOTE: when the old frame is emitted its bytes must be consumed by the caller before the next
// DataRead begins; a new DataRead cannot begin on this very byte (one byte never matches a 4-byte prefix with <=7 errors).
__device__ __forceinline__ uint32_t framer_input(const SameParams& p, StreamBlob* blob, Framer& f, uint32_t data,
                                                 bool restart, uint32_t& burst_len) {
  if (restart) {
    uint32_t out = framer_end(f, burst_len);
    f.st = 1; f.word = 0; f.count = 0;
    uint32_t dummy = 0;
    (void)framer_input_norestart(p, blob, f, data, dummy);
    return out == 3 ? 3u : 1u;
  }
  return framer_input_norestart(p, blob, f, data, burst_len);
}

// ----------------------------------------------------------------------------------------------------------------
// Header validation — sameplace message.rs:813-828
//   ^ZCZC-[[:alpha:]]{3}-[[:alpha:]]{3}(-[0-9]{6})+(\+[0-9]{4}-[0-9]{7}-.{3,8}-)     (leftmost-first / greedy)
// ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_alpha(uint32_t c) { return (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'); }
__device__ __forceinline__ bool is_digit(uint32_t c) { return c >= '0' && c <= '9'; }

static __device__ __noinline__ bool check_header(const uint8_t* h, uint32_t n, uint32_t& off_time, uint32_t& hdr_len) {
  if (n < 13) return false;
  if (!(h[0] == 'Z' && h[1] == 'C' && h[2] == 'Z' && h[3] == 'C' && h[4] == '-')) return false;
  if (!(is_alpha(h[5]) && is_alpha(h[6]) && is_alpha(h[7]) && h[8] == '-' && is_alpha(h[9]) && is_alpha(h[10]) &&
        is_alpha(h[11])))
    return false;
  // greedy location groups
  uint32_t q = 12, ngroups = 0;
  while (q + 7 <= n && h[q] == '-') {
    bool ok = true;
    for (uint32_t k = 1; k <= 6; ++k) ok = ok && is_digit(h[q + k]);
    if (!ok) break;
    q += 7; ngroups++;
  }
  // try the longest run of groups first, then backtrack
  for (; ngroups >= 1; --ngroups, q -= 7) {
    uint32_t r = q;
    if (r >= n || h[r] != '+') continue;
    r++;
    if (r + 4 + 1 + 7 + 1 > n) continue;
    bool ok = true;
    for (uint32_t k = 0; k < 4; ++k) ok = ok && is_digit(h[r + k]);
    ok = ok && h[r + 4] == '-';
    for (uint32_t k = 0; k < 7; ++k) ok = ok && is_digit(h[r + 5 + k]);
    ok = ok && h[r + 12] == '-';
    if (!ok) continue;
    r += 13;
    for (int len = 8; len >= 3; --len) {
      if (r + (uint32_t)len >= n) continue;
      bool nl = false;
      for (int k = 0; k < len; ++k) nl = nl || (h[r + k] == '\n');
      if (nl) continue;
      if (h[r + len] == '-') { off_time = q; hdr_len = r + (uint32_t)len + 1; return true; }
    }
  }
  return false;
}

// ----------------------------------------------------------------------------------------------------------------
// Transport state held in registers between symbols
// ----------------------------------------------------------------------------------------------------------------
struct Transport {
  uint32_t hist_n;            // bursts in history
  bool pending, have_prev, have_eom;
  uint32_t tr_state;          // last reported: 0 Idle, 1 Assembling, 2 Message
  unsigned long long next_deadline;  // min(pending deadline, history deadlines); ~0 if none
  unsigned long long eom_at;  // force_eom_at_sample
};

// result of combine / poll
struct MsgResult {
  uint32_t kind;   // 0 SOM, 1 EOM, 2 Err
  uint32_t err;    // 1 UnrecognizedPrefix, 2 NotAscii, 3 Malformed
  uint32_t len, parity, voting, offset;
};

__device__ __forceinline__ void recompute_next_deadline(const StreamBlob* b, Transport& t) {
  unsigned long long d = ~0ull;
  if (t.pending) d = b->pending_deadline;
  for (uint32_t k = 0; k < t.hist_n; ++k) d = min(d, b->hist[k].deadline);
  t.next_deadline = d;
}

// assembler.rs:357-363
static __device__ __noinline__ void prune_history(StreamBlob* b, Transport& t, unsigned long long now) {
  uint32_t w = 0;
  for (uint32_t k = 0; k < t.hist_n; ++k) {
    if (b->hist[k].deadline <= now) continue;
    if (w != k) {
      b->hist[w].deadline = b->hist[k].deadline; b->hist[w].len = b->hist[k].len;
      copy_words(b->hist[w].data, b->hist[k].data, b->hist[k].len);
    }
    ++w;
  }
  t.hist_n = w;
  while (t.hist_n > 2) {  // pop_front
    for (uint32_t k = 0; k + 1 < t.hist_n; ++k) {
      b->hist[k].deadline = b->hist[k + 1].deadline; b->hist[k].len = b->hist[k + 1].len;
      copy_words(b->hist[k].data, b->hist[k + 1].data, b->hist[k + 1].len);
    }
    t.hist_n -= 1;
  }
}

// combiner.rs:154-203 + 32-80.  Returns false for `None`.  The estimate is left in b->est[0..good_len).
// The three bursts are read one 32-bit word (4 message bytes) at a time and the estimate / burst-count / bit-error
// arrays are written one word at a time; the per-byte voting logic itself is the reference's, on register values.
static __device__ __noinline__ bool combine(StreamBlob* b, const Transport& t, MsgResult& res) {
  const uint32_t nb = min(t.hist_n, 3u);
  const uint32_t l0 = nb > 0 ? b->hist[0].len : 0u, l1 = nb > 1 ? b->hist[1].len : 0u, l2 = nb > 2 ? b->hist[2].len : 0u;
  const uint32_t* w0 = reinterpret_cast<const uint32_t*>(b->hist[0].data);
  const uint32_t* w1 = reinterpret_cast<const uint32_t*>(b->hist[1].data);
  const uint32_t* w2 = reinterpret_cast<const uint32_t*>(b->hist[2].data);
  uint32_t* est_w = reinterpret_cast<uint32_t*>(b->est);
  uint32_t* nb_w = reinterpret_cast<uint32_t*>(b->est_nb);
  uint32_t* er_w = reinterpret_cast<uint32_t*>(b->est_err);
  uint32_t n = 0;
  bool stop = false;
  for (uint32_t w = 0; w < SAME_MAX_MESSAGE_LENGTH / 4 && !stop; ++w) {
    const uint32_t base = 4u * w;
    const uint32_t x0 = base < l0 ? w0[w] : 0u, x1 = base < l1 ? w1[w] : 0u, x2 = base < l2 ? w2[w] : 0u;
    uint32_t oe = 0, on = 0, orr = 0;
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) {
      if (stop) break;
      const uint32_t i = base + j;
      // every burst iterator advances in lock step (one byte per output byte); exhausted bursts drop out
      const bool h0 = i < l0, h1 = i < l1, h2 = i < l2;
      const uint32_t nc = (h0 ? 1u : 0u) + (h1 ? 1u : 0u) + (h2 ? 1u : 0u);
      if (nc == 0) { stop = true; break; }
      const uint32_t r0 = (x0 >> (8 * j)) & 0xffu, r1 = (x1 >> (8 * j)) & 0xffu, r2 = (x2 >> (8 * j)) & 0xffu;
      // the (up to three) present bytes in burst order
      const uint32_t c0 = h0 ? r0 : (h1 ? r1 : r2);
      const uint32_t c1 = h0 ? (h1 ? r1 : r2) : r2;
      const uint32_t c2 = r2;
      const bool msb = ((h0 ? r0 : 0u) | (h1 ? r1 : 0u) | (h2 ? r2 : 0u)) & 0x80u;   // SAME bytes never have the MSb set
      const uint32_t a0 = c0 & 0x7fu, a1 = c1 & 0x7fu, a2 = c2 & 0x7fu;
      uint32_t est, be;
      if (nc == 1) { est = a0; be = 0; }
      else if (nc == 2) {  // bit_vote_detect combiner.rs:216-222
        const uint32_t x = a0 ^ a1;
        est = x ? 0u : a0; be = __popc(x);
      } else {             // bit_vote_correct combiner.rs:234-249
        const uint32_t p0 = ~(a0 ^ a1) & 0xffu, p1 = ~(a1 ^ a2) & 0xffu, p2 = ~(a0 ^ a2) & 0xffu;
        est = (a0 & p0) | (a2 & p1) | (a2 & p2);
        be = 8u - __popc(p0 & p1 & p2);
      }
      if (!is_allowed_byte(est)) { stop = true; break; }
      oe |= est << (8 * j); on |= nc << (8 * j); orr |= (be + (msb ? 1u : 0u)) << (8 * j);
      ++n;
    }
    est_w[w] = oe; nb_w[w] = on; er_w[w] = orr;
  }
  if (n == 0) return false;
  uint32_t good = 0;  // truncate_bytes_with_reference(msg, burst_count, 2)  combiner.rs:262-271
  while (good < n && b->est_nb[good] >= 2) ++good;
  // Message::try_from((good_msg, errs, bursts))  message.rs:718-736 (bytes are 7-bit, so from_utf8 cannot fail)
  const uint8_t* m = b->est;
  bool is_start = good >= 5 && m[0] == 'Z' && m[1] == 'C' && m[2] == 'Z' && m[3] == 'C' && m[4] == '-';
  uint32_t err;
  if (is_start) {
    uint32_t off, hl;
    if (check_header(m, good, off, hl)) {
      uint32_t pe = 0, vc = 0;  // new_with_errors / new_with_error_info  message.rs:209-254
      for (uint32_t i = 0; i < hl; ++i) { pe += b->est_err[i]; vc += (b->est_nb[i] >= 3) ? 1u : 0u; }
      res.kind = 0; res.err = 0; res.len = hl; res.parity = pe; res.voting = vc; res.offset = off;
      return true;
    }
    err = 3;  // Malformed
  } else if (good >= 2 && m[0] == 'N' && m[1] == 'N') {
    res.kind = 1; res.err = 0; res.len = 4; res.parity = 0; res.voting = 0; res.offset = 0;
    return true;
  } else {
    err = 1;  // UnrecognizedPrefix
  }
  if (n >= 2 && m[0] == 'N' && m[1] == 'N') {  // Fast EOM on the un-truncated estimate  combiner.rs:63-66,251-258
    res.kind = 1; res.err = 0; res.len = 4; res.parity = 0; res.voting = 0; res.offset = 0;
    return true;
  }
  if (good == 0) return false;
  res.kind = 2; res.err = err; res.len = 0; res.parity = 0; res.voting = 0; res.offset = 0;
  return true;
}

// PendingResult::accept  assembler.rs:294-331.  For SOM the text is b->est[0..len).
static __device__ __noinline__ void pending_accept(const SameParams& p, StreamBlob* b, Transport& t, const MsgResult& r,
                                            unsigned long long now) {
  unsigned long long dl = (r.kind == 1) ? now : now + p.interburst_symbols;
  bool store;
  if (t.pending) {
    if (b->pending_kind == 2) store = true;                                  // (Err(_), _)
    else if (b->pending_kind == 1 && r.kind == 0) store = true;              // (Ok(EOM), Ok(SOM))
    else if (b->pending_kind == 0 && r.kind == 0) store = r.voting >= b->pending_voting;  // (Ok(SOM), Ok(SOM))
    else store = false;
  } else store = true;
  if (!store) return;
  t.pending = true;
  b->pending_deadline = dl; b->pending_kind = (uint8_t)r.kind; b->pending_err = (uint8_t)r.err;
  b->pending_len = (uint16_t)r.len; b->pending_parity = (uint16_t)r.parity; b->pending_voting = (uint16_t)r.voting;
  b->pending_offset = (uint16_t)r.offset;
  if (r.kind == 0) copy_words(b->pending_text, b->est, r.len);
  else if (r.kind == 1) { b->pending_text[0] = 'N'; b->pending_text[1] = 'N'; b->pending_text[2] = 'N'; b->pending_text[3] = 'N'; }
}

// Assembler::idle  assembler.rs:205-234.  Returns transport kind (0 Idle, 1 Assembling, 2 Message); for Message the
// result is in `out` and its text in b->pending_text.
static __device__ __noinline__ uint32_t assembler_idle(const SameParams& p, StreamBlob* b, Transport& t,
                                                unsigned long long now, MsgResult& out) {
  prune_history(b, t, now);
  uint32_t kind;
  if (t.pending && b->pending_deadline <= now) {  // PendingResult::poll assembler.rs:339-348
    t.pending = false;
    out.kind = b->pending_kind; out.err = b->pending_err; out.len = b->pending_len; out.parity = b->pending_parity;
    out.voting = b->pending_voting; out.offset = b->pending_offset;
    if (out.kind != 2) {  // Some(Ok(msg)): remember for duplicate suppression
      t.have_prev = true;
      b->prev_deadline = now + p.history_symbols;
      b->prev_len = (uint16_t)out.len;
      copy_words(b->prev_text, b->pending_text, out.len);
    }
    kind = 2;
  } else kind = t.hist_n ? 1u : 0u;
  recompute_next_deadline(b, t);
  return kind;
}

// Assembler::assemble  assembler.rs:154-184.  `burst` = blob->burst[0..min(len,CAP)).
static __device__ __noinline__ uint32_t assembler_assemble(const SameParams& p, StreamBlob* b, Transport& t, uint32_t burst_len,
                                                    unsigned long long now, MsgResult& out) {
  if (burst_len == 0) return assembler_idle(p, b, t, now, out);
  prune_history(b, t, now);
  if (t.have_prev && b->prev_deadline <= now) t.have_prev = false;  // prune_previous assembler.rs:366-371
  uint32_t n = min(burst_len, (uint32_t)SAME_MAX_MESSAGE_LENGTH);
  BurstSlot& slot = b->hist[t.hist_n];  // hist_n <= 2 after pruning
  slot.deadline = now + p.history_symbols; slot.len = n;
  copy_words(slot.data, b->burst, n);
  t.hist_n += 1;
  MsgResult r;
  if (combine(b, t, r)) {
    bool keep = true;  // deduplicate assembler.rs:245-265: string-equal to the previous message
    if (r.kind != 2 && t.have_prev) {
      const uint8_t* txt = (r.kind == 0) ? b->est : (const uint8_t*)"NNNN";
      bool same = (b->prev_len == r.len);
      for (uint32_t i = 0; same && i < r.len; ++i) same = (b->prev_text[i] == txt[i]);
      if (same) keep = false;
    }
    if (keep) pending_accept(p, b, t, r, now);
  }
  return assembler_idle(p, b, t, now, out);
}

}  // namespace same_dev
