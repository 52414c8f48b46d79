/* same_synth.h — device-side generator of the synthetic SAME corpus used by bench.py and the GPU tests
 * (BASELINE.md §3 configs 3-5).  Tooling, not part of the receiver path: it only produces s16 input.
 *
 * Each stream is AWGN over its whole length plus continuous-phase AFSK bursts (mark 2083.3 Hz / space 1562.5 Hz,
 * 520.83 Bd with a FRACTIONAL number of samples per symbol, LSb first — waveform.rs:6-26; the reference's own test
 * modulator waveform.rs:73-104 uses an integer 42 samples/symbol) with a per-stream frequency offset.
 * Noise is Philox4x32-10 keyed by the per-stream seed (counter = sample index / 4) + Box-Muller.
 */
#ifndef SAME_SYNTH_H
#define SAME_SYNTH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct same_synth_burst {
  double start_sample;   /* first sample of the burst (fractional allowed) */
  uint32_t byte_offset;  /* into `bytes` */
  uint32_t n_bytes;      /* preamble + payload */
} same_synth_burst;

/* Writes samples [first_sample, first_sample + n_samples) of each of n_streams streams to d_out[stream * stride + k],
 * k = 0 .. n_samples-1 (device memory on `device`).  first_sample must be a multiple of 8; a stream generated in
 * several windows is bit-identical to the same stream generated at once (time-chunked config 4, 24 h config 5).  burst_begin has n_streams+1 entries (CSR into `bursts`).  All table pointers are HOST memory.
 * Returns 0 or a CUDA error code; `err_text` (>= 256 bytes, may be NULL) receives the message. */
int same_synth_generate(int device, int16_t* d_out, uint32_t n_streams, uint64_t stride, uint64_t first_sample,
                        uint32_t n_samples, uint32_t rate, const uint32_t* burst_begin, const same_synth_burst* bursts, uint32_t n_bursts,
                        const uint8_t* bytes, uint64_t n_bytes_total, const float* freq_offset_hz,
                        const uint32_t* seeds, float amplitude, float noise_sigma, char* err_text);

#ifdef __cplusplus
}
#endif
#endif
