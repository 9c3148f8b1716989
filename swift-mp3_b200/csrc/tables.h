// Host-side constant tables of the encode path.  Literal tables come from tables_gen.h (shared with the device
// code); the two gain tables are evaluated here in double like the reference does at run time
// (Sources/SwiftMP3/MP3Encoder.swift = SRC).
#pragma once
#include <cstdint>

namespace mp3b {

const float *host_inv_step();   // [256] 1 / Float(max(2^((g-210)/4), 1e-4)), SRC:798-800
const float *host_inv_step_iso();   // [320] ISO-mode quantizer scale per search gain (iso_mode.cuh)
const double *host_gain_thr();  // [256] 2^((g-210)/4): thresholds that replace log2 in computeGlobalGain (SRC:1004)
const int *host_sfb_cum();      // [3][21] cumulative long sfb widths for 44.1 / 48 / 32 kHz, SRC:1814-1820

const uint8_t *host_len15();   // [256] Huffman table 15 code lengths, SRC:2457-2473
const uint8_t *host_code15();  // [256] Huffman table 15 codes, SRC:2476-2493

// MP3Tables lookups, SRC:2509-2556
int bitrate_index(int bitrate, int sample_rate);
int bitrate_value(int index);
int sample_rate_index(int sample_rate);
int sfb_table_index(int sample_rate);  // 0: 44.1k (default), 1: 48k, 2: 32k  (SRC:1879-1888)

}  // namespace mp3b
