// Device pipeline of the MP3 encode path: data layout + kernel launchers (implemented in kernels.cu).
//
// Unit of work: one granule-channel (gc) = 576 PCM samples of one channel.  A "pass" processes up to Fc frames of
// every stream of a batch.  All per-pass arrays are stream-major.  Frame slot (s, r): r = 0 is the frame carried
// from the previous pass (the reference's `bufferedFrame`, SRC:246), r = 1 + f the f-th frame of this pass.
// gc slot (s, g * ch + c) with g = 2 * f + gr — the reference's encode order, gr-major / channel-minor (SRC:652-653).
// SRC = Sources/SwiftMP3/MP3Encoder.swift of the reference.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace mp3b {

constexpr int kMaxEntries = 20;     // gain-loop iterations, SRC:745
constexpr int kMdCarryCap = 8192;   // bytes of reservoir backlog that may cross a pass boundary
constexpr int kPsyMaxPart = 80;     // threshold-calculation partitions of the psychoacoustic model (ISO mode level 2)

// Tables of the psychoacoustic model (iso_psy.cuh) for one sample rate; built on the host (engine.cc: build_psy_tab), device resident.
struct PsyTab {
  int n_part;
  int sfb_line[23];                          // FFT-line boundaries of the 22 long scalefactor bands (MDCT line * 8 / 9)
  uint16_t part_lo[kPsyMaxPart], part_n[kPsyMaxPart];   // first FFT line and number of lines of a partition (1/3 Bark each)
  float rnorm[kPsyMaxPart], minval[kPsyMaxPart] /* dB */, qthr[kPsyMaxPart] /* absolute threshold, energy */;
  float s3t[kPsyMaxPart * kPsyMaxPart];      // spreading function, [source][target]
  uint8_t line_part[512];
  float hann1024[1024], hann256[256];
  float2 tw[768];                            // exp(-2 pi i j / 1024)
};

struct Config {              // constant for the lifetime of a batch; passed to kernels by value
  int n_streams, channels, fsc /* floats per frame = 1152*channels */;
  int sample_rate, base_kbps, vbr, mode, quality, crc, original, copyright;
  int sr_index, sfb_index, side_bytes, header_bytes /* 4 + crc + side */;
  int mode_bits, mode_ext, cbr_index;
  float f_one, f_neg0;                // 1.0f and -0.0f as run-time values (see k_spectrum phase A)
  int iso;                            // opt-in ISO mode (iso_mode.cuh): 1 = ISO quantizer, table selection, count1, real main_data_begin; 2 = + psychoacoustic model and scalefactor outer loop (iso_psy.cuh); 3 = + window switching (start / short / stop blocks)
  int iso_delay;                      // ISO mode level 3 (window switching): the filterbank reads the PCM this many samples late (576), 0 otherwise
  float ms_scale;                     // mid / side = (L +- R) * ms_scale: 0.5 like the reference (SRC:2148-2154), 1/sqrt(2) in ISO mode
  int frame_base[16], frame_rem[16];  // 144*kbps*1000 / sr and % sr per bitrate index
  uint8_t vbr_idx_of_kbps[324];       // bitrateIndex(kbps) for every VBR target 0...320
};

struct StreamPlan {          // per stream, per pass (host -> device)
  const float *cur;          // new PCM of this pass (device pointer), interleaved
  uint32_t cur_n;            // floats at cur
  uint32_t n_frames;         // frames encoded in this pass (including a final zero-padded one)
  uint32_t flags;            // bit0: last frame of the pass is the flush frame (isFinal, SRC:331)
                             // bit1: emit the buffered frame at the end of the pass (flush, SRC:335-347)
                             // bit2: first pass of an API call (output cursor restarts at 0)
  uint32_t head_n;           // floats valid in head: fsc (carried frame, zeros before the first) + pending partial
};

struct GcSide {              // side-info fields of one gc (GranuleInfo, SRC:2070-2085) + trace extras
  uint16_t part23;           // Huffman bit count before the 12-bit mask of the side-info writer
  uint16_t big_values;
  uint8_t global_gain, gain_used;
  uint8_t block_type;        // 0 long, 1 mixed, 2 short (raw values SRC:1923-1927)
  uint8_t sbg[3];
  uint8_t region0, region1, preflag, g0, iterations, pad;   // pad: ISO mode, search gain above 255 (gain_used + pad = the gain that quantized)
  uint16_t max_bits;
  uint8_t sfc, part2;        // ISO mode level 2: scalefac_compress and the scalefactor bits inside part23 (0 otherwise)
  float energy;
  uint8_t tsel[3], c1sel;    // ISO mode: table_select per region, count1table_select (the reference writes 15, 15, 15 and 0)
};
struct FrameRec {            // everything needed to emit a frame later (one-frame delay, SRC:546-562)
  uint8_t valid, br_index, padding, ms;
  uint16_t mdb, slot;        // main_data_begin as written; slot = main-data bytes of the frame
  uint8_t is_final, pad0[3];
  int32_t reservoir_bits, huff_bytes;
  float frame_energy;
  GcSide gc[4];
};
struct FrameEmit {           // emission of a frame slot in this pass: copy `take` bytes from md + src_off, zero-fill up to slot
  uint32_t emit, src_off, take, out_off;
};

struct StreamState {         // persistent per stream
  int32_t avail_bytes;       // BitReservoir.availableBytes, SRC:2096
  int32_t backlog;           // BitReservoir.stream.count, SRC:2093 (bytes waiting in md_carry)
  int32_t pad_rem;           // paddingRemainder, SRC:247
  uint32_t frame_count, total_bytes;   // SRC:256-257
  uint32_t out_pos;          // bytes written to this stream's output region during the current API call
  uint32_t frames_total;     // frames encoded so far
  int32_t vbr_n;             // valid entries of the 10-deep energy history, SRC:1141
  int32_t error;             // sticky engine-limit flag
  uint32_t ms_prev;          // stereo decision of the carried frame
  float vbr_hist[10];        // oldest first
  FrameRec buffered;         // bufferedFrame, SRC:246
};

struct PassBuffers {         // device arrays for one pass; Fc = frame capacity per stream, GC = Fc*2*ch
  int Fc, GC;
  int max_frames;            // largest n_frames of any stream in this pass (grid sizing)
  StreamPlan *plan;          // [S]
  StreamState *state;        // [S]
  float *head_in, *head_out; // [S][2*fsc] carried frame + pending partial (double buffered, swapped per pass)
  uint8_t *ms;               // [S][Fc+1]   ([0] = carried frame)
  float *frame_energy;       // [S][Fc]
  float *gc_energy;          // [S][10 + GC]  (first 10 = carried history, right aligned)
  uint16_t *gc_bt;           // [S][GC] block_type | sbg0<<2 | sbg1<<5 | sbg2<<8
  uint8_t *frame_br;         // [S][Fc] bitrate index
  float *sub;                // [S][ch][sub_rows][32] subband samples (K1 output): row 18 (g + 1) + t = step t of granule g,
                             // rows 0..17 = last granule of the previous pass (the MDCT overlap, SRC:1534-1535)
  int sub_rows;              // 18 * (1 + 2 * Fc)
  float *spec;               // [S][GC][576] MDCT spectrum — trace plane only (nullptr otherwise)
  float *smag;               // [S][GC][576] sign(x) * |x|^0.75 (K4 output, K5 input)
  uint32_t *gc_meta;         // [S][GC] g0 | n_entries<<8 | restart<<16 | preflag<<17   (ISO mode: first gain | n_entries<<8 | gain of entry 19<<24)
  uint16_t *gc_bits;         // [S][GC][20]
  uint16_t *gc_bv;           // [S][GC][20]
  uint32_t *gc_bitoff;       // [S][GC] bit offset inside the frame's main data
  uint32_t *gc_sel;          // [S][GC] gain_used | big_values<<8
  uint32_t *fr_md;           // [S][Fc][2] byte offset of the frame's main data in md, huff bytes
  FrameRec *rec;             // [S][Fc+1]
  FrameEmit *emit;           // [S][Fc+1]
  uint8_t *md;               // [S][md_stride] main-data byte stream of the pass (starts with the carried backlog)
  size_t md_stride;
  uint32_t *md_tail;         // [S][2] offset / length of the unconsumed tail after the scan
  uint8_t *md_carry;         // [S][kMdCarryCap]
  uint8_t *out;              // [S][out_stride] emitted frames of the API call
  size_t out_stride;
  uint16_t *emit_size;       // [S][Fc+1] sizes of the frames emitted by this pass, in order
  uint32_t *emit_n;          // [S]
  const float *tc_b;         // non-null: k_filterbank_tc; the analysis matrix split into three TF32 terms, pre-swizzled (filterbank_tc.cuh)
  // ISO mode level 2 (iso_psy.cuh)
  const PsyTab *psy;
  float *gc_psy;             // [S][GC][24] threshold / energy per long scalefactor band (22), perceptual entropy, mean tonality
  uint8_t *gc_sf;            // [S][GC][24] scalefactors (21), scalefac_compress, part2 bits, bands over | outer iterations << 4
  // optional traces
  int32_t *tr_ix;            // same
  float *tr_thr;             // same
};

cudaError_t upload_tables();   // __constant__ tables for the current device

// launchers; each returns the number of kernels launched (negative cudaError on failure)
int launch_prepass(const Config &cfg, const PassBuffers &pb, cudaStream_t st);
int launch_spectrum(const Config &cfg, const PassBuffers &pb, cudaStream_t st);
int launch_curve(const Config &cfg, const PassBuffers &pb, cudaStream_t st, bool fused_prepass);   // fused_prepass: launch_prepass did not run
// ISO mode level 2: psychoacoustic model (needs the PCM and the pre-pass's M/S decision) and the outer loop (after launch_curve)
int launch_blocktype(const Config &cfg, const PassBuffers &pb, cudaStream_t st);   // ISO mode level 3: block types, after launch_prepass
int launch_psy(const Config &cfg, const PassBuffers &pb, cudaStream_t st);
int launch_outer(const Config &cfg, const PassBuffers &pb, cudaStream_t st);
// ISO mode: the main-data FIFO of a pass must start out zeroed (stuffing bytes are never written)
int launch_clear_md(const Config &cfg, const PassBuffers &pb, cudaStream_t st);
int launch_scan(const Config &cfg, const PassBuffers &pb, cudaStream_t st);
int launch_pack(const Config &cfg, const PassBuffers &pb, cudaStream_t st);
int launch_frames(const Config &cfg, const PassBuffers &pb, cudaStream_t st);
// the pass boundary: what the next pass's head needs (after launch_curve) / what its scan needs (after launch_frames)
int launch_carry_head(const Config &cfg, const PassBuffers &pb, cudaStream_t st);
int launch_carry_tail(const Config &cfg, const PassBuffers &pb, cudaStream_t st);
int launch_thresholds(const Config &cfg, const PassBuffers &pb, cudaStream_t st);
// out_len[s] = state[s].out_pos; offsets = exclusive prefix sum; compact = gather of out rows
int launch_compact(const Config &cfg, const PassBuffers &pb, uint64_t *offsets /* [S+1] */, uint8_t *compact, int gather,
                   cudaStream_t st);
int launch_synth(float *d_pcm, size_t n_per_channel, int channels, int sample_rate, float f_left, float f_right,
                 float amp, float noise, uint64_t seed, cudaStream_t st);
int launch_table_dump(int which, void *d_out, cudaStream_t st);
int launch_selftest(unsigned long long *d_mismatch /* [3] */, cudaStream_t st);
int launch_widen_i16(const int16_t *in, float *out, size_t stride, const StreamPlan *d_plan, int n_streams, cudaStream_t st);

}  // namespace mp3b
