/* mp3b200.h — C ABI of the B200-native MP3 encode engine (libmp3b200.so).
 *
 * Drop-in boundary for the encode path of mierau/swift-mp3.  Citations are into the reference's
 * Sources/SwiftMP3/MP3Encoder.swift ("SRC").  No torch / C++ types cross this boundary: plain pointers and
 * sizes only.  All functions return 0 (MP3B_OK) or a negative mp3b_status; mp3b_last_error() gives the text
 * of the calling thread's most recent failure.
 *
 * Two planes:
 *   session plane — one handle == one reference `EncoderSession` (SRC:237-350); what a Swift facade binds.
 *   batch plane   — N independent sessions that share options and advance together; every call is ONE pass
 *                   of the device pipeline over all streams (this is what the GPU wants; a session is a
 *                   batch of one).
 * Handles are not thread-safe; distinct handles may be used from different threads.  There is no CPU
 * fallback: creation fails with MP3B_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef MP3B200_H
#define MP3B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MP3B_VERSION 100 /* 0.1.0 */

typedef enum {
  MP3B_OK = 0,
  MP3B_ERR_BAD_ARG = -1,          /* null pointer, unsupported option value, bad stream index */
  MP3B_ERR_CUDA = -2,             /* CUDA runtime failure (sticky for the handle) */
  MP3B_ERR_OOM = -3,              /* host or device allocation failed */
  MP3B_ERR_BUFFER_TOO_SMALL = -4, /* *written holds the size needed; the data stays pending in the handle */
  MP3B_ERR_INTERNAL = -5          /* an engine limit was exceeded (e.g. reservoir backlog > 8 KiB) */
} mp3b_status;

/* MP3EncoderOptions, SRC:57-116.  mode: 0 = .mono, 1 = .stereo, 2 = .jointStereo (SRC:59-63).
 * quality is clamped to 0...9 like SRC:110.  sample_rate must be > 0; a sample rate / bitrate pair whose
 * MPEG-1 frame would be smaller than header + side info (the reference traps there) is MP3B_ERR_BAD_ARG. */
typedef struct mp3b_options {
  int32_t sample_rate;   /* default 44100 */
  int32_t bitrate_kbps;  /* default 128; CBR rate or VBR base rate */
  int32_t vbr;           /* default 0 */
  int32_t mode;          /* default 1 (.stereo) */
  int32_t quality;       /* default 5 */
  int32_t crc_protected; /* default 0 */
  int32_t original;      /* default 1 */
  int32_t copyright;     /* default 0 */
} mp3b_options;

/* ID3Tag, SRC:8-54.  NULL string = nil; negative number = nil. */
typedef struct mp3b_id3 {
  const char *title, *artist, *album, *genre, *comment;
  int32_t track, track_total, year;
  const uint8_t *album_art;
  size_t album_art_len;
  const char *album_art_mime; /* NULL = "image/jpeg" (SRC:41) */
} mp3b_id3;

typedef struct mp3b_session mp3b_session;
typedef struct mp3b_batch mp3b_batch;

/* ---- library ---------------------------------------------------------------------------------------- */
int mp3b_version(void);
const char *mp3b_last_error(void);
/* Fills *opts with the reference defaults (MP3EncoderOptions.init, SRC:95-115). */
void mp3b_options_default(mp3b_options *opts);
/* Number of usable CUDA devices (compute capability 10.x). */
int mp3b_device_count(int *count);

/* ---- session plane: EncoderSession ------------------------------------------------------------------ */
/* MP3Encoder.newSession(), SRC:143-145 / EncoderSession.init, SRC:268-282.  device = CUDA ordinal. */
int mp3b_session_create(const mp3b_options *opts, int device, mp3b_session **out);
void mp3b_session_destroy(mp3b_session *s);
/* `var copy = session` — EncoderSession is a struct of values (SRC:237-258), so a copy is a complete, independent snapshot of the
 * encoder (checkpoint / fork).  The clone continues bit-identically to the original when fed the same samples. */
int mp3b_session_clone(const mp3b_session *s, mp3b_session **out);
/* EncoderSession.encode(samples:), SRC:297-310.  pcm = interleaved f32 in [-1, 1], any length (0 allowed).
 * Writes 0...k whole MP3 frames to out; the first full frame of a session yields 0 bytes (one-frame delay,
 * SRC:546-562).  If cap is too small: MP3B_ERR_BUFFER_TOO_SMALL, *written = bytes needed, and the frames
 * stay pending — fetch them with mp3b_session_take_output before the next encode/flush. */
int mp3b_session_encode(mp3b_session *s, const float *pcm, size_t n_floats, uint8_t *out, size_t cap, size_t *written);
/* EncoderSession.flush(), SRC:318-350.  Idempotent after the first call. */
int mp3b_session_flush(mp3b_session *s, uint8_t *out, size_t cap, size_t *written);
int mp3b_session_take_output(mp3b_session *s, uint8_t *out, size_t cap, size_t *written);
/* Upper bound of the bytes one encode(n_floats) / flush() call can return for this session. */
size_t mp3b_session_output_bound(const mp3b_session *s, size_t n_floats);
/* EncoderSession.generateXingHeader(), SRC:367-449 (uses the counters as of now). */
int mp3b_session_xing_header(const mp3b_session *s, uint8_t *out, size_t cap, size_t *written);
/* Size of the Xing placeholder MP3Encoder.encode(_:to:) reserves before the first audio frame, SRC:198-200:
 * 144 * bitrateValue(bitrateIndex(bitrate_kbps)) * 1000 / sample_rate, i.e. from the SNAPPED bitrate.  Host only. */
int mp3b_xing_frame_size(const mp3b_options *opts);
/* EncoderSession.encodedFrameCount / encodedByteCount, SRC:261-264. */
uint32_t mp3b_session_frame_count(const mp3b_session *s);
uint32_t mp3b_session_byte_count(const mp3b_session *s);
/* EncoderSession.generateID3Tag() / ID3TagWriter.build, SRC:355-358, 1040-1075.  Empty tag -> 0 bytes. */
int mp3b_id3_build(const mp3b_id3 *tag, uint8_t *out, size_t cap, size_t *written);

/* ---- batch plane ------------------------------------------------------------------------------------ */
/* n_streams independent EncoderSessions with the same options on one device. */
int mp3b_batch_create(const mp3b_options *opts, int n_streams, int device, mp3b_batch **out);
/* Same with an explicit pass size (frames of every stream processed per device pass; 0 = automatic). */
int mp3b_batch_create_ex(const mp3b_options *opts, int n_streams, int device, int frames_per_pass, mp3b_batch **out);
/* The same batch spread over several devices (sessions share nothing, SRC:237-258, so the batch is
 * partitioned BY STREAM with no collective): streams are cut into n_dev contiguous blocks, block k lives on devices[k] and is
 * driven by a host thread of its own, so uploads, kernels and downloads of all devices overlap.  Every mp3b_batch_* call
 * works on the handle as on a single-device batch, indexed by the global stream number.  A device may be listed more than
 * once.  frames_per_pass: 0 = automatic.  For mp3b_batch_encode_device, d_pcm[i] must live on mp3b_batch_stream_device(b, i). */
int mp3b_batch_create_multi(const mp3b_options *opts, int n_streams, const int *devices, int n_dev, int frames_per_pass, mp3b_batch **out);
int mp3b_batch_device_count(const mp3b_batch *b);
int mp3b_batch_stream_device(const mp3b_batch *b, int stream);
int mp3b_batch_frames_per_pass(const mp3b_batch *b);
void mp3b_batch_destroy(mp3b_batch *b);
int mp3b_batch_stream_count(const mp3b_batch *b);
/* Every stream back to a fresh EncoderSession (SRC:268-282) without reallocating anything. */
/* Snapshot of all n_streams sessions (see mp3b_session_clone). */
int mp3b_batch_clone(const mp3b_batch *b, mp3b_batch **out);
int mp3b_batch_reset(mp3b_batch *b);
/* The same for one stream (a session slot handed to a new caller). */
int mp3b_batch_reset_stream(mp3b_batch *b, int stream);
/* One encode(samples:) per stream: pcm[i] = HOST pointer to n_floats[i] interleaved floats (may be 0 / NULL).
 * flush != 0 additionally performs flush() on every stream afterwards (SRC:318-350).  flush_mask (may be
 * NULL) restricts the flush to streams with a non-zero byte.  The H2D copies, the device pipeline and the
 * D2H copy of the produced frames all happen inside the call; results are read with mp3b_batch_output. */
int mp3b_batch_encode(mp3b_batch *b, const float *const *pcm, const size_t *n_floats, int flush, const uint8_t *flush_mask);
/* Same for callers that stage every stream in one host arena: stream i = base + i * pitch_floats, n_floats[i] <= pitch_floats
 * floats of it are new.  All rows being readable up to the pitch, the upload is a single strided copy even when only some of
 * the streams have data (the session pool's case). */
int mp3b_batch_encode_strided(mp3b_batch *b, const float *base, size_t pitch_floats, const size_t *n_floats, int flush, const uint8_t *flush_mask);
/* Extension for 16-bit sources (not in the reference, whose input is [Float]): pcm[i] = n_samples[i] interleaved int16 values.
 * Exactly mp3b_batch_encode on Float(pcm[i][k]) / 32768 — the widening runs on the device, so half the bytes cross PCIe. */
int mp3b_batch_encode_i16(mp3b_batch *b, const int16_t *const *pcm, const size_t *n_samples, int flush, const uint8_t *flush_mask);
/* Same, but pcm[i] are DEVICE pointers (same device, 4-byte aligned) and the produced frames stay in device
 * memory; only the per-stream byte counts come back.  download != 0 also copies the frames to the host. */
int mp3b_batch_encode_device(mp3b_batch *b, const float *const *d_pcm, const size_t *n_floats, int flush, int download);
/* Result of the last call for stream i: *data points into batch-owned pinned host memory (valid until the next
 * encode on this batch), *len = bytes (whole frames). */
int mp3b_batch_output(const mp3b_batch *b, int stream, const uint8_t **data, size_t *len);
/* Device-side view of the same result (valid after either encode variant). */
int mp3b_batch_output_device(const mp3b_batch *b, int stream, const uint8_t **d_data, size_t *len);
/* Total bytes produced by the last call over all streams. */
size_t mp3b_batch_output_total(const mp3b_batch *b);
int mp3b_batch_xing_header(const mp3b_batch *b, int stream, uint8_t *out, size_t cap, size_t *written);
uint32_t mp3b_batch_frame_count(const mp3b_batch *b, int stream);
uint32_t mp3b_batch_byte_count(const mp3b_batch *b, int stream);

/* ---- ISO mode (extension, default off: with it off every byte equals the reference's) ------------------------------------
 * What north_star lists beyond the reference's live code — its dead stubs HuffmanEncoder.encode / writePair / selectTable
 * (SRC:1740-1806) and Huffman tables 1-13 (SRC:2288-2398) — done the ISO 11172-3 way: the ISO quantizer law on the decoder's
 * scale, global_gain by bit-count search, rzero / count1 / big_values partition, three regions with the cheapest of tables
 * 1-3, 5-13, 15, 16-31 (linbits escapes) each, count1 table A or B, a real main_data_begin back pointer with stuffing, M/S
 * signalled per frame with 1/sqrt(2) scaling, the ISO CRC.  An independent decoder then reconstructs the INPUT signal.  Long
 * blocks only at levels 1 and 2.  Only on fresh sessions (after create / reset).
 * on = 2 adds north_star stages (3) and (4) in full (csrc/iso_psy.cuh; the reference's stubs: ScaleFactorBands.scale SRC:1831-1876,
 * ScaleFactorCompression SRC:2017-2037, its unused masking thresholds SRC:1983-2013): a psychoacoustic model batched over
 * granule-channels — 1024-point FFT line energies, 256-point FFT unpredictability, 1/3-Bark partitions, spreading function,
 * tonality, masking thresholds, perceptual entropy — and the scalefactor outer loop (noise per band against the threshold,
 * amplification, scalefac_compress, part2 bits); the perceptual entropy steers each granule's share of the bit reservoir.
 * on = 3 adds window switching (north_star stage 2: "block type taken from a transient-detection kernel"): ISO start / short / stop
 * blocks with the windows the reference defines and never uses (SRC:1470-1503), short-block lines in scalefactor-band order,
 * two-region side info.  The start window needs one granule of look-ahead: at this level the signal is coded 576 samples late
 * (one more granule of encoder delay) and joint stereo is coded as L / R. */
int mp3b_batch_set_iso_mode(mp3b_batch *b, int on);
int mp3b_batch_iso_mode(const mp3b_batch *b);
int mp3b_session_set_iso_mode(mp3b_session *s, int on);

/* ---- matrixing of the polyphase filterbank (extension, default 0) -----------------------------------------------------------
 * north_star stage (1): the 32 x 64 cosine matrixing (PolyphaseFilterbank.analyze, SRC:1402-1408) "runs on tensor cores only if
 * 3xTF32 split-precision stays inside tolerance, and on FP32 FMA otherwise".  0 = FP32 FMA in the reference's summation order
 * (the default: every byte equals the oracle's); 1 = tcgen05 tensor cores, every FP32 operand split exactly into three TF32
 * terms, six products accumulated in FP32 — the subband samples then differ from the FP32 path in the last bits (tier 1 holds
 * with a wide margin; how many granules change their quantized values is measured by the tests and reported under profiles/). */
int mp3b_batch_set_matrixing(mp3b_batch *b, int mode);
int mp3b_batch_matrixing(const mp3b_batch *b);

/* ---- session pool: many threads, one step ------------------------------------------------------------- */
/* n_sessions EncoderSessions (SRC:237-350) driven from concurrent threads — one blocking call per session at a time,
 * like N reference sessions on N threads — and advanced on the GPU together: the calls that arrive within max_wait_us of
 * each other (or all open sessions, whichever happens first) become one step of the batch plane.  BASELINE config 5. */
typedef struct mp3b_pool mp3b_pool;
int mp3b_pool_create(const mp3b_options *opts, int n_sessions, int device, int max_wait_us, mp3b_pool **out);
void mp3b_pool_destroy(mp3b_pool *p);
/* MP3Encoder.newSession(): claims a free session of the pool; *slot identifies it in the calls below. */
int mp3b_pool_open(mp3b_pool *p, int *slot);
/* EncoderSession.encode(samples:) / flush() of session `slot`.  Blocks until the step that contains the request has run.
 * flush() closes the session (it no longer delays the steps of the others). */
int mp3b_pool_encode(mp3b_pool *p, int slot, const float *pcm, size_t n_floats, uint8_t *out, size_t cap, size_t *written);
int mp3b_pool_flush(mp3b_pool *p, int slot, uint8_t *out, size_t cap, size_t *written);
/* When encode / flush returned MP3B_ERR_BUFFER_TOO_SMALL (*written = the size needed) the bytes are kept: fetch them here
 * before the session's next call (which is refused until then).  Same contract as mp3b_session_take_output. */
int mp3b_pool_take_output(mp3b_pool *p, int slot, uint8_t *out, size_t cap, size_t *written);
/* Steps run so far and requests served by them (requests / steps = achieved coalescing). */
int mp3b_pool_stats(mp3b_pool *p, uint64_t *steps, uint64_t *requests);
const char *mp3b_pool_last_error(void);

/* Pinned host memory helpers (cudaHostAlloc / cudaFreeHost) so callers can stage PCM for full-speed H2D. */
int mp3b_host_alloc(size_t bytes, void **out);
void mp3b_host_free(void *p);
/* Device memory helpers for the device plane. */
int mp3b_device_alloc(int device, size_t bytes, void **out);
void mp3b_device_free(int device, void *p);
int mp3b_device_copy(int device, void *dst, const void *src, size_t bytes, int kind /* 0 H2D, 1 D2H, 2 D2D */);
int mp3b_device_sync(int device);

/* ---- measurement / test plane (not part of the reference API) ---------------------------------------- */
/* Per-stage device time of the last batch call, in milliseconds, summed over the call's passes.
 * Stage order: see MP3B_STAGE_*.  n = entries available in ms[]. */
enum {
  MP3B_STAGE_H2D = 0, MP3B_STAGE_PREPASS, MP3B_STAGE_SPECTRUM /* k_filterbank */, MP3B_STAGE_CURVE /* k_granule: MDCT + curve */, MP3B_STAGE_SCAN, MP3B_STAGE_PACK,
  MP3B_STAGE_FRAMES, MP3B_STAGE_D2H, MP3B_STAGE_TOTAL, MP3B_STAGE_COUNT
};
int mp3b_batch_stage_ms(const mp3b_batch *b, float *ms, int n);
/* Number of device passes of the last batch call (each stage kernel is launched once per pass). */
int mp3b_batch_pass_count(const mp3b_batch *b);
/* The CUDA stream (cudaStream_t) all work of this batch is issued on, so that callers can record their own events
 * (multi-device batch: the stream of its first device; stage times are then the maximum over the devices). */
void *mp3b_batch_stream(const mp3b_batch *b);
/* Number of kernel launches issued by the last batch call. */
int mp3b_batch_launch_count(const mp3b_batch *b);
/* Enable per-granule-channel traces (costs device memory and bandwidth; off by default).  bit0: keep MDCT
 * spectra, bit1: keep quantized ix, bit2: also run the (dead, SRC:737) masking-threshold kernel. */
int mp3b_batch_set_trace(mp3b_batch *b, int flags);
/* Granule-channel record of the last call, in encode order per stream (frame-major, then gr, then ch). */
typedef struct mp3b_gc_record {
  int32_t part23_length, big_values, global_gain, gain_used, block_type, subblock_gain[3];
  int32_t region0, region1, preflag, g0, max_bits, iterations;
  float energy;
  int32_t table_select[3], count1table_select;   /* 15, 15, 15 and 0 unless ISO mode is on */
  int32_t scalefac_compress, part2_length;       /* 0 and 0 unless ISO mode level 2 is on (part23_length includes part2_length) */
} mp3b_gc_record;
typedef struct mp3b_frame_record {
  int32_t bitrate_index, padding, frame_size, main_data_size, main_data_begin, reservoir_bits, huff_bytes, ms, is_final;
  float frame_energy;
} mp3b_frame_record;
/* Number of frames the last call encoded for `stream` (its granule-channel count is frames * 2 * channels). */
int mp3b_batch_trace_frames(const mp3b_batch *b, int stream);
int mp3b_batch_trace_frame_records(const mp3b_batch *b, int stream, mp3b_frame_record *out, int cap);
int mp3b_batch_trace_gc_records(const mp3b_batch *b, int stream, mp3b_gc_record *out, int cap);
/* kind: 0 spectrum f32[576], 1 ix i32[576], 2 thresholds f32[576]; out holds cap_gc * 576 elements.
 * ISO mode level 2: kind 3 = psychoacoustic record f32[24] (threshold / energy of the 22 long scalefactor bands, perceptual
 * entropy, mean tonality), kind 4 = scalefactor record i32[24] (21 scalefactors, scalefac_compress, part2 bits, bands left
 * over their threshold | outer iterations << 4); out holds cap_gc * 24 elements. */
int mp3b_batch_trace_gc_array(const mp3b_batch *b, int stream, int kind, void *out, int cap_gc);
/* Product tables for cross-checks against the oracle / the reference literals.
 * which: 0 window[512] f32, 1 analysis[32*64] f32, 2 mdct_long[18*36] f32, 3 mdct_short[6*12] f32,
 * 4 win_long[36] f32, 5 win_short[12] f32, 6 inv_step[256] f32, 7 len15[256] u8, 8 code15[256] u8,
 * 9 gain_threshold[256] f64, 10 alias_cs[8] f32, 11 alias_ca[8] f32, 12 sfb_cum[3*21] i32,
 * 13 len31s[31*32] u8, 14 tab31[31*32] u16: table 15 indexed by the kernels' unclamped quantizer value (u = min(floor(2t), 30),
 * q = (u+1)>>1, row stride 32): pair length incl. sign bits, and code | length << 8.
 * Returns the element count, or a negative status. */
int mp3b_table(int which, void *out, size_t cap_bytes);
/* Deterministic synthetic PCM written straight into device memory (bench / tests): stream `seed` of the
 * BASELINE C1/C4 recipe — L = a*sin(2*pi*fL*t) + n*N(0,1), R likewise with fR and phase 0.3, clipped to
 * [-1, 1]; interleaved when channels == 2. */
int mp3b_synth_fill(int device, float *d_pcm, size_t n_samples_per_channel, int channels, int sample_rate,
                    float f_left, float f_right, float amp, float noise, uint64_t seed);

/* Exhaustive on-device check of the two exact-arithmetic shortcuts of the kernels (tests only): mismatches[0] = floats in
 * [1e-10, FLT_MAX] where the fast |x|^0.75 differs from its IEEE-double definition, [1] / [2] = finite floats where the
 * FMA-based division by 9 / 3 differs from the IEEE division.  All three must be 0. */
int mp3b_selftest(int device, uint64_t mismatches[3]);

#ifdef __cplusplus
}
#endif
#endif /* MP3B200_H */
