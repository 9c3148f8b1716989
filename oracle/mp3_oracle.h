/* mp3_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, single-threaded per session) of the encode path of mierau/swift-mp3
 * (Sources/SwiftMP3/MP3Encoder.swift, "SRC" below).  It is the parity checker for the CUDA engine and the
 * timed CPU baseline; it is never linked into, imported by or called from the product library.
 *
 * PARITY UNPINNED: the reference is Swift + Apple Accelerate and cannot be built or run on this box, and
 * its own tests hold no numeric golden vectors.  The oracle is therefore pinned only on
 * the reference's structural known-answer tests (frame sizes, padding ratio, one-frame delay, counters,
 * main_data_begin layout, determinism; tests/test_oracle_kat.py).  Where Accelerate's behaviour is
 * unknowable the choice is marked ORACLE-DEFINED in mp3_oracle.c.
 */
#ifndef MP3_ORACLE_H
#define MP3_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* MP3EncoderOptions (SRC:57-116). mode: 0 mono, 1 stereo, 2 jointStereo. */
typedef struct {
  int32_t sample_rate, bitrate_kbps, vbr, mode, quality, crc_protected, original, copyright;
} orc_options;

/* ID3Tag (SRC:8-54); NULL string = nil, negative number = nil. */
typedef struct {
  const char *title, *artist, *album, *genre, *comment;
  int32_t track, track_total, year;
  const uint8_t *album_art; size_t album_art_len; const char *album_art_mime;
} orc_id3;

/* Per granule-channel trace (one record per call of quantizeToFitBudget, in encode order). */
typedef struct {
  float   spectrum[576];   /* MDCT.apply output, SRC:1512-1565 */
  float   subband[576];    /* analyzeSubbands output [sb*18+t] before the MDCT sign flip, SRC:917-944 */
  float   thresholds[576]; /* PsychoacousticModel.maskingThresholds, SRC:1983-2013 (dead output) */
  int32_t ix[576];         /* quantized values returned by quantizeToFitBudget, SRC:734-794 */
  float   energy;          /* FrameAnalysis.energy(granuleSamples), SRC:673 */
  float   sub_energy[3];   /* TransientDetector thirds, SRC:1947-1953 */
  int32_t block_type, mixed, window_switching, subblock_gain[3];
  int32_t g0;              /* computeGlobalGain, SRC:989-1006 */
  int32_t gain_out;        /* global_gain written to side info (SRC:712) */
  int32_t gain_used;       /* gain that produced ix (differs from gain_out on loop exit) */
  int32_t iterations;      /* quantizeWithGain calls made */
  int32_t bits;            /* part2_3_length before 12-bit masking */
  int32_t max_bits;        /* bitsPerGranule, SRC:650 */
  int32_t big_values, region0, region1, preflag;
  int32_t frame, gr, ch;
} orc_gc_trace;

/* Per frame trace (one record per encodeFrame call). */
typedef struct {
  float   frame_energy;    /* SRC:477 */
  int32_t ms;              /* StereoDecision chose mid/side, SRC:2158 */
  int32_t bitrate_kbps, bitrate_index, padding, frame_size, main_data_size;
  int32_t main_data_begin, reservoir_bits, huff_bytes, is_final;
} orc_frame_trace;

typedef struct orc_session orc_session;

orc_session *orc_create(const orc_options *opts);
void orc_destroy(orc_session *s);
/* Deep copy (EncoderSession is a Swift value type: copying it is a full snapshot, SRC:237-258). */
orc_session *orc_clone(const orc_session *s);

/* EncoderSession.encode(samples:) SRC:297-310 / flush() SRC:318-350.
 * The returned pointer is owned by the session and valid until the next call on it. */
const uint8_t *orc_encode(orc_session *s, const float *pcm, size_t n_floats, size_t *out_len);
const uint8_t *orc_flush(orc_session *s, size_t *out_len);
/* generateXingHeader SRC:367-420 (uses current counters / frame sizes). */
const uint8_t *orc_xing_header(orc_session *s, size_t *out_len);
uint32_t orc_frame_count(const orc_session *s);
uint32_t orc_byte_count(const orc_session *s);

/* ID3TagWriter.build SRC:1040-1075.  Caller frees with orc_free. */
uint8_t *orc_id3_build(const orc_id3 *tag, size_t *out_len);
void orc_free(void *p);

/* Tracing (off by default).  Records accumulate until orc_trace_clear. */
void orc_trace_enable(orc_session *s, int on);
size_t orc_trace_gc_count(const orc_session *s);
const orc_gc_trace *orc_trace_gc(const orc_session *s);
size_t orc_trace_frame_count(const orc_session *s);
const orc_frame_trace *orc_trace_frames(const orc_session *s);
void orc_trace_clear(orc_session *s);

/* Table access for cross-checks against the product tables. */
const float *orc_table_window(void);        /* 512 */
const float *orc_table_analysis(void);      /* 32*64 [k*64+n] */
const float *orc_table_mdct_long(void);     /* 18*36 */
const float *orc_table_mdct_short(void);    /* 6*12 */
const float *orc_table_win_long(void);      /* 36 */
const float *orc_table_win_short(void);     /* 12 */
const uint8_t *orc_table_len15(void);       /* 256 */
const uint8_t *orc_table_code15(void);      /* 256 */
float orc_inv_step(int gain);               /* 1/quantizerStep, SRC:798-800 */
float orc_pow34(float a);                   /* ORACLE-DEFINED |x|^0.75 */
int   orc_bitrate_index(int bitrate, int sample_rate); /* SRC:2509-2523 */
int   orc_bitrate_value(int index);                    /* SRC:2526-2530 */

/* Multi-threaded driver for the CPU baseline: encodes n_streams independent streams (encode + flush),
 * one session per stream, streams distributed over n_threads.  Returns total output bytes; out_digest
 * (optional) receives an order-independent checksum. */
size_t orc_encode_streams(const orc_options *opts, const float *const *pcm, const size_t *n_floats,
                          size_t n_streams, int n_threads, uint64_t *out_digest);

/* Parity driver: like orc_encode_streams, but every stream's bytes are compared with expect[i][0 .. expect_len[i]) (the
 * output of the implementation under test); chunk_floats > 0 feeds the stream in chunks of that many floats.
 * first_diff[i] (optional) = -1 when identical, else the offset of the first differing byte.  Returns the number of
 * streams that differ. */
size_t orc_compare_streams(const orc_options *opts, const float *const *pcm, const size_t *n_floats, size_t n_streams,
                           size_t chunk_floats, int n_threads, const uint8_t *const *expect, const size_t *expect_len,
                           int64_t *first_diff);

/* CPU twin of the product's synthetic-PCM generator (mp3b_synth_fill): bit-identical floats from the same arguments. */
void orc_synth_fill(float *pcm, size_t n_per_channel, int channels, int sample_rate, float f_left, float f_right,
                    float amp, float noise, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
