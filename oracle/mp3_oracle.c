/* mp3_oracle.c — TEST INFRASTRUCTURE ONLY (see mp3_oracle.h).
 *
 * Plain-C restatement of the encode path of mierau/swift-mp3, Sources/SwiftMP3/MP3Encoder.swift ("SRC").
 * Every function cites the SRC lines it follows.  PARITY UNPINNED against the real reference (Swift +
 * Apple Accelerate cannot be built here, and the reference's tests hold no numeric vectors): the oracle is
 * pinned on the reference's structural known-answer tests only (tests/test_oracle_kat.py).
 *
 * ORACLE-DEFINED choices (where Accelerate / Darwin libm behaviour is unknowable); each is an exactly
 * reproducible sequence of IEEE-754 operations so that the CUDA engine can be compared bit for bit:
 *   [OD1] vDSP_dotpr / vDSP_sve (short, fixed-length: 64/36/12-term dot products, 8-term strided sums, the
 *         <=10-term VBR history sum): accumulator starts at +0 and walks the index upwards; dot products
 *         use one fused multiply-add per element (fmaf), plain sums one rounded add per element.
 *   [OD1b] vDSP_svesq (sums of squares over 144...2304 elements: all energies): Accelerate is a SIMD
 *         library, so a strictly serial order is no more faithful than a laned one.  The oracle defines
 *         32 interleaved partial sums — element i of the segment (counted from the segment start) goes to
 *         partial i mod 32, ascending i, one fmaf each — combined by a fixed butterfly tree
 *         p[j] += p[j ^ 16], then ^8, ^4, ^2, ^1 (result = p[0]).
 *   [OD2] vDSP_vmul / vsmul / vadd / vsub: one correctly rounded FP32 operation per element.
 *   [OD3] vvpowf(x, 0.75) and powf(peak, 0.75): pow34(a) = (float)(sqrt(d) * sqrt(sqrt(d))), d=(double)a,
 *         all in IEEE double.  This equals the correctly rounded powf(a, 0.75f) except when a^0.75 lies
 *         within ~2 double-ulps of an FP32 rounding boundary (never observed in 5e7 random trials, see
 *         tests/test_oracle_kat.py::test_pow34_matches_libm).
 *   [OD4] Float.rounded() = roundf (ties away from zero); Int(x) = truncation toward zero.
 *   [OD5] SIMD8 alias butterflies: two rounded multiplies and one rounded add/sub (no contraction).
 * Build with -ffp-contract=off and without -ffast-math (oracle/Makefile).
 */
#define _GNU_SOURCE
#include "mp3_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "iso_tables.inc"

/* ------------------------------------------------------------------------------------------------ */
/* growable byte buffer (Swift `Data`)                                                              */
typedef struct { uint8_t *p; size_t n, cap; } bytes_t;
static void by_reserve(bytes_t *b, size_t extra) {
  if (b->n + extra <= b->cap) return;
  size_t c = b->cap ? b->cap * 2 : 1024;
  while (c < b->n + extra) c *= 2;
  b->p = (uint8_t *)realloc(b->p, c); b->cap = c;
}
static void by_push(bytes_t *b, uint8_t v) { by_reserve(b, 1); b->p[b->n++] = v; }
static void by_append(bytes_t *b, const uint8_t *src, size_t n) { if (!n) return; by_reserve(b, n); memcpy(b->p + b->n, src, n); b->n += n; }
static void by_zeros(bytes_t *b, size_t n) { if (!n) return; by_reserve(b, n); memset(b->p + b->n, 0, n); b->n += n; }
static void by_drop_front(bytes_t *b, size_t n) { memmove(b->p, b->p + n, b->n - n); b->n -= n; }
static void by_free(bytes_t *b) { free(b->p); b->p = NULL; b->n = b->cap = 0; }
static bytes_t by_clone(const bytes_t *b) { bytes_t c = {0}; by_append(&c, b->p, b->n); return c; }

/* BitstreamWriter, SRC:2219-2275 */
typedef struct { bytes_t data; uint32_t buffer; int bits_in_buffer; } bitw_t;
static void bw_write(bitw_t *w, int bits, int count) {           /* SRC:2230-2252 */
  if (!(count > 0 && count <= 24)) {                              /* SRC:2231-2236 (bit-at-a-time fallback) */
    for (int i = count - 1; i >= 0; --i) {
      w->buffer = (w->buffer << 1) | (uint32_t)((bits >> i) & 1);
      if (++w->bits_in_buffer == 8) { by_push(&w->data, (uint8_t)(w->buffer & 0xFF)); w->buffer = 0; w->bits_in_buffer = 0; }
    }
    return;
  }
  w->buffer = (w->buffer << count) | (uint32_t)(bits & ((1 << count) - 1));
  w->bits_in_buffer += count;
  while (w->bits_in_buffer >= 8) {
    w->bits_in_buffer -= 8;
    by_push(&w->data, (uint8_t)((w->buffer >> w->bits_in_buffer) & 0xFF));
  }
  if (w->bits_in_buffer > 0) w->buffer &= (1u << w->bits_in_buffer) - 1; else w->buffer = 0;
}
static void bw_pad(bitw_t *w) {                                   /* SRC:2266-2274 */
  if (w->bits_in_buffer > 0) {
    w->buffer <<= (8 - w->bits_in_buffer);
    by_push(&w->data, (uint8_t)(w->buffer & 0xFF));
    w->buffer = 0; w->bits_in_buffer = 0;
  }
}
static int bw_bitcount(const bitw_t *w) { return (int)w->data.n * 8 + w->bits_in_buffer; }  /* SRC:2225-2227 */

/* ------------------------------------------------------------------------------------------------ */
/* tables                                                                                           */
static float T_window[512], T_analysis[32 * 64], T_analysis_t[64 * 32];
static float T_mdct_long[18 * 36], T_mdct_long_t[36 * 18], T_mdct_short[6 * 12], T_win_long[36], T_win_short[12];
static float T_inv_step[256];
#ifdef ORC_VARIANTS
#include "variants.inc"   /* alternative summation orders / pow / split-TF32 matrixing for tools/order_sensitivity.py; never in the default build */
#endif
static uint16_t T_crc[256];
static pthread_once_t tables_once = PTHREAD_ONCE_INIT;

static void build_tables(void) {
  /* SRC:1209-1354: the literals as printed (9 decimals), parsed to FP32 like the Swift compiler does. */
  for (int i = 0; i < 512; ++i) {
    char buf[32]; snprintf(buf, sizeof buf, "%.9f", (double)ISO_WINDOW_K[i] / 2097152.0);
    T_window[i] = strtof(buf, NULL);
  }
  for (int k = 0; k < 32; ++k)                                     /* SRC:1197-1206 */
    for (int n = 0; n < 64; ++n) {
      double angle = M_PI / 64.0 * (double)(2 * k + 1) * ((double)n - 16.0);
      T_analysis[k * 64 + n] = T_analysis_t[n * 32 + k] = (float)cos(angle);
    }
  for (int m = 0; m < 18; ++m)                                     /* SRC:1422-1433 */
    for (int k = 0; k < 36; ++k) {
      double angle = M_PI / (double)(2 * 36) * (double)(2 * k + 1 + 36 / 2) * (double)(2 * m + 1);
      T_mdct_long[m * 36 + k] = T_mdct_long_t[k * 18 + m] = (float)cos(angle);
    }
  for (int m = 0; m < 6; ++m)                                      /* SRC:1436-1447 */
    for (int k = 0; k < 12; ++k) {
      double angle = M_PI / (double)(2 * 12) * (double)(2 * k + 1 + 12 / 2) * (double)(2 * m + 1);
      T_mdct_short[m * 12 + k] = (float)cos(angle);
    }
  for (int i = 0; i < 36; ++i) T_win_long[i] = (float)sin(M_PI / 36.0 * ((double)i + 0.5));   /* SRC:1450-1457 */
  for (int i = 0; i < 12; ++i) T_win_short[i] = (float)sin(M_PI / 12.0 * ((double)i + 0.5));  /* SRC:1460-1467 */
  for (int g = 0; g < 256; ++g) {                                  /* SRC:798-800 */
    double step_power = (double)(g - 210) / 4.0;
    float step = (float)fmax(pow(2.0, step_power), 0.0001);
    T_inv_step[g] = 1.0f / step;
  }
  for (int i = 0; i < 256; ++i) {                                  /* SRC:2191-2205 */
    uint16_t crc = (uint16_t)(i << 8);
    for (int b = 0; b < 8; ++b) crc = (crc & 0x8000) ? (uint16_t)((crc << 1) ^ 0x8005) : (uint16_t)(crc << 1);
    T_crc[i] = crc;
  }
}
static void tables(void) { pthread_once(&tables_once, build_tables); }

const float *orc_table_window(void) { tables(); return T_window; }
const float *orc_table_analysis(void) { tables(); return T_analysis; }
const float *orc_table_mdct_long(void) { tables(); return T_mdct_long; }
const float *orc_table_mdct_short(void) { tables(); return T_mdct_short; }
const float *orc_table_win_long(void) { tables(); return T_win_long; }
const float *orc_table_win_short(void) { tables(); return T_win_short; }
const uint8_t *orc_table_len15(void) { return ISO_HUFF15_LEN; }
const uint8_t *orc_table_code15(void) { return ISO_HUFF15_CODE; }
float orc_inv_step(int gain) { tables(); return T_inv_step[gain < 0 ? 0 : gain > 255 ? 255 : gain]; }

/* [OD3] */
float orc_pow34(float a) {
#ifdef ORC_VARIANTS
  if (g_variant == 4 || g_variant == 5) return pow34_var(a);
#endif
  double d = (double)a; double r = sqrt(d); return (float)(r * sqrt(r));
}

/* MP3Tables.bitrateIndex SRC:2509-2523 */
int orc_bitrate_index(int bitrate, int sample_rate) {
  static const int t1[16] = {0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 0};
  static const int t2[16] = {0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160, 0};
  const int *t = sample_rate >= 32000 ? t1 : t2;
  for (int i = 0; i < 16; ++i) if (t[i] == bitrate) return i;
  int best = 0;                                                    /* min(by:) keeps the first minimum */
  for (int i = 1; i < 16; ++i) if (abs(t[i] - bitrate) < abs(t[best] - bitrate)) best = i;
  return best;
}
/* MP3Tables.bitrateValue SRC:2526-2530 */
int orc_bitrate_value(int index) {
  static const int t1[16] = {0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 0};
  return (index >= 0 && index < 16) ? t1[index] : 128;
}
static int sample_rate_index(int sr) { return sr == 44100 ? 0 : sr == 48000 ? 1 : sr == 32000 ? 2 : 0; }  /* SRC:2533-2544 */
static const uint8_t *band_table(int sr) {                        /* SRC:1879-1888 */
  return sr == 48000 ? ISO_SFB_LONG_48000 : sr == 32000 ? ISO_SFB_LONG_32000 : ISO_SFB_LONG_44100;
}

/* FrameAnalysis.energy SRC:1902-1907  [OD1b] */
static float sumsq(const float *x, int n) {                       /* vDSP_svesq [OD1b] */
#ifdef ORC_VARIANTS
  if (g_variant == 9 || g_variant == 10) return sumsq_var(x, n);
#endif
  float p[32], q[32];
  for (int j = 0; j < 32; ++j) p[j] = 0.0f;
  for (int i = 0; i < n; ++i) p[i & 31] = fmaf(x[i], x[i], p[i & 31]);
  for (int m = 16; m >= 1; m >>= 1) {
    for (int j = 0; j < 32; ++j) q[j] = p[j] + p[j ^ m];
    memcpy(p, q, sizeof p);
  }
  return p[0];
}
static float energy(const float *x, int n) {
  if (n <= 0) return 0.0f;
  return sumsq(x, n) / (float)n;
}

/* ------------------------------------------------------------------------------------------------ */
/* session                                                                                          */
typedef struct { bytes_t header_and_side; int slot_size; int present; } buffered_frame;

struct orc_session {
  orc_options opt;
  int channels;
  float *pcm; size_t pcm_n, pcm_cap;                     /* pcmBuffer SRC:243 */
  float vbr_energy[10]; int vbr_n;                       /* VBRState.energyHistory SRC:1141 (gainHistory is never read) */
  bytes_t res_stream; int res_available;                 /* BitReservoir SRC:2091-2096 */
  buffered_frame buffered;                               /* SRC:246 */
  int padding_remainder;                                 /* SRC:247 */
  float fb[2][512];                                      /* filterbankBuffers SRC:249 */
  float overlap[2][32][18];                              /* mdctOverlap SRC:250 */
  uint32_t frame_count, total_bytes;                     /* SRC:256-257 */
  int *frame_sizes; size_t fs_n, fs_cap;                 /* SRC:258 */
  bytes_t out;                                           /* return buffer of encode/flush/xing */
  int trace_on;
  orc_gc_trace *tg; size_t tg_n, tg_cap;
  orc_frame_trace *tf; size_t tf_n, tf_cap;
};

orc_session *orc_create(const orc_options *opts) {                /* EncoderSession.init SRC:268-282 */
  tables();
  orc_session *s = (orc_session *)calloc(1, sizeof *s);
  s->opt = *opts;
  if (s->opt.quality < 0) s->opt.quality = 0;                      /* SRC:110 */
  if (s->opt.quality > 9) s->opt.quality = 9;
  s->channels = s->opt.mode == 0 ? 1 : 2;
  return s;
}
void orc_destroy(orc_session *s) {
  if (!s) return;
  free(s->pcm); by_free(&s->res_stream); by_free(&s->buffered.header_and_side); free(s->frame_sizes);
  by_free(&s->out); free(s->tg); free(s->tf); free(s);
}
orc_session *orc_clone(const orc_session *s) {
  orc_session *c = (orc_session *)malloc(sizeof *c);
  memcpy(c, s, sizeof *c);
  c->pcm = NULL; c->pcm_cap = 0;
  if (s->pcm_n) { c->pcm = (float *)malloc(s->pcm_n * sizeof(float)); memcpy(c->pcm, s->pcm, s->pcm_n * sizeof(float)); c->pcm_cap = s->pcm_n; }
  c->res_stream = by_clone(&s->res_stream);
  c->buffered.header_and_side = by_clone(&s->buffered.header_and_side);
  c->frame_sizes = NULL; c->fs_cap = 0;
  if (s->fs_n) { c->frame_sizes = (int *)malloc(s->fs_n * sizeof(int)); memcpy(c->frame_sizes, s->frame_sizes, s->fs_n * sizeof(int)); c->fs_cap = s->fs_n; }
  memset(&c->out, 0, sizeof c->out);
  c->tg = NULL; c->tg_n = c->tg_cap = 0; c->tf = NULL; c->tf_n = c->tf_cap = 0;
  return c;
}
uint32_t orc_frame_count(const orc_session *s) { return s->frame_count; }
uint32_t orc_byte_count(const orc_session *s) { return s->total_bytes; }

void orc_trace_enable(orc_session *s, int on) { s->trace_on = on; }
size_t orc_trace_gc_count(const orc_session *s) { return s->tg_n; }
const orc_gc_trace *orc_trace_gc(const orc_session *s) { return s->tg; }
size_t orc_trace_frame_count(const orc_session *s) { return s->tf_n; }
const orc_frame_trace *orc_trace_frames(const orc_session *s) { return s->tf; }
void orc_trace_clear(orc_session *s) { s->tg_n = 0; s->tf_n = 0; }
static orc_gc_trace *trace_gc_new(orc_session *s) {
  if (s->tg_n == s->tg_cap) { s->tg_cap = s->tg_cap ? s->tg_cap * 2 : 64; s->tg = (orc_gc_trace *)realloc(s->tg, s->tg_cap * sizeof *s->tg); }
  orc_gc_trace *t = &s->tg[s->tg_n++]; memset(t, 0, sizeof *t); return t;
}
static orc_frame_trace *trace_frame_new(orc_session *s) {
  if (s->tf_n == s->tf_cap) { s->tf_cap = s->tf_cap ? s->tf_cap * 2 : 64; s->tf = (orc_frame_trace *)realloc(s->tf, s->tf_cap * sizeof *s->tf); }
  orc_frame_trace *t = &s->tf[s->tf_n++]; memset(t, 0, sizeof *t); return t;
}

/* PolyphaseFilterbank.analyze SRC:1367-1411 */
static void filterbank_step(const float *new32, float *buffer, float *out32) {
#ifdef ORC_VARIANTS
  if (g_variant) { filterbank_step_var(new32, buffer, out32); return; }
#endif
  memmove(buffer, buffer + 32, 480 * sizeof(float));               /* SRC:1373 */
  memcpy(buffer + 480, new32, 32 * sizeof(float));                 /* SRC:1375-1381 */
  float z[512];
  for (int i = 0; i < 512; ++i) z[i] = buffer[511 - i] * T_window[i];   /* SRC:1386-1389 [OD2] */
  float y[64];
  for (int j = 0; j < 64; ++j) {                                   /* SRC:1392-1399 [OD1] */
    float sum = 0.0f;
    for (int i = 0; i < 8; ++i) sum = sum + z[j + 64 * i];
    y[j] = sum;
  }
  float acc[32];                                                   /* SRC:1402-1408 [OD1]: per k, n ascending */
  for (int k = 0; k < 32; ++k) acc[k] = 0.0f;
  for (int n = 0; n < 64; ++n) {
    const float *col = T_analysis_t + n * 32; float yn = y[n];
    for (int k = 0; k < 32; ++k) acc[k] = fmaf(yn, col[k], acc[k]);
  }
  memcpy(out32, acc, sizeof acc);
}

/* MDCT.mdctLong SRC:1619-1636 */
static void mdct_long(const float *comb, float *out18) {
#ifdef ORC_VARIANTS
  if (g_variant) { mdct_long_var(comb, out18); return; }
#endif
  float w[36];
  for (int k = 0; k < 36; ++k) w[k] = comb[k] * T_win_long[k];     /* SRC:1625 [OD2] */
  float acc[18];
  for (int m = 0; m < 18; ++m) acc[m] = 0.0f;
  for (int k = 0; k < 36; ++k) {                                   /* SRC:1629-1633 [OD1]: per m, k ascending */
    const float *col = T_mdct_long_t + k * 18; float wk = w[k];
    for (int m = 0; m < 18; ++m) acc[m] = fmaf(wk, col[m], acc[m]);
  }
  for (int m = 0; m < 18; ++m) out18[m] = acc[m] / 9.0f;           /* SRC:1632 */
}
/* MDCT.mdctShort SRC:1639-1662 */
static void mdct_short(const float *comb, float *out18) {
#ifdef ORC_VARIANTS
  if (g_variant) { mdct_short_var(comb, out18); return; }
#endif
  for (int w = 0; w < 3; ++w) {
    int offset = w * 6 + 6;
    float seg[12];
    for (int i = 0; i < 12; ++i) seg[i] = comb[offset + i] * T_win_short[i];   /* SRC:1649-1652 */
    for (int m = 0; m < 6; ++m) {
      float r = 0.0f;
      for (int k = 0; k < 12; ++k) r = fmaf(seg[k], T_mdct_short[m * 12 + k], r);  /* SRC:1656 [OD1] */
      out18[w + m * 3] = r / 3.0f;                                 /* SRC:1657 */
    }
  }
}
/* MDCT.apply SRC:1512-1565 + applyAliasingReduction SRC:1581-1616 */
static void mdct_apply(float sub[32][18], float overlap[32][18], int block_type, float *out576) {
  for (int sb = 0; sb < 32; ++sb) {
    float cur[18], comb[36];
    memcpy(cur, sub[sb], sizeof cur);
    if (sb & 1) for (int k = 1; k < 18; k += 2) cur[k] *= -1.0f;   /* SRC:1520-1524 */
    memcpy(comb, overlap[sb], 18 * sizeof(float));                 /* SRC:1526-1532 */
    memcpy(comb + 18, cur, 18 * sizeof(float));
    memcpy(overlap[sb], cur, 18 * sizeof(float));                  /* SRC:1534-1539 */
    int use_long = block_type == 0 || (block_type == 1 && sb < 2); /* SRC:1542-1553 (mixed raw value 1) */
    if (use_long) mdct_long(comb, out576 + sb * 18); else mdct_short(comb, out576 + sb * 18);
  }
  if (block_type == 0) {                                           /* SRC:1560-1562 */
    for (int sb = 0; sb < 31; ++sb)                                /* SRC:1582-1615 [OD5] */
      for (int i = 0; i < 8; ++i) {
        int iu = sb * 18 + 17 - i, il = (sb + 1) * 18 + i;
        float upper = out576[iu], lower = out576[il];
        float a = lower * ISO_ALIAS_CA[i], b = upper * ISO_ALIAS_CS[i];
        float c = lower * ISO_ALIAS_CS[i], d = upper * ISO_ALIAS_CA[i];
        out576[iu] = a + b;
        out576[il] = c - d;
      }
  }
}

/* TransientDetector.analyze SRC:1944-1968 (samples.count == 576 always on the live path) */
static void transient(const float *x, float *e3, int *block_type, int *sbg3) {
  for (int i = 0; i < 3; ++i) e3[i] = energy(x + 192 * i, 192);
  float mx = e3[0], mn = e3[0];
  for (int i = 1; i < 3; ++i) { if (e3[i] > mx) mx = e3[i]; if (e3[i] < mn) mn = e3[i]; }
  float ratio = mx / fmaxf(mn, 0.0001f);
  if (ratio > 6.0f) *block_type = (e3[0] == mx) ? 1 /* mixed */ : 2 /* short */;
  else *block_type = 0;
  for (int i = 0; i < 3; ++i) {
    float normalized = fminf(fmaxf(e3[i] / fmaxf(mx, 0.0001f), 0.0f), 1.0f);
    sbg3[i] = (int)((1.0f - normalized) * 7.0f);
  }
}

/* PsychoacousticModel.maskingThresholds SRC:1983-2013 (output is never read by the quantizer, SRC:737) */
static void masking_thresholds(const float *spec, int sample_rate, int quality, float *thr) {
  const uint8_t *bands = band_table(sample_rate);
  double qd = (double)(10 - quality) / 10.0;
  float quality_scale = (float)(qd > 0.1 ? qd : 0.1);
  for (int i = 0; i < 576; ++i) thr[i] = 0.0001f;
  int cursor = 0;
  for (int b = 0; b < 21; ++b) {
    int start = cursor, end = cursor + bands[b]; if (end > 576) end = 576;
    int size = end - start;
    if (size <= 0) { cursor = end; continue; }
    float e = sumsq(spec + start, size);
    float average = e / (float)size;
    float t = fmaxf(average * quality_scale, 0.0001f);
    for (int i = start; i < end; ++i) thr[i] = t;
    cursor = end;
    if (cursor >= 576) break;
  }
}

/* computeGlobalGain SRC:989-1006 */
static int compute_global_gain(const float *spec) {
  float peak = 0.0f;
  for (int i = 0; i < 576; ++i) { float a = fabsf(spec[i]); if (a > peak) peak = a; }
  if (!(peak > 0.0f)) return 210;
  float peak_pow = orc_pow34(peak);                                /* SRC:997 [OD3] */
  float ratio = peak_pow / 15.0f;
  if (ratio <= 0.0f) return 210;
  int gain = 210 + (int)(4.0 * log2((double)ratio));               /* SRC:1004 [OD4] */
  return gain < 0 ? 0 : gain > 255 ? 255 : gain;
}

/* quantizeWithGain SRC:797-825.  pow34 does not depend on the gain (SRC:804-813), so it is passed in. */
static void quantize_with_gain(const float *spec, const float *mag, int gain, int32_t *ix) {
  float inv = T_inv_step[gain];
  for (int i = 0; i < 576; ++i) {
    float scaled = mag[i] * inv;                                   /* SRC:816 [OD2] */
    float r = roundf(scaled);                                      /* SRC:820 [OD4] */
    int q = r >= 15.0f ? 15 : (int)r;
    ix[i] = spec[i] < 0.0f ? -q : q;                               /* SRC:821 */
  }
}
static int last_nonzero(const int32_t *ix) {                      /* SRC:750-756 */
  for (int i = 575; i >= 0; --i) if (ix[i] != 0) return i + 1;
  return 0;
}
/* countHuffmanBits SRC:828-853 (count is always even here) */
static int count_bits15(const int32_t *v, int count) {
  int bits = 0;
  for (int i = 0; i + 1 < count; i += 2) {
    int ax = abs(v[i]); if (ax > 15) ax = 15;
    int ay = abs(v[i + 1]); if (ay > 15) ay = 15;
    bits += ISO_HUFF15_LEN[ax * 16 + ay];
    if (ax) ++bits;
    if (ay) ++bits;
  }
  return bits;
}
/* HuffmanEncoder.encodeWithTable15 SRC:1705-1737 */
static int encode_table15(const int32_t *v, int count, bitw_t *w) {
  int start = bw_bitcount(w);
  for (int i = 0; i + 1 < count; i += 2) {
    int x = v[i], y = v[i + 1];
    int ax = abs(x); if (ax > 15) ax = 15;
    int ay = abs(y); if (ay > 15) ay = 15;
    bw_write(w, ISO_HUFF15_CODE[ax * 16 + ay], ISO_HUFF15_LEN[ax * 16 + ay]);
    if (ax) bw_write(w, x < 0 ? 1 : 0, 1);
    if (ay) bw_write(w, y < 0 ? 1 : 0, 1);
  }
  return bw_bitcount(w) - start;
}

/* quantizeToFitBudget SRC:734-794 */
static void quantize_to_fit(const float *spec, int initial_gain, int max_bits, bitw_t *w,
                            int *gain_out, int32_t *ix, int *bits_out, int *gain_used, int *iterations) {
  float mag[576];
  for (int i = 0; i < 576; ++i) mag[i] = orc_pow34(fmaxf(fabsf(spec[i]), 1e-10f));   /* SRC:805-813 [OD3] */
  int gain = initial_gain < 0 ? 0 : initial_gain > 255 ? 255 : initial_gain;
  int used = gain, iters = 0;
  memset(ix, 0, 576 * sizeof(int32_t));
  for (int it = 0; it < 20; ++it) {
    quantize_with_gain(spec, mag, gain, ix); used = gain; ++iters;
    int last = last_nonzero(ix);
    if (last == 0 && it == 0) { gain = gain - 40 > 0 ? gain - 40 : 0; continue; }   /* SRC:758-761 */
    int significant = (last + 1) & ~1; if (significant > 576) significant = 576;
    int big_values = significant / 2; if (big_values > 288) big_values = 288;
    int est = count_bits15(ix, big_values * 2);
    if (est <= max_bits) break;
    gain = gain + 4 < 255 ? gain + 4 : 255;                        /* SRC:772-775 */
    if (gain >= 255) break;
  }
  int last = last_nonzero(ix);
  int significant = (last + 1) & ~1; if (significant > 576) significant = 576;
  int big_values = significant / 2; if (big_values > 288) big_values = 288;
  *bits_out = encode_table15(ix, big_values * 2, w);
  *gain_out = gain; *gain_used = used; *iterations = iters;
}

/* calculateRegionCounts SRC:856-887 */
static void region_counts(int big_values, int sample_rate, int *r0, int *r1) {
  int region = big_values * 2;
  const uint8_t *bands = band_table(sample_rate);
  int boundaries[21], cum = 0;
  for (int i = 0; i < 21; ++i) { cum += bands[i]; boundaries[i] = cum; }
  int region0 = 0;
  for (int i = 0; i < 15; ++i) { if (boundaries[i] <= region) region0 = i; else break; }
  int region1 = 0, start = region0 + 1;
  int lim = start + 7 < 21 ? start + 7 : 21;
  for (int i = start; i < lim; ++i) { if (boundaries[i] <= region) region1 = i - region0 - 1; else break; }
  *r0 = region0 < 15 ? region0 : 15; *r1 = region1 < 7 ? region1 : 7;
}

/* PreEmphasis.shouldEnable SRC:2042-2066 (scalefactor average is always 1.0 > 0.5) */
static int preflag(const float *spec) {
  float high = sumsq(spec + 432, 144);
  float low = sumsq(spec, 432);
  return high > low * 1.5f;
}

typedef struct {
  int part23, big_values, global_gain, window_switching, block_type, mixed, sbg[3], region0, region1, preflag;
} granule_info;                                                   /* GranuleInfo SRC:2070-2085 (constant fields implied) */

/* VBRState.chooseBitrate SRC:1177-1189 */
static int choose_bitrate(const orc_session *s, int base, float e, int quality) {
  float average;
  if (s->vbr_n == 0) average = e;
  else { float sum = 0.0f; for (int i = 0; i < s->vbr_n; ++i) sum = sum + s->vbr_energy[i]; average = sum / (float)s->vbr_n; }
  float ratio = fminf(fmaxf(e / fmaxf(average, 0.0001f), 0.5f), 2.0f);
  float quality_factor = (float)(9 - quality) / 9.0f;
  int max_adjustment = (int)(32.0f + 32.0f * quality_factor);
  int adjustment = (int)((ratio - 1.0f) * (float)max_adjustment);
  int min_bitrate = base - 64 + quality * 8; if (min_bitrate < 32) min_bitrate = 32;
  int max_bitrate = base + 64 - quality * 4; if (max_bitrate > 320) max_bitrate = 320;
  int v = base + adjustment; if (v > max_bitrate) v = max_bitrate;
  return v > min_bitrate ? v : min_bitrate;
}
static void vbr_update(orc_session *s, float e) {                 /* SRC:1144-1153 */
  if (s->vbr_n == 10) { memmove(s->vbr_energy, s->vbr_energy + 1, 9 * sizeof(float)); s->vbr_n = 9; }
  s->vbr_energy[s->vbr_n++] = e;
}

/* buildMainData SRC:628-731 */
static void build_main_data(orc_session *s, const float *frame, int target_main_data_size, int reservoir_bits,
                            bytes_t *huff, granule_info gi[2][2], orc_frame_trace *ft, int frame_no) {
  int ch_n = s->channels;
  float chan[2][1152];
  if (ch_n == 1) memcpy(chan[0], frame, 1152 * sizeof(float));     /* deinterleave SRC:890-914 */
  else for (int i = 0; i < 1152; ++i) { chan[0][i] = frame[2 * i]; chan[1][i] = frame[2 * i + 1]; }
  int ms = 0;
  if (s->opt.mode == 2) {                                          /* StereoDecision.make SRC:2140-2162 */
    static __thread float mid[1152], side[1152];
    for (int i = 0; i < 1152; ++i) {
      mid[i] = (chan[0][i] + chan[1][i]) * 0.5f;                   /* SRC:2148-2150 */
      side[i] = (chan[0][i] - chan[1][i]) * 0.5f;                  /* SRC:2153-2154: vDSP_vsub(B,A) = A - B → left - right */
    }
    float me = energy(mid, 1152), se = energy(side, 1152);
    if (se < me * 0.4f) { ms = 1; memcpy(chan[0], mid, sizeof mid); memcpy(chan[1], side, sizeof side); }
  }
  if (ft) ft->ms = ms;

  bitw_t w; memset(&w, 0, sizeof w);
  int usable = (reservoir_bits * 9) / 10;                          /* SRC:647 */
  int total_bits = target_main_data_size * 8 + usable;
  int bits_per_granule = total_bits / (2 * ch_n);                  /* SRC:650 */

  for (int gr = 0; gr < 2; ++gr)
    for (int ch = 0; ch < ch_n; ++ch) {
      const float *g = chan[ch] + gr * 576;
      /* encodeSpectrum SRC:947-986 */
      float sub[32][18];
      for (int t = 0; t < 18; ++t) {                               /* analyzeSubbands SRC:917-944 */
        float o[32];
        filterbank_step(g + 32 * t, s->fb[ch], o);
        for (int sb = 0; sb < 32; ++sb) sub[sb][t] = o[sb];
      }
      float e3[3]; int bt, sbg[3];
      transient(g, e3, &bt, sbg);
      float spec[576], thr[576];
      mdct_apply(sub, s->overlap[ch], bt, spec);
      masking_thresholds(spec, s->opt.sample_rate, s->opt.quality, thr);
      int g0 = compute_global_gain(spec);
      float ge = energy(g, 576);
      vbr_update(s, ge);                                           /* SRC:671-674 */
      int gain, bits, used, iters; int32_t ix[576];
      quantize_to_fit(spec, g0, bits_per_granule, &w, &gain, ix, &bits, &used, &iters);
      int pf = preflag(spec);
      int last = last_nonzero(ix);                                 /* SRC:692-700 */
      int big_values = ((last + 1) & ~1) / 2; if (big_values > 288) big_values = 288;
      int r0, r1; region_counts(big_values, s->opt.sample_rate, &r0, &r1);
      granule_info *q = &gi[gr][ch];
      q->part23 = bits; q->big_values = big_values; q->global_gain = gain;
      q->window_switching = bt != 0; q->block_type = bt; q->mixed = bt == 1;
      memcpy(q->sbg, sbg, sizeof sbg); q->region0 = r0; q->region1 = r1; q->preflag = pf;
      if (s->trace_on) {
        orc_gc_trace *t = trace_gc_new(s);
        memcpy(t->spectrum, spec, sizeof spec);
        for (int sb = 0; sb < 32; ++sb) memcpy(t->subband + sb * 18, sub[sb], 18 * sizeof(float));
        memcpy(t->thresholds, thr, sizeof thr); memcpy(t->ix, ix, sizeof ix);
        t->energy = ge; memcpy(t->sub_energy, e3, sizeof e3);
        t->block_type = bt; t->mixed = bt == 1; t->window_switching = bt != 0; memcpy(t->subblock_gain, sbg, sizeof sbg);
        t->g0 = g0; t->gain_out = gain; t->gain_used = used; t->iterations = iters; t->bits = bits;
        t->max_bits = bits_per_granule; t->big_values = big_values; t->region0 = r0; t->region1 = r1; t->preflag = pf;
        t->frame = frame_no; t->gr = gr; t->ch = ch;
      }
    }
  bw_pad(&w);                                                      /* SRC:729 */
  *huff = w.data;
}

/* buildSideInfo SRC:571-625 */
static void build_side_info(int ch_n, granule_info gi[2][2], int main_data_begin, bytes_t *out) {
  bitw_t w; memset(&w, 0, sizeof w);
  bw_write(&w, main_data_begin < 511 ? main_data_begin : 511, 9);
  bw_write(&w, 0, ch_n == 1 ? 5 : 3);
  for (int ch = 0; ch < ch_n; ++ch) for (int b = 0; b < 4; ++b) bw_write(&w, 0, 1);   /* scfsi all zero SRC:644 */
  for (int gr = 0; gr < 2; ++gr)
    for (int ch = 0; ch < ch_n; ++ch) {
      const granule_info *q = &gi[gr][ch];
      bw_write(&w, q->part23, 12); bw_write(&w, q->big_values, 9); bw_write(&w, q->global_gain, 8);
      bw_write(&w, 0, 4);                                          /* scalefac_compress SRC:685 */
      bw_write(&w, q->window_switching, 1);
      if (q->window_switching == 1) {
        bw_write(&w, q->block_type, 2); bw_write(&w, q->mixed, 1);
        bw_write(&w, 15, 5); bw_write(&w, 15, 5);                  /* table_select SRC:717 */
        bw_write(&w, q->sbg[0], 3); bw_write(&w, q->sbg[1], 3); bw_write(&w, q->sbg[2], 3);
      } else {
        bw_write(&w, 15, 5); bw_write(&w, 15, 5); bw_write(&w, 15, 5);
        bw_write(&w, q->region0, 4); bw_write(&w, q->region1, 3);
      }
      bw_write(&w, q->preflag, 1); bw_write(&w, 0, 1); bw_write(&w, 0, 1);
    }
  bw_pad(&w);
  size_t want = (size_t)(ch_n == 1 ? 136 : 256) / 8;
  by_append(out, w.data.p, w.data.n);
  if (w.data.n < want) by_zeros(out, want - w.data.n);
  by_free(&w.data);
}

static uint16_t crc16_mpeg(const uint8_t *p, size_t n) {          /* SRC:2208-2215 */
  uint16_t crc = 0xFFFF;
  for (size_t i = 0; i < n; ++i) crc = (uint16_t)((crc << 8) ^ T_crc[((crc >> 8) ^ p[i]) & 0xFF]);
  return crc;
}
static void mode_bits(int mode, int *mb, int *me) {               /* SRC:2547-2556 */
  if (mode == 0) { *mb = 3; *me = 0; } else if (mode == 2) { *mb = 1; *me = 2; } else { *mb = 0; *me = 0; }
}
static void write_header(bitw_t *h, int protection, int bitrate_index, int sr_index, int padding, int mode,
                         int copyright, int original) {           /* SRC:523-536 / 379-392 */
  int mb, me; mode_bits(mode, &mb, &me);
  bw_write(h, 0x7FF, 11); bw_write(h, 3, 2); bw_write(h, 1, 2); bw_write(h, protection, 1);
  bw_write(h, bitrate_index, 4); bw_write(h, sr_index, 2); bw_write(h, padding, 1); bw_write(h, 0, 1);
  bw_write(h, mb, 2); bw_write(h, me, 2); bw_write(h, copyright, 1); bw_write(h, original, 1); bw_write(h, 0, 2);
}

/* BitReservoir.fillSlot SRC:2110-2121 */
static void fill_slot(orc_session *s, int slot, bytes_t *out) {
  if (slot <= 0) return;
  if (s->res_stream.n >= (size_t)slot) { by_append(out, s->res_stream.p, slot); by_drop_front(&s->res_stream, slot); }
  else { size_t have = s->res_stream.n; by_append(out, s->res_stream.p, have); by_zeros(out, slot - have); s->res_stream.n = 0; }
}
static void note_frame_size(orc_session *s, int n) {
  if (s->fs_n == s->fs_cap) { s->fs_cap = s->fs_cap ? s->fs_cap * 2 : 1024; s->frame_sizes = (int *)realloc(s->frame_sizes, s->fs_cap * sizeof(int)); }
  s->frame_sizes[s->fs_n++] = n;
}

/* encodeFrame SRC:475-568; appends the emitted (previous) frame to s->out */
static void encode_frame(orc_session *s, const float *frame, int is_final) {
  int ch_n = s->channels;
  float frame_energy = energy(frame, 1152 * ch_n);                 /* SRC:477 */
  int target = s->opt.vbr ? choose_bitrate(s, s->opt.bitrate_kbps, frame_energy, s->opt.quality) : s->opt.bitrate_kbps;
  int br_index = orc_bitrate_index(target, s->opt.sample_rate);
  int sr_index = sample_rate_index(s->opt.sample_rate);
  int br_value = orc_bitrate_value(br_index);
  int side_size = ch_n == 1 ? 17 : 32, crc_size = s->opt.crc_protected ? 2 : 0;
  int numerator = 144 * br_value * 1000;                           /* SRC:490-496 */
  int base_size = numerator / s->opt.sample_rate, remainder = numerator % s->opt.sample_rate;
  int padding = 0;
  s->padding_remainder += remainder;                               /* shouldPad SRC:456-463 */
  if (s->padding_remainder >= s->opt.sample_rate) { s->padding_remainder -= s->opt.sample_rate; padding = 1; }
  int frame_size = base_size + padding;
  int main_data_size = frame_size - 4 - crc_size - side_size;
  int mdb = is_final ? 0 : (s->res_stream.n < 511 ? (int)s->res_stream.n : 511);   /* SRC:499, 2099-2101 */
  int res_bits = is_final ? 0 : s->res_available * 8;              /* SRC:500 */

  orc_frame_trace *ft = s->trace_on ? trace_frame_new(s) : NULL;
  int frame_no = (int)(s->trace_on ? s->tf_n - 1 : 0);
  bytes_t huff = {0}; granule_info gi[2][2]; memset(gi, 0, sizeof gi);
  build_main_data(s, frame, main_data_size, res_bits, &huff, gi, ft, frame_no);
  if (s->trace_on) ft = &s->tf[s->tf_n - 1];
  by_append(&s->res_stream, huff.p, huff.n);                       /* SRC:511 */

  bytes_t hs = {0};
  bitw_t h; memset(&h, 0, sizeof h);
  write_header(&h, s->opt.crc_protected ? 0 : 1, br_index, sr_index, padding, s->opt.mode, s->opt.copyright ? 1 : 0,
               s->opt.original ? 1 : 0);
  by_append(&hs, h.data.p, h.data.n); by_free(&h.data);
  if (s->opt.crc_protected) {                                      /* SRC:540-543: CRC over the 4 header bytes */
    uint16_t crc = crc16_mpeg(hs.p, hs.n);
    by_push(&hs, (uint8_t)(crc >> 8)); by_push(&hs, (uint8_t)(crc & 0xFF));
  }
  build_side_info(ch_n, gi, mdb, &hs);

  if (s->buffered.present) {                                       /* SRC:548-556 */
    size_t before = s->out.n;
    by_append(&s->out, s->buffered.header_and_side.p, s->buffered.header_and_side.n);
    fill_slot(s, s->buffered.slot_size, &s->out);
    int emitted = (int)(s->out.n - before);
    s->frame_count += 1; s->total_bytes += (uint32_t)emitted; note_frame_size(s, emitted);
  }
  by_free(&s->buffered.header_and_side);                           /* SRC:559-562 */
  s->buffered.header_and_side = hs; s->buffered.slot_size = main_data_size; s->buffered.present = 1;
  s->res_available += main_data_size - (int)huff.n;                /* SRC:565, 2125-2128 */
  if (s->res_available < 0) s->res_available = 0;
  if (s->res_available > 511) s->res_available = 511;
  if (ft) {
    ft->frame_energy = frame_energy; ft->bitrate_kbps = br_value; ft->bitrate_index = br_index; ft->padding = padding;
    ft->frame_size = frame_size; ft->main_data_size = main_data_size; ft->main_data_begin = mdb;
    ft->reservoir_bits = res_bits; ft->huff_bytes = (int)huff.n; ft->is_final = is_final;
  }
  by_free(&huff);
}

/* EncoderSession.encode(samples:) SRC:297-310 */
const uint8_t *orc_encode(orc_session *s, const float *pcm, size_t n_floats, size_t *out_len) {
  s->out.n = 0;
  if (s->pcm_n + n_floats > s->pcm_cap) { s->pcm_cap = (s->pcm_n + n_floats) * 2 + 4096; s->pcm = (float *)realloc(s->pcm, s->pcm_cap * sizeof(float)); }
  if (n_floats) memcpy(s->pcm + s->pcm_n, pcm, n_floats * sizeof(float));
  s->pcm_n += n_floats;
  size_t fsc = (size_t)1152 * s->channels, pos = 0;
  while (s->pcm_n - pos >= fsc) { encode_frame(s, s->pcm + pos, 0); pos += fsc; }
  if (pos) { memmove(s->pcm, s->pcm + pos, (s->pcm_n - pos) * sizeof(float)); s->pcm_n -= pos; }  /* SRC:305 without the quadratic cost */
  *out_len = s->out.n;
  return s->out.p;
}
/* EncoderSession.flush() SRC:318-350 */
const uint8_t *orc_flush(orc_session *s, size_t *out_len) {
  s->out.n = 0;
  if (s->pcm_n) {
    size_t fsc = (size_t)1152 * s->channels;
    float *frame = (float *)calloc(fsc, sizeof(float));
    memcpy(frame, s->pcm, s->pcm_n * sizeof(float));               /* SRC:325-328 (pcm_n < fsc here) */
    s->pcm_n = 0;
    encode_frame(s, frame, 1);
    free(frame);
  }
  if (s->buffered.present) {                                       /* SRC:335-347 */
    size_t before = s->out.n;
    by_append(&s->out, s->buffered.header_and_side.p, s->buffered.header_and_side.n);
    fill_slot(s, s->buffered.slot_size, &s->out);
    int emitted = (int)(s->out.n - before);
    s->frame_count += 1; s->total_bytes += (uint32_t)emitted; note_frame_size(s, emitted);
    by_free(&s->buffered.header_and_side); s->buffered.present = 0;
  }
  *out_len = s->out.n;
  return s->out.p;
}

/* generateXingHeader SRC:367-420 + generateTOC SRC:423-449 */
const uint8_t *orc_xing_header(orc_session *s, size_t *out_len) {
  s->out.n = 0;
  int side = s->channels == 1 ? 17 : 32;
  int br_index = orc_bitrate_index(s->opt.bitrate_kbps, s->opt.sample_rate);
  int sr_index = sample_rate_index(s->opt.sample_rate);
  int frame_size = (144 * orc_bitrate_value(br_index) * 1000) / s->opt.sample_rate;
  bitw_t h; memset(&h, 0, sizeof h);
  write_header(&h, 1, br_index, sr_index, 0, s->opt.mode, 0, 1);
  by_append(&s->out, h.data.p, h.data.n); by_free(&h.data);
  by_zeros(&s->out, side);
  by_append(&s->out, (const uint8_t *)(s->opt.vbr ? "Xing" : "Info"), 4);
  uint32_t words[3] = {0x07u, s->frame_count + 1u, s->total_bytes + (uint32_t)frame_size};
  for (int i = 0; i < 3; ++i) { uint8_t b[4] = {(uint8_t)(words[i] >> 24), (uint8_t)(words[i] >> 16), (uint8_t)(words[i] >> 8), (uint8_t)words[i]}; by_append(&s->out, b, 4); }
  long total = 0;
  for (size_t i = 0; i < s->fs_n; ++i) total += s->frame_sizes[i];
  if (s->fs_n == 0 || total <= 0) { for (int p = 0; p < 100; ++p) by_push(&s->out, (uint8_t)(p * 255 / 99)); }
  else {
    long *cum = (long *)malloc(s->fs_n * sizeof(long)); long sum = 0;
    for (size_t i = 0; i < s->fs_n; ++i) { sum += s->frame_sizes[i]; cum[i] = sum; }
    for (int p = 0; p < 100; ++p) {
      size_t target = ((size_t)p * s->fs_n) / 100;
      long pos = target > 0 ? cum[target - 1] : 0;
      long scaled = pos * 255 / total;
      by_push(&s->out, (uint8_t)(scaled < 255 ? scaled : 255));
    }
    free(cum);
  }
  if ((int)s->out.n < frame_size) by_zeros(&s->out, frame_size - s->out.n);
  *out_len = s->out.n;
  return s->out.p;
}

/* ID3TagWriter SRC:1037-1136 */
static void id3_frame_header(bytes_t *b, const char *id, uint32_t size) {   /* SRC:1127-1135 */
  by_append(b, (const uint8_t *)id, 4);
  uint8_t h[6] = {(uint8_t)(size >> 24), (uint8_t)(size >> 16), (uint8_t)(size >> 8), (uint8_t)size, 0, 0};
  by_append(b, h, 6);
}
static void id3_text(bytes_t *b, const char *id, const char *v) {  /* SRC:1078-1086 */
  size_t n = strlen(v);
  id3_frame_header(b, id, (uint32_t)(1 + n)); by_push(b, 0x03); by_append(b, (const uint8_t *)v, n);
}
uint8_t *orc_id3_build(const orc_id3 *tag, size_t *out_len) {      /* SRC:1040-1075 */
  bytes_t f = {0}; char num[32];
  if (tag->title) id3_text(&f, "TIT2", tag->title);
  if (tag->artist) id3_text(&f, "TPE1", tag->artist);
  if (tag->album) id3_text(&f, "TALB", tag->album);
  if (tag->genre) id3_text(&f, "TCON", tag->genre);
  if (tag->year >= 0) { snprintf(num, sizeof num, "%d", tag->year); id3_text(&f, "TYER", num); }
  if (tag->track >= 0) {
    if (tag->track_total >= 0) snprintf(num, sizeof num, "%d/%d", tag->track, tag->track_total); else snprintf(num, sizeof num, "%d", tag->track);
    id3_text(&f, "TRCK", num);
  }
  if (tag->comment) {                                              /* SRC:1089-1099 */
    size_t n = strlen(tag->comment);
    id3_frame_header(&f, "COMM", (uint32_t)(1 + 3 + 1 + n));
    by_push(&f, 0x03); by_append(&f, (const uint8_t *)"eng", 3); by_push(&f, 0); by_append(&f, (const uint8_t *)tag->comment, n);
  }
  if (tag->album_art) {                                            /* SRC:1102-1114 */
    const char *mime = tag->album_art_mime ? tag->album_art_mime : "image/jpeg"; size_t mn = strlen(mime);
    id3_frame_header(&f, "APIC", (uint32_t)(1 + mn + 1 + 1 + 1 + tag->album_art_len));
    by_push(&f, 0x03); by_append(&f, (const uint8_t *)mime, mn); by_push(&f, 0); by_push(&f, 0x03); by_push(&f, 0);
    by_append(&f, tag->album_art, tag->album_art_len);
  }
  if (f.n == 0) { *out_len = 0; by_free(&f); return NULL; }
  bytes_t o = {0};
  uint32_t sz = (uint32_t)f.n;
  uint8_t hdr[10] = {0x49, 0x44, 0x33, 0x03, 0x00, 0x00, (uint8_t)((sz >> 21) & 0x7F), (uint8_t)((sz >> 14) & 0x7F),
                     (uint8_t)((sz >> 7) & 0x7F), (uint8_t)(sz & 0x7F)};
  by_append(&o, hdr, 10); by_append(&o, f.p, f.n); by_free(&f);
  *out_len = o.n;
  return o.p;
}
void orc_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------------ */
/* Synthetic PCM, CPU twin of the product's mp3b_synth_fill (swift-mp3_b200/csrc/kernels.cu, k_synth): the BASELINE C1/C4
 * recipe a*sin(2 pi f t) + noise*N(0,1) from a counter-based generator, every step an IEEE-754 double operation with
 * one rounding (built with -ffp-contract=off; fma() is the fused operation), so both sides produce the same floats.    */
static uint64_t sm64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
static double synth_ln(uint64_t k) {                                 /* ln(k / 2^53), 1 <= k <= 2^53 */
  int e = 63 - __builtin_clzll(k);
  double m = ldexp((double)k, -e);                                   /* exact, [1, 2) */
  if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
  const double s = (m - 1.0) / (m + 1.0), z = s * s;
  static const double odd[11] = {1.0, 1.0 / 3.0, 1.0 / 5.0, 1.0 / 7.0, 1.0 / 9.0, 1.0 / 11.0, 1.0 / 13.0, 1.0 / 15.0, 1.0 / 17.0, 1.0 / 19.0, 1.0 / 21.0};
  double p = odd[10];
  for (int i = 9; i >= 0; --i) p = fma(p, z, odd[i]);
  return fma((double)(e - 53), 0.6931471805599453, (2.0 * s) * p);
}
static void synth_sincos(double t, double *sn, double *cs) {         /* (sin, cos)(2 pi t), 0 <= t < 1 */
  static const double fs[12] = {1.0, -1.0 / 6.0, 1.0 / 120.0, -1.0 / 5040.0, 1.0 / 362880.0, -1.0 / 39916800.0, 1.0 / 6227020800.0,
                                -1.0 / 1307674368000.0, 1.0 / 355687428096000.0, -1.0 / 121645100408832000.0,
                                1.0 / 51090942171709440000.0, -1.0 / 25852016738884976640000.0};
  static const double fc[13] = {1.0, -0.5, 1.0 / 24.0, -1.0 / 720.0, 1.0 / 40320.0, -1.0 / 3628800.0, 1.0 / 479001600.0,
                                -1.0 / 87178291200.0, 1.0 / 20922789888000.0, -1.0 / 6402373705728000.0,
                                1.0 / 2432902008176640000.0, -1.0 / 1124000727777607680000.0, 1.0 / 620448401733239439360000.0};
  const double q4 = floor(t * 4.0);
  const double x = 6.283185307179586 * (t - q4 * 0.25), z = x * x;
  double ps = fs[11], pc = fc[12];
  for (int i = 10; i >= 0; --i) ps = fma(ps, z, fs[i]);
  for (int i = 11; i >= 0; --i) pc = fma(pc, z, fc[i]);
  const double s0 = x * ps;
  switch ((int)q4) {
    case 0: *sn = s0; *cs = pc; break;
    case 1: *sn = pc; *cs = -s0; break;
    case 2: *sn = -s0; *cs = -pc; break;
    default: *sn = -pc; *cs = s0; break;
  }
}
void orc_synth_fill(float *pcm, size_t n_per_channel, int channels, int sample_rate, float f_left, float f_right,
                    float amp, float noise, uint64_t seed) {
  for (size_t i = 0; i < n_per_channel; ++i) {
    const uint64_t h1 = sm64(seed * 0x100000001B3ull + i), h2 = sm64(h1);
    const double rad = sqrt(-2.0 * synth_ln((h1 >> 11) + 1));
    double g[2];                                                     /* g[0] = cos part (left), g[1] = sin part (right) */
    synth_sincos((double)(h2 >> 11) * 1.1102230246251565e-16, &g[1], &g[0]);
    for (int c = 0; c < channels; ++c) {
      const double cyc = ((double)(c == 0 ? f_left : f_right) * (double)i) / (double)sample_rate;
      double t = cyc - floor(cyc);
      if (c == 1) { t = t + 0.0477464829275686; if (t >= 1.0) t = t - 1.0; }
      double sn, cs;
      synth_sincos(t, &sn, &cs);
      const double v = (double)amp * sn + (double)noise * (rad * g[c]);
      float f = (float)v;
      pcm[i * (size_t)channels + c] = f < -1.0f ? -1.0f : f > 1.0f ? 1.0f : f;
    }
  }
}

/* ------------------------------------------------------------------------------------------------ */
/* multi-threaded CPU-baseline driver                                                               */
typedef struct {
  const orc_options *opts; const float *const *pcm; const size_t *n_floats; size_t n_streams;
  size_t next; pthread_mutex_t mu; size_t bytes; uint64_t digest;
  const uint8_t *const *expect; const size_t *expect_len; int64_t *first_diff; size_t chunk_floats;   /* compare mode */
} job_t;
static uint64_t fnv1a(const uint8_t *p, size_t n, uint64_t h) { for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; } return h; }
static void *worker(void *arg) {
  job_t *j = (job_t *)arg;
  for (;;) {
    pthread_mutex_lock(&j->mu); size_t i = j->next++; pthread_mutex_unlock(&j->mu);
    if (i >= j->n_streams) break;
    orc_session *s = orc_create(j->opts);
    size_t n1 = 0, n2; uint64_t h = 1469598103934665603ull;
    int64_t diff = -1;                       /* compare mode: offset of the first byte that differs from expect[i], -1 = identical */
    const uint8_t *ex = j->expect ? j->expect[i] : NULL; const size_t exn = j->expect ? j->expect_len[i] : 0;
    /* one encode(samples:) call, or — chunk_floats > 0 — the stream fed chunk by chunk like a streaming client (SRC:297-310) */
    const size_t total = j->n_floats[i], step = j->chunk_floats ? j->chunk_floats : (total ? total : 1);
    for (size_t off = 0; off < total || off == 0; off += step) {
      size_t nn; const size_t take = total - off < step ? total - off : step;
      const uint8_t *p = orc_encode(s, j->pcm[i] + off, take, &nn); h = fnv1a(p, nn, h);
      if (j->expect && diff < 0) for (size_t k = 0; k < nn; ++k) if (n1 + k >= exn || ex[n1 + k] != p[k]) { diff = (int64_t)(n1 + k); break; }
      n1 += nn;
      if (total == 0) break;
    }
    const uint8_t *p = orc_flush(s, &n2); h = fnv1a(p, n2, h);
    if (j->expect && diff < 0) for (size_t k = 0; k < n2; ++k) if (n1 + k >= exn || ex[n1 + k] != p[k]) { diff = (int64_t)(n1 + k); break; }
    if (j->expect && diff < 0 && n1 + n2 != exn) diff = (int64_t)(n1 + n2);
    if (j->first_diff) j->first_diff[i] = diff;
    orc_destroy(s);
    pthread_mutex_lock(&j->mu); j->bytes += n1 + n2; j->digest += h * (2 * (uint64_t)i + 1); pthread_mutex_unlock(&j->mu);
  }
  return NULL;
}
static void run_job(job_t *j, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
  for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, worker, j);
  for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
  free(th);
}
size_t orc_encode_streams(const orc_options *opts, const float *const *pcm, const size_t *n_floats, size_t n_streams,
                          int n_threads, uint64_t *out_digest) {
  job_t j = {opts, pcm, n_floats, n_streams, 0, PTHREAD_MUTEX_INITIALIZER, 0, 0, NULL, NULL, NULL, 0};
  run_job(&j, n_threads);
  if (out_digest) *out_digest = j.digest;
  return j.bytes;
}

/* Parity driver: encodes every stream (encode + flush, or chunk_floats at a time) and compares the bytes with
 * expect[i][0 .. expect_len[i]) — the output of the implementation under test.  first_diff[i] = -1 when identical, else the
 * offset of the first differing byte (or the shorter length).  Returns the number of streams that differ. */
size_t orc_compare_streams(const orc_options *opts, const float *const *pcm, const size_t *n_floats, size_t n_streams,
                           size_t chunk_floats, int n_threads, const uint8_t *const *expect, const size_t *expect_len,
                           int64_t *first_diff) {
  int64_t *fd = first_diff ? first_diff : (int64_t *)malloc(sizeof(int64_t) * (n_streams ? n_streams : 1));
  job_t j = {opts, pcm, n_floats, n_streams, 0, PTHREAD_MUTEX_INITIALIZER, 0, 0, expect, expect_len, fd, chunk_floats};
  run_job(&j, n_threads);
  size_t bad = 0;
  for (size_t i = 0; i < n_streams; ++i) bad += fd[i] >= 0;
  if (!first_diff) free(fd);
  return bad;
}

#ifdef ORC_VARIANTS
/* quantizeWithGain of one spectrum at a given gain, with the variant's arithmetic (tools/order_sensitivity.py). */
void orc_requantize(const float *spec, int gain, int32_t *ix) {
  tables();
  float mag[576];
  for (int i = 0; i < 576; ++i) mag[i] = orc_pow34(fmaxf(fabsf(spec[i]), 1e-10f));
  quantize_with_gain(spec, mag, gain < 0 ? 0 : gain > 255 ? 255 : gain, ix);
}
#endif
