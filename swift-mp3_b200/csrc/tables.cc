#include "tables.h"

#include <cmath>
#include <cstdlib>
#include <mutex>

#include "tables_gen.h"

namespace mp3b {

static float g_inv_step[256], g_inv_step_iso[320];
static double g_gain_thr[256];
static int g_sfb_cum[3][21];
static std::once_flag g_once;

static void build() {
  for (int g = 0; g < 256; ++g) {
    double p = std::pow(2.0, (double)(g - 210) / 4.0);
    g_gain_thr[g] = p;
    float step = (float)(p > 0.0001 ? p : 0.0001);
    g_inv_step[g] = 1.0f / step;
  }
  // ISO mode: ix = nint((|xr| * 32768 / 2^((G - 210) / 4))^0.75 - 0.0946) for the written global_gain G (the decoder's PCM scale
  // is 32768 times the [-1, 1] floats the encoder is fed), i.e. |xr|^0.75 * 2^((180 - 3 (G - 210)) / 16); 320 entries: iso_mode.cuh
  for (int g = 0; g < 320; ++g) g_inv_step_iso[g] = (float)std::pow(2.0, (180.0 - 3.0 * (double)(g - 210)) / 16.0);
  for (int r = 0; r < 3; ++r) { int c = 0; for (int i = 0; i < 21; ++i) { c += tab::kSfbLong[r][i]; g_sfb_cum[r][i] = c; } }
}

const float *host_inv_step() { std::call_once(g_once, build); return g_inv_step; }
const float *host_inv_step_iso() { std::call_once(g_once, build); return g_inv_step_iso; }
const double *host_gain_thr() { std::call_once(g_once, build); return g_gain_thr; }
const uint8_t *host_len15() { return tab::kHuff15Len; }
const uint8_t *host_code15() { return tab::kHuff15Code; }
const int *host_sfb_cum() { std::call_once(g_once, build); return &g_sfb_cum[0][0]; }

static const int kBitrates1[16] = {0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 0};
static const int kBitrates2[16] = {0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160, 0};

int bitrate_index(int bitrate, int sample_rate) {   // exact match, else nearest (first minimum wins), SRC:2516-2521
  const int *t = sample_rate >= 32000 ? kBitrates1 : kBitrates2;
  for (int i = 0; i < 16; ++i) if (t[i] == bitrate) return i;
  int best = 0;
  for (int i = 1; i < 16; ++i) if (std::abs(t[i] - bitrate) < std::abs(t[best] - bitrate)) best = i;
  return best;
}
int bitrate_value(int index) { return (index >= 0 && index < 16) ? kBitrates1[index] : 128; }
int sample_rate_index(int sr) { return sr == 44100 ? 0 : sr == 48000 ? 1 : sr == 32000 ? 2 : 0; }
int sfb_table_index(int sr) { return sr == 48000 ? 1 : sr == 32000 ? 2 : 0; }

}  // namespace mp3b
