#include "tables.h"

#include <cmath>
#include <cstdlib>
#include <mutex>

#include "tables_gen.h"

namespace mp3b {

static float g_inv_step[256];
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
  for (int r = 0; r < 3; ++r) { int c = 0; for (int i = 0; i < 21; ++i) { c += tab::kSfbLong[r][i]; g_sfb_cum[r][i] = c; } }
}

const float *host_inv_step() { std::call_once(g_once, build); return g_inv_step; }
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
