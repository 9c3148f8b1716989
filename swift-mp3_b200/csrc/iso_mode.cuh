// Opt-in ISO mode (north_star stages (4) and (5)): what the reference leaves as dead stubs —
// HuffmanEncoder.encode / writePair / selectTable (SRC:1740-1806), Huffman tables other than 15 (SRC:2288-2398) — done the ISO
// 11172-3 way, so that an independent decoder reconstructs the INPUT, not just "something":
//   * the ISO quantizer  ix = nint((|xr| / 2^((global_gain - 210) / 4))^0.75 - 0.0946)  on the decoder's scale (xr * 32768),
//     no clamp at 15 (ESC tables with linbits carry values up to 8206);
//   * per granule-channel: rzero / count1 / big_values partition, three regions on scalefactor-band boundaries, for every
//     region the cheapest of the ISO tables that can hold its largest value (1-3, 5-13, 15, 16-23, 24-31), count1 table A or B;
//   * global_gain by search: the smallest gain whose bit count fits the granule's budget (binary search for the largest budget
//     the frame can have, then the curve gain by gain down to the smallest — the serial scan picks the entry);
//   * a real main_data_begin (the bit reservoir as a back pointer, with stuffing when it would exceed 511 bytes or one slot).
// Long blocks only in this slice (the reference's short-block line order and its switch without start / stop windows, SURVEY
// Q7-Q10, are not ISO); no scalefactors (part2 = 0 bits), no psychoacoustic model.  One warp per granule-channel, everything
// warp-reduced.  Included by kernels.cu.
#pragma once
#include "iso_huffman.inc"

namespace mp3b {

// Quantizer scale per search gain G: 2^((180 - 3 (G - 210)) / 16) = (32768 / 2^((G - 210) / 4))^0.75.  The side info can
// express G <= 255; the search runs on to kIsoGainMax so that a budget too small even for global_gain 255 (full-scale noise at
// the lowest bitrates) still ends in a valid stream — the written gain is then 255 and the granule decodes too quiet.
constexpr int kIsoGainMax = 319;
__constant__ float c_inv_step_iso[kIsoGainMax + 1];

cudaError_t upload_iso_tables(const float *inv_step_iso) { return cudaMemcpyToSymbol(c_inv_step_iso, inv_step_iso, sizeof(float) * (kIsoGainMax + 1)); }

constexpr int kIsoMaxValue = 8191 + 15;
__device__ __forceinline__ int iso_quant(float mag /* |xr|^0.75 */, float inv) {
  return min(__float2int_rd(__fadd_rn(__fmul_rn(mag, inv), 0.4054f)), kIsoMaxValue);      // nint(t - 0.0946) = floor(t + 0.4054)
}

// up to three candidate tables for a region whose largest value is m (ISO 11172-3 Table B.7: tables grouped by their range)
__device__ __forceinline__ void iso_candidates(int m, int c[3]) {
  c[0] = 0; c[1] = c[2] = -1;
  if (m == 0) return;
  if (m == 1) { c[0] = 1; return; }
  if (m == 2) { c[0] = 2; c[1] = 3; return; }
  if (m == 3) { c[0] = 5; c[1] = 6; return; }
  if (m <= 5) { c[0] = 7; c[1] = 8; c[2] = 9; return; }
  if (m <= 7) { c[0] = 10; c[1] = 11; c[2] = 12; return; }
  if (m <= 15) { c[0] = 13; c[1] = 15; return; }
  const int need = m - 15;                                         // linbits must hold it: need < 2^linbits
  int a = 0, b = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if ((1 << ((kLinbits16Packed >> (4 * i)) & 15u)) <= need) a = i + 1;
    if ((1 << ((kLinbits24Packed >> (4 * i)) & 15u)) <= need) b = i + 1;
  }
  c[0] = 16 + min(a, 7); c[1] = 24 + min(b, 7);
}
__device__ __forceinline__ uint32_t iso_desc(int t) {             // first entry | row length << 16 | linbits << 24; 0xFFFFFFFF = no table
  return t < 0 ? 0xFFFFFFFFu : (uint32_t)kHuffBaseC[t] | (uint32_t)kHuffDimC[t] << 16 | (uint32_t)kHuffLinbitsC[t] << 24;
}
__device__ __forceinline__ int quad_len_a(int idx) { return (int)((kQuadLenAPacked >> (4 * idx)) & 15ull); }

struct IsoChoice {            // everything the side info and the bit packer need to know about one granule-channel at one gain
  int bits;                   // part2_3_length (part2 = 0)
  int bv, c1;                 // big_values (pairs), count1 (quadruples)
  int c1sel;                  // count1table_select
  int r0, r1;                 // region0_count, region1_count
  int a1, a2;                 // first line of region 1 / region 2
  int tsel[3];
};

// Bit count and all choices for the quantized pairs (qx[j], qy[j]) = lines 2 p, 2 p + 1 of pair p = lane + 32 j.
// s_len: the concatenated length tables in shared memory; s_c: 288 bytes of warp scratch; sfb: cumulative band ends (21).
// ws: the granule is window-switched (block type start / short / stop): the side info then has two regions only, region 0 = the
// first 36 lines (ISO 11172-3 2.4.2.7: region0_count 7 resp. 8 and region1_count 13 are implied), two table_selects.
// Q: q(j, x, y) delivers the quantized pair lane + 32 j — from registers (U = 9: fully unrolled, the bit packer's single call) or from
// the warp's scratch in shared memory (U = 3: the search function, whose code has to stay small enough for the instruction cache:
// fully unrolled it is 36 KB that every one of ~40 calls per granule-channel streams through).
template <int U, class Q> __device__ __forceinline__ IsoChoice iso_evaluate_core(Q q, int lane, const uint8_t *s_len, uint8_t *s_c, const int *sfb, bool ws) {
  IsoChoice ch;
  int top = 0, big = 0;
#pragma unroll U
  for (int j = 0; j < 9; ++j) {
    const int p = lane + 32 * j;
    int qx, qy; q(j, qx, qy);
    if (qx | qy) top = p + 1;
    if (qx > 1 || qy > 1) big = p + 1;
    s_c[p] = (uint8_t)((qx & 1) << 1 | (qy & 1));                  // only read for pairs of the count1 region (values 0 / 1)
  }
  top = warp_max_i(top); big = warp_max_i(big);
  ch.c1 = (top - big) >> 1;                                        // quadruples of |value| <= 1 from the top down
  ch.bv = top - 2 * ch.c1;
  const int bv2 = 2 * ch.bv;
  // three regions on band boundaries: a third of the bands that lie inside the big_values region each (ISO leaves the split
  // to the encoder; region0_count <= 15, region1_count <= 7)
  int nb = 0;
#pragma unroll
  for (int i = 0; i < 21; ++i) nb += sfb[i] <= bv2;
  const int k0 = min(max((nb + 1) / 3, 1), 16), k1 = min(max((nb + 1) / 3, 1), 8);
  ch.r0 = k0 - 1; ch.r1 = k1 - 1;
  ch.a1 = sfb[k0 - 1]; ch.a2 = k0 + k1 - 1 < 21 ? sfb[k0 + k1 - 1] : 576;
  if (ws) { ch.r0 = 0; ch.r1 = 0; ch.a1 = 36; ch.a2 = 576; }
  int m0 = 0, m1 = 0, m2 = 0;
#pragma unroll U
  for (int j = 0; j < 9; ++j) {
    int qx, qy; q(j, qx, qy);
    const int p = lane + 32 * j, m = max(qx, qy);
    if (p < ch.bv) { if (2 * p < ch.a1) m0 = max(m0, m); else if (2 * p < ch.a2) m1 = max(m1, m); else m2 = max(m2, m); }
  }
  m0 = warp_max_i(m0); m1 = warp_max_i(m1); m2 = warp_max_i(m2);
  int cand[3][3];
  iso_candidates(m0, cand[0]); iso_candidates(m1, cand[1]); iso_candidates(m2, cand[2]);
  uint32_t d[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int k = 0; k < 3; ++k) d[r][k] = iso_desc(cand[r][k]);
  // per candidate k: code lengths summed per region in 16-bit fields of a 64-bit accumulator (a region holds at most 288 pairs of
  // at most 19 bits); values >= 15 per region (10-bit fields) for the linbits; sign bits; count1 bits with table A and with table B
  unsigned long long acc[3] = {0ull, 0ull, 0ull};
  uint32_t n15 = 0, misc = 0;                                       // misc: signs of big values (<= 576) | count1 A bits << 10 (<= 1440) | count1 B bits << 21 (<= 1152)
#pragma unroll U
  for (int j = 0; j < 9; ++j) {
    const int p = lane + 32 * j;
    if (p < ch.bv) {
      int qx, qy; q(j, qx, qy);
      const int r = (2 * p >= ch.a1) + (2 * p >= ch.a2);
      const int cx = min(qx, 15), cy = min(qy, 15);
      n15 += (uint32_t)((qx >= 15) + (qy >= 15)) << (10 * r);
      misc += (uint32_t)((qx != 0) + (qy != 0));
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const uint32_t dk = r == 0 ? d[0][k] : r == 1 ? d[1][k] : d[2][k];
        if (dk != 0xFFFFFFFFu && (dk & 0xFFFFu) + ((dk >> 16) & 255u) != 0u)      // (table 0: nothing is coded)
          acc[k] += (unsigned long long)s_len[(dk & 0xFFFFu) + cx * ((dk >> 16) & 255u) + cy] << (16 * r);
      }
    }
  }
  __syncwarp();
#pragma unroll U
  for (int j = 0; j < 9; ++j) {
    const int p = lane + 32 * j;
    if (p >= ch.bv && p < ch.bv + 2 * ch.c1 && !((p - ch.bv) & 1)) {
      int qx, qy; q(j, qx, qy);
      const int idx = ((qx & 1) << 3) | ((qy & 1) << 2) | s_c[p + 1];
      const int sg = __popc(idx);
      misc += (uint32_t)(quad_len_a(idx) + sg) << 10 | (uint32_t)(4 + sg) << 21;
    }
  }
  __syncwarp();
  uint32_t lo[3], hi[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { lo[k] = (uint32_t)warp_sum_i((int)(uint32_t)acc[k]); hi[k] = (uint32_t)warp_sum_i((int)(uint32_t)(acc[k] >> 32)); }
  n15 = (uint32_t)warp_sum_i((int)n15); misc = (uint32_t)warp_sum_i((int)misc);
  int bits = (int)(misc & 1023u);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    int best = 1 << 30, sel = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (cand[r][k] < 0) continue;
      const int sum = r == 0 ? (int)(lo[k] & 0xFFFFu) : r == 1 ? (int)(lo[k] >> 16) : (int)(hi[k] & 0xFFFFu);
      const int tot = sum + (int)((d[r][k] >> 24) & 255u) * (int)((n15 >> (10 * r)) & 1023u);
      if (tot < best) { best = tot; sel = cand[r][k]; }
    }
    ch.tsel[r] = sel; bits += best;
  }
  const int ca = (int)((misc >> 10) & 2047u), cb = (int)(misc >> 21);
  ch.c1sel = cb < ca;
  ch.bits = bits + min(ca, cb);
  return ch;
}
// ONE copy of the evaluation in the library, called from every search site: inlined into each of them (bisections, galloping
// searches, the curve — nine sites in k_outer), the kernels were 350 KB of straight-line code and their warps spent most of the
// time between two instructions waiting for instruction fetch (ncu: 35 of 40 cycles).
__device__ __noinline__ IsoChoice iso_evaluate(const int qx[9], const int qy[9], int lane, const uint8_t *s_len, uint8_t *s_c, const int *sfb, bool ws = false) {
  return iso_evaluate_core<9>([&](int j, int &x, int &y) { x = qx[j]; y = qy[j]; }, lane, s_len, s_c, sfb, ws);
}
// The search form: quantize the warp's 288 pairs of (amplified) magnitudes m[lane + 32 j] at gain G and count.  Returns
// min(bits, 65535) | big_values << 16 — one register; the magnitudes and the quantized values (sq: 288 words of scratch per warp,
// a lane reads back only what it wrote) stay in shared memory.
// U: unrolling of the pair loops — 3 for k_outer (~40 calls per granule-channel: the code has to stay in the instruction cache), 9 for
// k_granule's level-1 search (fewer calls, more resident warps: the unrolled form measured 6 % faster there).
template <int U> __device__ __noinline__ uint32_t iso_eval_gain(int G, const float2 *m, uint32_t *sq, int lane, const uint8_t *s_len, uint8_t *s_c, const int *sfb, bool ws) {
  const float inv = c_inv_step_iso[G];
#pragma unroll U
  for (int j = 0; j < 9; ++j) { const float2 v = m[lane + 32 * j]; sq[lane + 32 * j] = (uint32_t)iso_quant(v.x, inv) | (uint32_t)iso_quant(v.y, inv) << 16; }
  const IsoChoice c = iso_evaluate_core<U>([&](int j, int &x, int &y) { const uint32_t w = sq[lane + 32 * j]; x = (int)(w & 0xFFFFu); y = (int)(w >> 16); }, lane, s_len, s_c, sfb, ws);
  return (uint32_t)min(c.bits, 65535) | (uint32_t)c.bv << 16;
}


// ---- level 3: window switching ----------------------------------------------------------------------------------------------------
// ISO 11172-3 block types: 0 normal, 1 start, 2 short (three 12-point windows), 3 stop — the windows the reference defines and never
// uses (SRC:1470-1503; it switches with plain sine windows).  A short granule needs a start window in the granule
// BEFORE it, i.e. one granule of look-ahead.  The look-ahead costs no engine change: at level 3 the filterbank (and the
// psychoacoustic window) read the PCM 576 samples late — inside the carried frame every pass already holds — while the transient
// detector (the reference's, SRC:1944-1968: thirds of 192 samples, max / min > 6) looks at the newest granule, which for the
// delayed signal IS the next granule.  With a(j) = attack in PCM granule j of either channel, delayed granule g (current = PCM
// granule g - 1):  short if a(g-1);  else start if a(g), short instead if a(g-2) too (a granule cannot be stop and start at once);
// else stop if a(g-2);  else normal.  Pure function of the PCM in [carried frame, this pass], so chunking cannot change it.
__constant__ float c_iso_win[4][36];               // MDCT windows by block type (type 2 unused: kWinShort)
__constant__ uint16_t c_short_pos[3][192];         // short blocks: line of (frequency 6 sb + m, window 0) in scalefactor-band order
__constant__ uint8_t c_short_width[3][192];        // ... + window * width of its short scalefactor band

cudaError_t upload_iso_switch_tables() {
  float win[4][36];
  const double pi = 3.14159265358979323846;
  for (int i = 0; i < 36; ++i) {
    const double l = sin(pi / 36.0 * (i + 0.5));
    win[0][i] = (float)l; win[2][i] = 0.0f;
    win[1][i] = (float)(i < 18 ? l : i < 24 ? 1.0 : i < 30 ? sin(pi / 12.0 * (i - 18 + 0.5)) : 0.0);
    win[3][i] = (float)(i < 6 ? 0.0 : i < 12 ? sin(pi / 12.0 * (i - 6 + 0.5)) : i < 18 ? 1.0 : l);
  }
  static const int sfs[3][14] = {{0, 4, 8, 12, 16, 22, 30, 40, 52, 66, 84, 106, 136, 192},       // 44.1 kHz (ISO 11172-3 Table B.8)
                                 {0, 4, 8, 12, 16, 22, 28, 38, 50, 64, 80, 100, 126, 192},       // 48 kHz
                                 {0, 4, 8, 12, 16, 22, 30, 42, 58, 78, 104, 138, 180, 192}};     // 32 kHz
  uint16_t pos[3][192]; uint8_t wid[3][192];
  for (int r = 0; r < 3; ++r)
    for (int b = 0; b < 13; ++b)
      for (int f = sfs[r][b]; f < sfs[r][b + 1]; ++f) { pos[r][f] = (uint16_t)(3 * sfs[r][b] + (f - sfs[r][b])); wid[r][f] = (uint8_t)(sfs[r][b + 1] - sfs[r][b]); }
  cudaError_t e;
  if ((e = cudaMemcpyToSymbol(c_iso_win, win, sizeof win))) return e;
  if ((e = cudaMemcpyToSymbol(c_short_pos, pos, sizeof pos))) return e;
  return cudaMemcpyToSymbol(c_short_width, wid, sizeof wid);
}

}  // namespace mp3b
