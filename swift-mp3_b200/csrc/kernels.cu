// sm_100a kernels of the MP3 encode path.  SRC = Sources/SwiftMP3/MP3Encoder.swift of the reference.
//
// Numerical contract (bit-exact against oracle/mp3_oracle.c): every floating-point operation below is an explicit
// IEEE-754 round-to-nearest intrinsic (__fmul_rn / __fadd_rn / __fmaf_rn / __fdiv_rn) in the order the oracle
// defines ([OD1]..[OD5] in its header); the file is compiled with -fmad=false so nothing else is contracted.
// Constant tables are literals (tables_gen.h): the MDCT matrices are instruction immediates of fully unrolled loops,
// the 32x64 analysis matrix and the 512-tap window are staged once per CTA in shared memory for the packed FP32x2 loop.
#include "kernels.h"

#include <cstdio>
#include <cstdlib>

#include "tables_gen.h"

namespace mp3b {

__constant__ float c_inv_step[256];     // 1 / Float(max(2^((g-210)/4), 1e-4)), SRC:798-800
__constant__ double c_gain_thr[256];    // 2^((g-210)/4) in double: replaces log2 in computeGlobalGain, SRC:1004
__constant__ int c_sfb_cum[3][21];      // cumulative long sfb widths, SRC:1814-1820

extern const float *host_inv_step();    // tables.cc
extern const double *host_gain_thr();
extern const int *host_sfb_cum();

extern const float *host_inv_step_iso();
cudaError_t upload_iso_tables(const float *inv_step_iso);   // iso_mode.cuh
cudaError_t upload_psy_constants();                         // iso_psy.cuh
cudaError_t upload_iso_switch_tables();                     // iso_mode.cuh

cudaError_t upload_tables() {
  cudaError_t e;
  if ((e = upload_iso_tables(host_inv_step_iso()))) return e;
  if ((e = upload_psy_constants())) return e;
  if ((e = upload_iso_switch_tables())) return e;
  if ((e = cudaMemcpyToSymbol(c_inv_step, host_inv_step(), sizeof(float) * 256))) return e;
  if ((e = cudaMemcpyToSymbol(c_gain_thr, host_gain_thr(), sizeof(double) * 256))) return e;
  if ((e = cudaMemcpyToSymbol(c_sfb_cum, host_sfb_cum(), sizeof(int) * 63))) return e;
  return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------------------
// helpers

// Logical PCM sequence of a stream in one pass: head[0, head_n) ++ cur[0, cur_n) ++ zeros (flush padding, SRC:322-328).
struct PcmView {
  const float *head, *cur;
  uint32_t head_n, cur_n;
  __device__ __forceinline__ float at(int64_t q) const {
    if (q < (int64_t)head_n) return head[q];
    q -= head_n;
    return q < (int64_t)cur_n ? __ldg(cur + q) : 0.0f;
  }
};
__device__ __forceinline__ PcmView pcm_view(const Config &cfg, const PassBuffers &pb, int s) {
  const StreamPlan &p = pb.plan[s];
  PcmView v;
  v.head = pb.head_in + (size_t)s * 2 * cfg.fsc;
  v.cur = p.cur; v.head_n = p.head_n; v.cur_n = p.cur_n;
  return v;
}

// [OD1b] butterfly tree over the 32 lane partials; every lane ends with the same bits (a + b == b + a).
__device__ __forceinline__ float lane_tree(float p) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) p = __fadd_rn(p, __shfl_xor_sync(0xffffffffu, p, m));
  return p;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, m));
  return v;
}
// lane_tree of two values for the price of one and a half: the upper half-warp reduces b while the lower one reduces a (the
// same additions as the butterfly performs in those lanes), then both results are broadcast.
__device__ __forceinline__ void lane_tree_pair(float &a, float &b, int lane) {
  const bool up = lane & 16;
  float keep = up ? b : a;
  keep = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, up ? a : b, 16));
#pragma unroll
  for (int m = 8; m >= 1; m >>= 1) keep = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, keep, m));
  a = __shfl_sync(0xffffffffu, keep, 0); b = __shfl_sync(0xffffffffu, keep, 16);
}
// maximum of non-negative floats: their bit patterns order like unsigned integers (one REDUX instead of five shuffles and maxima)
__device__ __forceinline__ float warp_max_nonneg(float v) { return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(v))); }
__device__ __forceinline__ int warp_max_i(int v) { return __reduce_max_sync(0xffffffffu, v); }   // REDUX: one instruction
__device__ __forceinline__ int warp_sum_i(int v) { return __reduce_add_sync(0xffffffffu, v); }

// d^-1/2 in double for arguments in the normal range (no zero / subnormal / huge inputs, which is all this file ever feeds
// it): reciprocal-square-root seed and one third-order Newton step.  The library's sqrt.rn.f64 continues with d * y, the
// correction g + (d - g * g) * (y / 2), a range guard and a slow-path call to deliver the correctly rounded root; pow34
// needs none of that, because the last-bit error never survives the rounding of the result to FP32: mp3b_selftest
// compares pow34 with the IEEE definition on EVERY finite float >= 1e-10 (0 of 2.3e9 differ; with a second-order step
// instead it counts 5052, so the test has teeth).
__device__ __forceinline__ double drsqrt_fast(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = __fma_rn(d, -__dmul_rn(y, y), 1.0);             // 1 - d y^2
  const double p = __fma_rn(e, 0.375, 0.5);
  return __fma_rn(p, __dmul_rn(y, e), y);                          // y (1 + e / 2 + 3 e^2 / 8)
}

// [OD3] |x|^0.75 = (float)(sqrt(d) * sqrt(sqrt(d))) in IEEE double, a >= 1e-10f; evaluated as d * (d^1/2)^-1/2.
// exact widening of a POSITIVE NORMAL float with integer operations: the conversion instruction (F2F) runs on the XU pipe, one
// warp instruction per 8 cycles per scheduler, which MUFU.RSQ64H and the quantizer's F2I already keep busy
__device__ __forceinline__ double widen_normal(float a) {
  const uint32_t b = __float_as_uint(a);
  return __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29));
}
template <bool NORMAL = true> __device__ __forceinline__ float pow34(float a) {
  const double d = NORMAL ? widen_normal(a) : (double)a;
#if MP3B_POW34_V1
  const double r = __dmul_rn(d, drsqrt_fast(d));                   // d^1/2
  return __double2float_rn(__dmul_rn(d, drsqrt_fast(r)));          // d * d^-1/4
#else
  // d * d^-1/4 with ONE Newton step: seed q0 = y * y^-1/2 from two reciprocal-square-root approximations (y ~ d^-1/2, so
  // q0 ~ d^-1/4 to about 2^-21), then the third-order step for the inverse fourth root, e = 1 - d q0^4,
  // q = q0 (1 + e / 4 + 5 e^2 / 32): 8 FP64 instructions instead of 12 — the same exhaustive proof (mp3b_selftest)
  double y, z;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(z) : "d"(y));
  const double q0 = __dmul_rn(y, z);
  const double q2 = __dmul_rn(q0, q0);
  const double e = __fma_rn(-d, __dmul_rn(q2, q2), 1.0);
  const double p = __fma_rn(e, 0.15625, 0.25);
  const double q = __fma_rn(p, __dmul_rn(q0, e), q0);
  return __double2float_rn(__dmul_rn(d, q));
#endif
}
__device__ __forceinline__ float pow34_reference(float a) {      // the same with the library's IEEE square root
  double d = (double)a;
  double r = __dsqrt_rn(d);
  return __double2float_rn(__dmul_rn(r, __dsqrt_rn(r)));
}

// quantizeWithGain SRC:816-821: q = min(Int32(roundf(mag * inv)), 15); roundf = ties away from zero [OD4].
// inv2 = 2 * inv: scaling by two commutes with the rounding of the product (no overflow / underflow at these
// magnitudes), so floor(RN(mag * inv2)) = floor(2 t) with t = RN(mag * inv), and roundf(t) = (floor(2 t) + 1) >> 1, t >= 0.
// The kernels keep the quantizer as an index: u = min(floor(2 t), 30), and q = (u + 1) >> 1 (u = 29, 30 -> 15).  The
// Huffman tables indexed by (ux, uy) (tab::kLen31s, tab::kTab31, row stride 32) save the increment, shift and clamp per value;
// q != 0 <=> u != 0.
__device__ __forceinline__ int quant30(float mag, float inv2) { return min(__float2int_rd(__fmul_rn(mag, inv2)), 30); }
// the same index without the conversion instruction (XU pipe): 2^23 + floor(min(t, 30.5)) by an addition that rounds down; the
// index of a pair is ((mx << 5) + my) & 1023 on the bit patterns (2^23 as a float is 0x4B000000: nothing of it survives the mask)
__device__ __forceinline__ uint32_t quant30m(float mag, float inv2) { return __float_as_uint(__fadd_rd(fminf(__fmul_rn(mag, inv2), 30.5f), 8388608.0f)); }
__device__ __forceinline__ int pair_index(float mx, float my, float inv2) { return (int)(((quant30m(mx, inv2) << 5) + quant30m(my, inv2)) & 1023u); }
// ... with both products in one packed multiply (the same IEEE product per half)
__device__ __forceinline__ int pair_index2(float2 m, float inv2) {
  const float2 t = __fmul2_rn(m, make_float2(inv2, inv2));
  const uint32_t ux = __float_as_uint(__fadd_rd(fminf(t.x, 30.5f), 8388608.0f)), uy = __float_as_uint(__fadd_rd(fminf(t.y, 30.5f), 8388608.0f));
  return (int)(((ux << 5) + uy) & 1023u);
}

// a / d correctly rounded for d = 9 and d = 3 (r = RN(1 / d)): quotient estimate, exact residual, one correction.
// tools/check_div.c compares it with the IEEE division for all 2^32 floats (signed zeros and denormals included).
__device__ __forceinline__ float div_exact(float a, float d, float r) {
  const float q = __fmul_rn(a, r);
  const float e = __fmaf_rn(d, q, -a);
  return __fmaf_rn(-e, r, q);
}

}  // namespace mp3b
#include "iso_mode.cuh"   // opt-in ISO mode: quantizer, partition, table selection (needs the warp helpers above)
namespace mp3b {

// x / 192 as (x / 3) / 64: the division by 64 is exact unless the quotient is subnormal (then the IEEE division)
__device__ __forceinline__ float div192(float x) {
  return x >= 1e-30f ? __fmul_rn(div_exact(x, 3.0f, 1.0f / 3.0f), 0.015625f) : __fdiv_rn(x, 192.0f);
}

// ------------------------------------------------------------------------------------------------------------
// K0: pre-pass.  One warp per frame: frame energy (SRC:477), stereo decision (SRC:2140-2162), granule energies
// (SRC:673), transient thirds -> block type + subblock_gain (SRC:1944-1968).

// ALL_SBG = false: subblock_gain only where it is coded (window-switched granules, SRC:586-624); long granules get 0
template <bool ALL_SBG = true> __device__ __forceinline__ void transient_decide(const float e3[3], int &bt, int sbg[3]) {
  float mx = fmaxf(e3[0], fmaxf(e3[1], e3[2])), mn = fminf(e3[0], fminf(e3[1], e3[2]));
  float ratio = __fdiv_rn(mx, fmaxf(mn, 0.0001f));
  if (ratio > 6.0f) bt = (e3[0] == mx) ? 1 : 2; else bt = 0;
  if (!ALL_SBG && bt == 0) { sbg[0] = sbg[1] = sbg[2] = 0; return; }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float normalized = fminf(fmaxf(__fdiv_rn(e3[i], fmaxf(mx, 0.0001f)), 0.0f), 1.0f);
    sbg[i] = (int)__fmul_rn(__fsub_rn(1.0f, normalized), 7.0f);
  }
}

// Frame access for the pre-pass: MODE 0 = generic (frame straddles head / cur / zero padding), 1 = contiguous in cur,
// 2 = contiguous and 8-byte aligned (stereo pairs as one 64-bit load).
template <int MODE> struct FrameLoad {
  PcmView pv; int64_t q0; const float *p;
  __device__ __forceinline__ float one(int i) const { return MODE ? __ldg(p + i) : pv.at(q0 + i); }
  __device__ __forceinline__ float2 pair(int n) const {
    if (MODE == 2) return __ldg(reinterpret_cast<const float2 *>(p) + n);
    return make_float2(one(2 * n), one(2 * n + 1));
  }
};

// JOINT / CH are compile-time so that the mid / side variants (and the second channel of mono) cost nothing when unused.
template <int MODE, bool JOINT, int CH> __device__ __forceinline__ void prepass_frame(const Config &cfg, const PassBuffers &pb, const FrameLoad<MODE> &ld,
                                                                  int s, int f, int lane) {
  constexpr int ch = CH;
  constexpr int NV = JOINT ? 4 : CH;               // signal variants: L, R (or mono), mid, side
  // Frame energy over the interleaved frame (SRC:477) [OD1b]: partial p sums the floats at interleaved positions = p mod 32,
  // ascending.  Mono: lane = partial.  Stereo: position 2 n + c, so partial 2 q + c walks n = q, q + 16, q + 32, ... — values
  // that lanes q and q + 16 of the pair loop below hold alternately.  Lane q (< 16) therefore accumulates partial 2 q from its
  // own left sample and lane q + 16's, lane q + 16 accumulates partial 2 q + 1 from lane q's right sample and its own: one
  // shuffle per pair instead of a second pass over the frame.
  float pf = 0.0f;

  // per-channel signals; variant 0/1 = L/R (or mono), 2/3 = mid/side
  constexpr bool joint = JOINT;
  float eg[4][2], e3[4][2][3];
  float pm = 0.0f, ps = 0.0f;
#pragma unroll
  for (int gr = 0; gr < 2; ++gr) {
    float ag[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int th = 0; th < 3; ++th) {
      float a3[4] = {0.f, 0.f, 0.f, 0.f};
      float2 lr[6];
#pragma unroll
      for (int jj = 0; jj < 6; ++jj) {
        int n = gr * 576 + th * 192 + jj * 32 + lane;
        lr[jj] = ch == 1 ? make_float2(ld.one(n), 0.0f) : ld.pair(n);
      }
#pragma unroll
      for (int jj = 0; jj < 6; ++jj) {
        float v[4] = {lr[jj].x, lr[jj].y, 0.0f, 0.0f};
        if (ch == 1) pf = __fmaf_rn(v[0], v[0], pf);
        else {
          const float got = __shfl_xor_sync(0xffffffffu, lane < 16 ? v[1] : v[0], 16);
          const float first = lane < 16 ? v[0] : got, second = lane < 16 ? got : v[1];
          pf = __fmaf_rn(first, first, pf); pf = __fmaf_rn(second, second, pf);
        }
        if (joint) {
          v[2] = __fmul_rn(__fadd_rn(v[0], v[1]), 0.5f);       // SRC:2148-2150
          v[3] = __fmul_rn(__fsub_rn(v[0], v[1]), 0.5f);       // SRC:2153-2154
          pm = __fmaf_rn(v[2], v[2], pm); ps = __fmaf_rn(v[3], v[3], ps);
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) { a3[k] = __fmaf_rn(v[k], v[k], a3[k]); ag[k] = __fmaf_rn(v[k], v[k], ag[k]); }
      }
#pragma unroll
      for (int k = 0; k < NV; ++k) e3[k][gr][th] = __fdiv_rn(lane_tree(a3[k]), 192.0f);
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) eg[k][gr] = __fdiv_rn(lane_tree(ag[k]), 576.0f);
  }
  // butterfly tree over the partials in partial-index order 16, 8, 4, 2, 1; stereo lane L holds partial 2 (L & 15) + (L >> 4),
  // so those are lane distances 8, 4, 2, 1, 16
  if (ch == 2) {
#pragma unroll
    for (int m = 8; m >= 1; m >>= 1) pf = __fadd_rn(pf, __shfl_xor_sync(0xffffffffu, pf, m));
    pf = __fadd_rn(pf, __shfl_xor_sync(0xffffffffu, pf, 16));
  } else pf = lane_tree(pf);
  const float frame_energy = __fdiv_rn(pf, (float)cfg.fsc);
  int ms = 0;
  if (joint) {
    float me = __fdiv_rn(lane_tree(pm), 1152.0f), se = __fdiv_rn(lane_tree(ps), 1152.0f);
    ms = se < __fmul_rn(me, 0.4f) ? 1 : 0;                      // SRC:2156-2158
  }
  if (lane == 0) {
    pb.frame_energy[(size_t)s * pb.Fc + f] = frame_energy;
    pb.ms[(size_t)s * (pb.Fc + 1) + 1 + f] = (uint8_t)ms;
    if (!cfg.vbr) pb.frame_br[(size_t)s * pb.Fc + f] = (uint8_t)cfg.cbr_index;
  }
  if (lane < 2 * ch) {
    int gr = lane / ch, c = lane % ch, k = c + (ms ? 2 : 0);
    float e[3] = {0.f, 0.f, 0.f}; float g = 0.0f;
#pragma unroll
    for (int kk = 0; kk < NV; ++kk)
#pragma unroll
      for (int g2 = 0; g2 < 2; ++g2)
        if (kk == k && g2 == gr) { e[0] = e3[kk][g2][0]; e[1] = e3[kk][g2][1]; e[2] = e3[kk][g2][2]; g = eg[kk][g2]; }
    int bt, sbg[3];
    transient_decide(e, bt, sbg);
    int gci = (2 * f + gr) * ch + c;
    pb.gc_energy[(size_t)s * (10 + pb.GC) + 10 + gci] = g;
    pb.gc_bt[(size_t)s * pb.GC + gci] = (uint16_t)(bt | sbg[0] << 2 | sbg[1] << 5 | sbg[2] << 8);
  }
}

__global__ void __launch_bounds__(128) k_prepass(Config cfg, PassBuffers pb) {
  const int s = blockIdx.x, lane = threadIdx.x & 31;
  const int f = blockIdx.y * 4 + (threadIdx.x >> 5);
  const StreamPlan &plan = pb.plan[s];
  if (f >= (int)plan.n_frames) return;
  const PcmView pv = pcm_view(cfg, pb, s);
  const int64_t q0 = (int64_t)(1 + f) * cfg.fsc;
  if (f == 0 && lane < 10) {  // carried VBR history, right aligned in front of this pass's gc energies
    const StreamState &stt = pb.state[s];
    int n = stt.vbr_n;
    pb.gc_energy[(size_t)s * (10 + pb.GC) + lane] = lane >= 10 - n ? stt.vbr_hist[lane - (10 - n)] : 0.0f;
  }
  const int64_t rel = q0 - (int64_t)pv.head_n;
  int mode = 0; const float *p = nullptr;
  if (rel >= 0 && rel + cfg.fsc <= (int64_t)pv.cur_n) { p = pv.cur + rel; mode = (reinterpret_cast<uintptr_t>(p) & 7) == 0 ? 2 : 1; }
#define MP3B_PREPASS(M, J, C) { FrameLoad<M> ld{pv, q0, p}; prepass_frame<M, J, C>(cfg, pb, ld, s, f, lane); }
#define MP3B_PREPASS_M(J, C) { if (mode == 2) MP3B_PREPASS(2, J, C) else if (mode == 1) MP3B_PREPASS(1, J, C) else MP3B_PREPASS(0, J, C) }
  if (cfg.channels == 1) MP3B_PREPASS_M(false, 1)
  else if (cfg.mode == 2) MP3B_PREPASS_M(true, 2)
  else MP3B_PREPASS_M(false, 2)
#undef MP3B_PREPASS_M
#undef MP3B_PREPASS
}

// VBRState.chooseBitrate SRC:1177-1189 + MP3Tables.bitrateIndex SRC:2509-2523: one thread per frame.
__global__ void k_bitrate(Config cfg, PassBuffers pb) {
  const int s = blockIdx.x, f = blockIdx.y * blockDim.x + threadIdx.x;
  if (f >= (int)pb.plan[s].n_frames) return;
  const float *hist = pb.gc_energy + (size_t)s * (10 + pb.GC);
  const int before = f * 2 * cfg.channels;             // gc energies of this pass appended before frame f
  int count = pb.state[s].vbr_n + before; if (count > 10) count = 10;
  const float e = pb.frame_energy[(size_t)s * pb.Fc + f];
  float average;
  if (count == 0) average = e;
  else {
    float sum = 0.0f;
    for (int i = 10 + before - count; i < 10 + before; ++i) sum = __fadd_rn(sum, hist[i]);
    average = __fdiv_rn(sum, (float)count);
  }
  float ratio = fminf(fmaxf(__fdiv_rn(e, fmaxf(average, 0.0001f)), 0.5f), 2.0f);
  float quality_factor = __fdiv_rn((float)(9 - cfg.quality), 9.0f);
  int max_adjustment = (int)__fadd_rn(32.0f, __fmul_rn(32.0f, quality_factor));
  int adjustment = (int)__fmul_rn(__fsub_rn(ratio, 1.0f), (float)max_adjustment);
  int lo = max(32, cfg.base_kbps - 64 + cfg.quality * 8), hi = min(320, cfg.base_kbps + 64 - cfg.quality * 4);
  int kbps = max(lo, min(cfg.base_kbps + adjustment, hi));
  kbps = min(max(kbps, 0), 320);
  pb.frame_br[(size_t)s * pb.Fc + f] = cfg.vbr_idx_of_kbps[kbps];
}

// ------------------------------------------------------------------------------------------------------------
constexpr int kLook = 15;                         // 480 samples of look-back = 15 rows of 32

__device__ __forceinline__ int gain_from_peak(float peak) {      // computeGlobalGain SRC:989-1006
  if (!(peak > 0.0f)) return 210;
  float ratio = __fdiv_rn(pow34<false>(peak), 15.0f);
  if (ratio <= 0.0f) return 210;
  double r = (double)ratio;
  // 210 + Int(4*log2(r)): Int() truncates toward zero.  i = largest index with 2^((i-210)/4) <= r.
  if (r < c_gain_thr[0]) return 0;
  // start from the hardware logarithm (any estimate would do) and walk to the largest index whose threshold is <= r: zero or
  // one step instead of the eight of a bisection
  int lo = min(max(__float2int_rd(__fmaf_rn(__log2f(ratio), 4.0f, 210.0f)), 0), 255);
  while (lo < 255 && c_gain_thr[lo + 1] <= r) ++lo;
  while (lo > 0 && c_gain_thr[lo] > r) --lo;
  int gain = lo;                                                   // floor
  if (r < 1.0 && c_gain_thr[lo] != r) gain = lo + 1;               // negative values truncate upwards
  return gain > 255 ? 255 : gain;
}

// ------------------------------------------------------------------------------------------------------------
// K1: polyphase analysis filterbank (SRC:917-944, 1367-1411), PCM -> subband samples sub[s][c][step][32] in HBM.
// Same arithmetic as the reference's direct form, operation for operation, organised so that the FP32 pipe and not the
// shared-memory pipe is the limiter (tools/microbench/mb.cu holds the measurements the shapes below come from):
//   * a CTA (4 warps) owns a run of R granules of one (stream, channel) and walks it in tiles of 256 filterbank steps;
//   * windowing: a thread owns one window phase n and slides along the steps two at a time, so every PCM sample is
//     read from shared memory once per (n, parity) instead of once per tap; the pair (Y[n][u], Y[n][u+1]) is one FFMA2;
//   * matrixing: a register-tiled 32 x 256 x 64 product, thread = 16 subbands x 4 steps (64 accumulators): five
//     LDS.128 (twelve shared-memory wavefronts) feed 32 FFMA2; the reduction over n stays ascending, one fused
//     multiply-add per term;
//   * n is processed in two halves (window 0..31, matrix 0..31, window 32..63, matrix 32..63) so that Y is 32 KB and
//     three CTAs fit an SM;
//   * the PCM rows of tile j+1 are fetched with cp.async while tile j is in its second matrixing half.
// The MDCT lives in k_granule: it is warp-local work that wants many resident warps, this kernel wants shared memory.
// compile-time loop: f(integral_constant<int, 0>) ... f(integral_constant<int, N - 1>)
template <int V> struct IntC { static constexpr int value = V; };
template <int I, int N, class F> __device__ __forceinline__ void static_for(F &&f) {
  if constexpr (I < N) { f(IntC<I>{}); static_for<I + 1, N>(f); }
}
// one 4-byte cp.async with both offsets as instruction immediates
template <int DST_OFF, int SRC_OFF> __device__ __forceinline__ void cp_async4_imm(uint32_t dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0+%2], [%1+%3], 4;" ::"r"(dst), "l"(src), "n"(DST_OFF), "n"(SRC_OFF));
}
constexpr int kTile = 256;                        // filterbank steps per tile
constexpr int kPRows = kTile + kLook;             // 271 PCM rows of 32 samples: 15 rows of look-back + 256 new
constexpr int kFbThreads = 128;
constexpr int kFbWarps = kFbThreads / 32;
constexpr int kFbSmemFloats = 64 * 32 + kPRows * 32 + 32 * kTile;
constexpr int kFbSmemBytes = kFbSmemFloats * 4;

__global__ void __launch_bounds__(kFbThreads, 3) k_filterbank(Config cfg, PassBuffers pb, int Rdbg) {
  const int R = Rdbg & 0xFFFF, dbg = Rdbg >> 16;     // dbg: timing experiments only (tools/stage_times.py), 0 in production
  extern __shared__ __align__(16) float sm[];
  float *sM = sm;                                  // [64 n][32 k] analysis matrix, transposed
  float *P = sM + 64 * 32;                         // [271][32] PCM rows of the tile
  float *Y = P + kPRows * 32;                      // [32 n][256 t] of the current n half, 16-byte chunks XOR-swizzled with n & 7
  const int c = blockIdx.x, s = blockIdx.y, run = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const StreamPlan &plan = pb.plan[s];
  const int ngr = 2 * (int)plan.n_frames;
  const int g_begin = run * R;
  if (g_begin >= ngr) return;
  const int g_cnt = min(R, ngr - g_begin);
  const int U = 18 * g_cnt;                        // steps of the run
  const int n_tiles = (U + kTile - 1) / kTile;
  const int rows_total = kLook + U;                // run row r holds samples n_start + 32 r ... + 31; step u reads rows u ... u + 15
  const int ch = cfg.channels;
  const PcmView pv = pcm_view(cfg, pb, s);
  const uint8_t *msrow = pb.ms + (size_t)s * (pb.Fc + 1);
  const uint32_t ms_prev = pb.state[s].ms_prev;
  const bool joint = cfg.mode == 2;
  const int n_start = 576 * g_begin - 480 - cfg.iso_delay;   // (ISO mode level 3 reads the PCM one granule late: iso_mode.cuh)
  float *out = pb.sub + ((size_t)(s * ch + c) * pb.sub_rows + 18 * (g_begin + 1)) * 32;

  // PCM rows [ra, rb) of the run -> P rows slot0 ...  Fast path: the rows are contiguous in this pass's PCM and need no
  // mid/side transform: one 4-byte cp.async per sample, nothing waits until the next tile starts.
  auto load_rows = [&](int ra, int rb, int slot0) {
    rb = (dbg & 4) ? ra : min(rb, rows_total);
    const int ra0 = ra;
    const int lane_off = ch == 1 ? lane : 2 * lane + c;
    const int64_t rel0 = (int64_t)(n_start + 32 * ra + 1152) * ch - (int64_t)pv.head_n;
    if (!joint && ra < rb && rel0 >= 0 && rel0 + (int64_t)(rb - ra) * 32 * ch <= (int64_t)pv.cur_n) {
      // the whole range is inside this pass's PCM (the common case): no per-row address arithmetic
      const float *src = pv.cur + rel0 + lane_off + (size_t)warp * 32 * ch;
      uint32_t dst = (uint32_t)__cvta_generic_to_shared(P + (slot0 + warp) * 32 + lane);
      const int step = kFbWarps * 32 * ch;
      for (int r = warp; r < rb - ra; r += kFbWarps, src += step, dst += kFbWarps * 128)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src));
      ra = rb;
    }
    // otherwise row by row: a row (32 ch consecutive floats) that lies entirely in this pass's PCM or entirely in the
    // carried head and needs no mid/side transform is still a cp.async; only straddling / zero-padded / mid-side rows are
    // loaded synchronously
    for (int r = ra + warp; r < rb; r += kFbWarps) {
      const int nrow = n_start + 32 * r;
      const int64_t q = (int64_t)(nrow + 1152) * ch, rel = q - (int64_t)pv.head_n;
      float *dstp = P + (slot0 + r - ra0) * 32 + lane;
      const float *src = nullptr;
      if (!joint) {
        if (rel >= 0 && rel + 32 * ch <= (int64_t)pv.cur_n) src = pv.cur + rel + lane_off;
        else if (q >= 0 && q + 32 * ch <= (int64_t)pv.head_n) src = pv.head + q + lane_off;
      }
      if (src) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dstp)), "l"(src));
      } else {
        float v;
        if (ch == 1) v = pv.at(q + lane);
        else {
          const float l = pv.at(q + 2 * lane), rr = pv.at(q + 2 * lane + 1);
          const int fr = nrow >= 0 ? nrow / 1152 : -1;
          const bool ms = joint && (fr < 0 ? ms_prev != 0 : msrow[1 + fr] != 0);
          if (!ms) v = c == 0 ? l : rr;
          else v = c == 0 ? __fmul_rn(__fadd_rn(l, rr), cfg.ms_scale) : __fmul_rn(__fsub_rn(l, rr), cfg.ms_scale);   // SRC:2148-2154 (0.5; ISO mode: 1 / sqrt 2)
        }
        *dstp = v;
      }
    }
    asm volatile("cp.async.commit_group;");
  };
  load_rows(0, kPRows, 0);
  {                                                // analysis matrix -> shared memory, asynchronously as well
    const uint32_t dstm = (uint32_t)__cvta_generic_to_shared(sM);
    for (int e = tid; e < 64 * 32 / 4; e += kFbThreads)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dstm + e * 16), "l"(tab::kAnalysisT + e * 4));
    asm volatile("cp.async.commit_group;");
  }

  // windowing role: n = 32 H + lane in half H, steps 64 warp ... 64 warp + 63 of the tile
  float wc[2][8];                                  // C[32 H + lane + 64 i]
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int i = 0; i < 8; ++i) wc[h][i] = __ldg(tab::kWindow + 32 * h + lane + 64 * i);
  const float2 neg0 = make_float2(cfg.f_neg0, cfg.f_neg0), one = make_float2(cfg.f_one, cfg.f_one);
  // matrixing role: subbands 16 kg ... 16 kg + 15, steps 64 warp + 4 tg + j of the tile
  const int tg = lane & 15, kg = lane >> 4;
  int yo[8];                                       // XOR-swizzled chunk of this thread's four steps in row n, n & 7 = b
#pragma unroll
  for (int b = 0; b < 8; ++b) yo[b] = b * kTile + (((16 * warp + tg) ^ b) << 2);

  const float *pf_src = nullptr; uint32_t pf_dst = 0; int pf_left = 0;   // cp.async of the next tile still to be issued by this warp
  const int pf_step = kFbWarps * 32 * ch;
  for (int tile = 0; tile < n_tiles; ++tile) {
    float2 acc[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[j][i] = make_float2(0.0f, 0.0f);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    const int valid = min(kTile, U - kTile * tile);
#pragma unroll
    for (int H = 0; H < 2; ++H) {
      // A warp windows and matrixes the same 64 steps, so its part of Y is private: warp-level ordering is enough
      // between the phases.  Block-wide barriers only guard P: its arrival here, its reuse after the last windowing.
      if (H == 0) __syncthreads(); else __syncwarp();
      // ---- windowing (SRC:1386-1399): X[n + 64 i] of step u = sample at tile row u + 15 - 2 i - H, column 31 - lane
      if (64 * warp < valid && !(dbg & 1)) {                     // warp-uniform: a short last tile skips the steps beyond the run
        const float *Pc = P + (64 * warp + 1 - H) * 32 + (31 - lane);   // oldest row of the first step pair
        float2 q[8];
#pragma unroll
        for (int k = 1; k <= 7; ++k) { q[k].x = Pc[(2 * k - 2) * 32]; q[k].y = Pc[(2 * k - 1) * 32]; }
        Pc += 14 * 32;
        float *yrow = Y + lane * kTile;
#pragma unroll 1
        for (int o = 0; o < 4; ++o, Pc += 16 * 32) {
          float2 ypair[8];
#pragma unroll
          for (int ii = 0; ii < 8; ++ii) {
            q[ii].x = Pc[(2 * ii) * 32]; q[ii].y = Pc[(2 * ii + 1) * 32];
            float2 y;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              // fma(x, w, -0) == RN(x * w) and fma(y, 1, z) == RN(y + z): two roundings, as the reference's vDSP_vmul +
              // vDSP_sve; the constants are run-time values so that ptxas cannot contract the pair into one FFMA2.
              const float2 z = __ffma2_rn(q[(ii - i) & 7], make_float2(wc[H][i], wc[H][i]), neg0);
              y = i == 0 ? z : __ffma2_rn(y, one, z);
            }
            ypair[ii] = y;
          }
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int chunk = (16 * warp + 4 * o + h) ^ (lane & 7);
            *reinterpret_cast<float4 *>(yrow + 4 * chunk) = make_float4(ypair[2 * h].x, ypair[2 * h].y, ypair[2 * h + 1].x, ypair[2 * h + 1].y);
          }
        }
      }
      if (H == 0) __syncwarp(); else __syncthreads();
      if (H == 1 && tile + 1 < n_tiles) {
        // look-back of the next tile = last 15 rows of this one; each warp moves the rows its own cp.async is about to
        // overwrite (program order inside the warp), so no barrier is needed in between
        for (int r = kTile + ((warp - (kTile - kLook)) & (kFbWarps - 1)); r < kPRows; r += kFbWarps) P[(r - kTile) * 32 + lane] = P[r * 32 + lane];
        // Common case (the next tile's rows lie wholly in this pass's PCM): the 64 cp.async of this warp are not issued
        // in one burst — which fills the memory-instruction queue that the matrixing's LDS.128 of all resident CTAs go
        // through — but two per n inside the matrixing loop below.
        const int ra = kTile * (tile + 1) + kLook, rb = min(kTile * (tile + 2) + kLook, rows_total);
        const int64_t rel0 = (int64_t)(n_start + 32 * ra + 1152) * ch - (int64_t)pv.head_n;
        if (!joint && ra < rb && rel0 >= 0 && rel0 + (int64_t)(rb - ra) * 32 * ch <= (int64_t)pv.cur_n) {
          pf_src = pv.cur + rel0 + (ch == 1 ? lane : 2 * lane + c) + (size_t)warp * 32 * ch;
          pf_dst = (uint32_t)__cvta_generic_to_shared(P + (kLook + warp) * 32 + lane);
          pf_left = (rb - ra - warp + kFbWarps - 1) / kFbWarps;
        } else {
          load_rows(ra, rb, kLook);
        }
      }
      // ---- matrixing (SRC:1402-1408): S[k] = sum over ascending n of M[k][n] * Y[n], one fused multiply-add per term.
      // PFM = how the next tile's PCM rows are fetched from inside the loop (second half only): 0 nothing, 1 / 2 the
      // whole tile is contiguous stereo / mono PCM of this pass (64 rows per warp: two cp.async per n whose addresses are
      // instruction immediates — no pointer arithmetic, no branches), 3 whatever is left, one predicated copy at a time.
      if (64 * warp < valid && !(dbg & 2)) {
        auto matrixing = [&](auto Mc) {
          constexpr int PFM = decltype(Mc)::value;
          const float4 *mrow = reinterpret_cast<const float4 *>(sM + (32 * H) * 32 + kg * 16);
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            static_for<0, 8>([&](auto bc) {
              constexpr int b = decltype(bc)::value;
              const int nl = 8 * a + b;
              if constexpr (PFM == 1 || PFM == 2) {
                constexpr int SB = kFbWarps * 32 * 4 * (PFM == 1 ? 2 : 1);   // bytes between the rows of one warp
                cp_async4_imm<(2 * b) * kFbWarps * 128, (2 * b) * SB>(pf_dst, pf_src);
                cp_async4_imm<(2 * b + 1) * kFbWarps * 128, (2 * b + 1) * SB>(pf_dst, pf_src);
              } else if constexpr (PFM == 3) {
#pragma unroll
                for (int u = 0; u < 2; ++u)
                  if (pf_left > 0) {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(pf_dst), "l"(pf_src));
                    pf_src += pf_step; pf_dst += kFbWarps * 128; --pf_left;
                  }
              }
              const float4 m0 = mrow[nl * 8], m1 = mrow[nl * 8 + 1], m2 = mrow[nl * 8 + 2], m3 = mrow[nl * 8 + 3];
              const float4 y = *reinterpret_cast<const float4 *>(Y + a * 8 * kTile + yo[b]);
              const float yv[4] = {y.x, y.y, y.z, y.w};
              const float mv[16] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w, m2.x, m2.y, m2.z, m2.w, m3.x, m3.y, m3.z, m3.w};
#pragma unroll
              for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[j][i] = __ffma2_rn(make_float2(yv[j], yv[j]), make_float2(mv[2 * i], mv[2 * i + 1]), acc[j][i]);
            });
            if constexpr (PFM == 1 || PFM == 2) { pf_src += 16 * pf_step; pf_dst += 16 * kFbWarps * 128; }
          }
          if constexpr (PFM == 1 || PFM == 2) pf_left = 0;
        };
        if (H == 0 || pf_left == 0) matrixing(IntC<0>{});
        else if (pf_left == kTile / kFbWarps) { if (ch == 2) matrixing(IntC<1>{}); else matrixing(IntC<2>{}); }
        else matrixing(IntC<3>{});
      }
    }
    while (pf_left > 0) {                           // (a warp that skipped the matrixing of a short tile)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(pf_dst), "l"(pf_src));
      pf_src += pf_step; pf_dst += kFbWarps * 128; --pf_left;
    }
    asm volatile("cp.async.commit_group;");
    // subband samples -> HBM.  A thread owns 16 of the 32 subbands of four steps; written directly that is 32 scattered
    // 16-byte pieces per store instruction.  The warp's own 8 KB of Y (it alone reads it, and is done with it) is used to
    // transpose: 16-byte chunk c of step row t goes to slot t * 8 + (c ^ ((t >> 2) & 7)) — conflict-free both ways — and
    // comes back as whole 128-byte rows, four rows per store instruction.
    if (64 * warp < valid) {
      __syncwarp();
      auto slot = [&](int t, int cidx) {             // float offset of a chunk inside the warp's slice of Y
        const int q = t * 8 + (cidx ^ ((t >> 2) & 7));
        return (q >> 4) * kTile + ((16 * warp + (q & 15)) << 2);
      };
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<float4 *>(Y + slot(4 * tg + j, 4 * kg + i)) = make_float4(acc[j][2 * i].x, acc[j][2 * i].y, acc[j][2 * i + 1].x, acc[j][2 * i + 1].y);
      __syncwarp();
      const int rows = min(64, valid - 64 * warp);
      // row t = 4 it + (lane >> 3), chunk lane & 7: slot(t, chunk) = fixed per lane + it * 2 kTile + ((chunk ^ (it & 7)) << 2)
      const int l3 = lane >> 3, cidx = lane & 7;
      float4 *dst = reinterpret_cast<float4 *>(out + (size_t)(kTile * tile + 64 * warp) * 32) + l3 * 8 + cidx;
      const float *ysrc = Y + (l3 >> 1) * kTile + ((16 * warp + 8 * (l3 & 1)) << 2);
#pragma unroll
      for (int it = 0; it < 16; ++it)
        if (4 * it + l3 < rows && !(dbg & 8)) dst[it * 32] = *reinterpret_cast<const float4 *>(ysrc + it * 2 * kTile + ((cidx ^ (it & 7)) << 2));
    }
  }
}

}  // namespace mp3b
#include "filterbank_tc.cuh"   // K1 with the matrixing on the tensor cores (tcgen05, opt-in)
namespace mp3b {

// ------------------------------------------------------------------------------------------------------------
// K2+K4: MDCT / alias reduction (SRC:1512-1662), then the bits-vs-gain curve of quantizeToFitBudget (SRC:734-794).
// One warp per gc, everything between the subband samples and the curve stays in the warp.  The loop's gain sequence does not
// depend on maxBits (only where it stops does), so the warp evaluates it until the bit count fits the smallest
// budget the frame can have (no reservoir, no padding) or the loop's own exits fire; the serial scan (K_scan) then
// only looks entries up.
__device__ __forceinline__ int lo_bits_of(const Config &cfg, int bri) {
  return ((cfg.frame_base[bri] - cfg.header_bytes) * 8) >> cfg.channels;     // / (2 * channels), channels = 1 or 2; the dividend is positive
}

constexpr int kGranulePerWarp = 1;   // 4 measured slower (6.0 vs 5.7 ms): one granule-channel per warp keeps the tail short
// TRACE: also leave the MDCT spectrum behind.  PRE: no k_prepass ran (CBR, not joint stereo, no trace: nothing but the block
// type is needed from the PCM) — the warp reads its granule's 576 samples itself and decides the block type (SRC:1944-1968).
// CH: the channel count as a compile-time constant (0 = cfg.channels) — the product variant's 18 PCM loads then carry their
// strides as immediates instead of computing 18 addresses.
constexpr int kGrWarps = 8;                      // warps (= granule-channels) per CTA (4 and 16 measured slower: 39.4 / 41.4 vs 38.4 ms)
template <bool TRACE, bool PRE, bool ISO, int CH = 0> __global__ void __launch_bounds__(32 * kGrWarps, 32 / kGrWarps) k_granule(Config cfg, PassBuffers pb) {
  __shared__ __align__(16) uint8_t len31[ISO ? 16 : 31 * 32];   // table-15 code length of a pair + its sign bits (SRC:828-853), indexed by quant30
  __shared__ __align__(16) uint8_t iso_len[ISO ? (kHuffEntries + 15) / 16 * 16 : 16];   // ISO mode: all Huffman length tables
  __shared__ uint8_t iso_c[ISO ? kGrWarps : 1][ISO ? 288 : 1];
  __shared__ uint32_t iso_q[ISO ? kGrWarps : 1][ISO ? 288 : 1];   // ISO mode: the quantized pairs of the gain being evaluated
  __shared__ uint16_t s_spos[ISO ? 192 : 1];                    // ISO short blocks: line order by scalefactor band and window
  __shared__ uint8_t s_swid[ISO ? 192 : 1];
  __shared__ __align__(8) float smg[kGrWarps][576];
  const int s = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ch = CH ? CH : cfg.channels, chs = ch - 1;   // channels = 1 or 2: / ch is >> chs
  // a warp walks kGranulePerWarp granule-channels: the table staging below and the CTA start-up are paid once for all
  for (int rep = 0; rep < kGranulePerWarp; ++rep) {
  const int gci = (blockIdx.y * kGranulePerWarp + rep) * kGrWarps + warp;
  // ---- MDCT (SRC:1512-1565): lane = subband; the 36 time samples are the previous and the current granule's rows of
  // the subband array (row 18 (g + 1) + t = step t of granule g; rows 0..17 = last granule of the previous pass).
  // The first rows are requested before anything else — their addresses need nothing from memory —, the frame count and the
  // bitrate index right behind them.  On the fused path the last 14 rows are requested after the block-type decision: the
  // MDCT needs them ~400 instructions later, and until then their registers hold the PCM samples of the decision (all 54
  // loads in flight at once spill, and a spill store of a loaded value holds up every load behind it).
  constexpr int kEarlyRows = PRE ? 22 : 36;
  float v[36];
  const float *prev_late;
  {
    const int gcl = min(gci, pb.GC - 1);             // the grid is rounded up to whole CTAs: stay inside the array
    const int g = gcl >> chs, c = gcl & chs;
    const float *prev = pb.sub + ((size_t)(s * ch + c) * pb.sub_rows + 18 * g) * 32 + lane;
#pragma unroll
    for (int k = 0; k < kEarlyRows; ++k) v[k] = __ldg(prev + k * 32);
    prev_late = prev;
  }
  // (requested here, not where they are used: after the barrier below their latency would stand in front of everything)
  const int n_gc = (int)pb.plan[s].n_frames * 2 * ch;
  const int bri_f = pb.frame_br[(size_t)s * pb.Fc + min(gci >> (chs + 1), pb.Fc - 1)];
  float pc[PRE ? 18 : 1];                            // PRE: sample 32 k + lane of the granule-channel
  if (PRE) {
    const PcmView pv = pcm_view(cfg, pb, s);
    const int f = gci >> (chs + 1), gr = (gci >> chs) & 1, c = gci & chs;
    const int64_t q0 = (int64_t)(1 + f) * cfg.fsc + (int64_t)(gr * 576 + lane) * ch + c;   // element of sample `lane`
    const int64_t rel = q0 - (int64_t)pv.head_n;
    if (rel >= 0 && rel + (int64_t)(17 * 32 * ch) < (int64_t)pv.cur_n) {                     // the usual case: all of it in this pass's PCM
      const float *p = pv.cur + rel;
#pragma unroll
      for (int k = 0; k < 18; ++k) pc[k] = __ldg(p + k * 32 * ch);
    } else {
#pragma unroll
      for (int k = 0; k < 18; ++k) pc[k] = pv.at(q0 + (int64_t)k * 32 * ch);
    }
  }
  if (rep == 0) {
    if (ISO) {
      for (int i = threadIdx.x; i < kHuffEntries; i += 32 * kGrWarps) iso_len[i] = kHuffLenFlat[i];
      for (int i = threadIdx.x; i < 192; i += 32 * kGrWarps) { s_spos[i] = c_short_pos[cfg.sfb_index][i]; s_swid[i] = c_short_width[cfg.sfb_index][i]; }
    }
    else for (int i = threadIdx.x; i < 31 * 32 / 4; i += 32 * kGrWarps) reinterpret_cast<uint32_t *>(len31)[i] = reinterpret_cast<const uint32_t *>(tab::kLen31s)[i];
    __syncthreads();
  }
  if (gci >= n_gc) return;
  const size_t gslot = (size_t)s * pb.GC + gci;
  const int lo_bits = lo_bits_of(cfg, bri_f);
  int bt_gc = 0;
  {
    int bt;
    if (PRE) {
      // TransientDetector.analyze: thirds of 192 samples = 6 rows of 32; [OD1b] partial = lane, ascending rows, butterfly tree
      float e3[3]; int sbg[3];
#pragma unroll
      for (int th = 0; th < 3; ++th) {
        float a = 0.0f;
#pragma unroll
        for (int jj = 0; jj < 6; ++jj) a = __fmaf_rn(pc[6 * th + jj], pc[6 * th + jj], a);
        e3[th] = a;
      }
      lane_tree_pair(e3[0], e3[1], lane);
      e3[2] = lane_tree(e3[2]);
#pragma unroll
      for (int th = 0; th < 3; ++th) e3[th] = div192(e3[th]);
      transient_decide<false>(e3, bt, sbg);      // the trace plane, which shows subblock_gain of every granule, runs k_prepass
      if (lane == 0) pb.gc_bt[gslot] = (uint16_t)(bt | sbg[0] << 2 | sbg[1] << 5 | sbg[2] << 8);
    } else if (ISO) {
      if (cfg.iso >= 3) bt = pb.gc_bt[gslot] & 3;                 // level 3: ISO block type 0 / 1 start / 2 short / 3 stop from k_iso_blocktype
      else { bt = 0; if (lane == 0) pb.gc_bt[gslot] = 0; }        // levels 1, 2: long blocks only
    } else bt = pb.gc_bt[gslot] & 3;
    bt_gc = bt;
#pragma unroll
    for (int k = kEarlyRows; k < 36; ++k) v[k] = __ldg(prev_late + k * 32);
    float *X = smg[warp];
    const int sb = lane;
    const bool flip = sb & 1;
    const bool use_long = ISO ? bt != 2 : (bt == 0 || (bt == 1 && sb < 2));      // SRC:1542-1553 (ISO: start and stop are long transforms)
    if (use_long) {                                             // mdctLong SRC:1619-1636
      float a[18];
#pragma unroll
      for (int m = 0; m < 18; ++m) a[m] = 0.0f;
#pragma unroll
      for (int k = 0; k < 36; ++k) {
        float x = v[k];
        if (flip && (k & 1)) x = -x;                            // SRC:1520-1524
        float w = __fmul_rn(x, ISO ? c_iso_win[bt][k] : tab::kWinLong[k]);
#pragma unroll
        for (int m = 0; m < 18; ++m) a[m] = __fmaf_rn(w, tab::kMdctLong[m][k], a[m]);
      }
#pragma unroll
      for (int m = 0; m < 18; ++m) X[sb * 18 + m] = div_exact(a[m], 9.0f, 1.0f / 9.0f);
    }
    if (!use_long) {                                            // mdctShort SRC:1639-1662
#pragma unroll
      for (int w3 = 0; w3 < 3; ++w3) {
        float sg[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          const int k = w3 * 6 + 6 + i;
          float x = v[k];
          if (flip && (k & 1)) x = -x;
          sg[i] = __fmul_rn(x, tab::kWinShort[i]);
        }
#pragma unroll
        for (int m = 0; m < 6; ++m) {
          float r = 0.0f;
#pragma unroll
          for (int k = 0; k < 12; ++k) r = __fmaf_rn(sg[k], tab::kMdctShort[m][k], r);
          // the reference interleaves the windows inside the subband; ISO orders by scalefactor band, then window
          const int dst = ISO ? s_spos[6 * sb + m] + w3 * s_swid[6 * sb + m] : sb * 18 + w3 + 3 * m;
          X[dst] = div_exact(r, 3.0f, 1.0f / 3.0f);
        }
      }
    }
    __syncwarp();
    if ((ISO ? bt != 2 : bt == 0) && sb < 31) {                 // applyAliasingReduction SRC:1581-1616 [OD5]
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int iu = sb * 18 + 17 - i, il = (sb + 1) * 18 + i;
        float upper = X[iu], lower = X[il];
        X[iu] = __fadd_rn(__fmul_rn(lower, tab::kAliasCa[i]), __fmul_rn(upper, tab::kAliasCs[i]));
        X[il] = __fsub_rn(__fmul_rn(lower, tab::kAliasCs[i]), __fmul_rn(upper, tab::kAliasCa[i]));
      }
    }
    __syncwarp();
  }
  // |x|^0.75 (SRC:805-813 [OD3]), peak -> g0 (SRC:989-1006), preflag (SRC:2042-2066); line i = lane + 32 j
  uint32_t meta;
  {
    float *spec = TRACE ? pb.spec + gslot * 576 : nullptr;       // trace plane only; no predicated-off stores otherwise
    float *smag = pb.smag + gslot * 576;
    float x[18];
#pragma unroll
    for (int j = 0; j < 18; ++j) x[j] = smg[warp][lane + 32 * j];
    __syncwarp();
    if (TRACE) {
#pragma unroll
      for (int j = 0; j < 18; ++j) spec[lane + 32 * j] = x[j];
    }
    float peak = 0.0f, plo = 0.0f, phi = 0.0f;
#pragma unroll
    for (int j = 0; j < 18; ++j) {
      const int i = lane + 32 * j;
      const float ax = fabsf(x[j]);
      peak = fmaxf(peak, ax);
      // [OD1b]: partial index = (i - segment start) mod 32; 432 mod 32 = 16 is an xor permutation of the lanes, which
      // the butterfly tree is invariant under.
      if (i < 432) plo = __fmaf_rn(x[j], x[j], plo); else phi = __fmaf_rn(x[j], x[j], phi);
      const float mag = pow34(fmaxf(ax, 1e-10f));
      smag[i] = __uint_as_float(__float_as_uint(mag) | (__float_as_uint(x[j]) & 0x80000000u));   // sign of x on mag > 0 (a -0 line quantizes to 0 either way: no sign bit is coded)
      smg[warp][i] = mag;
    }
    peak = warp_max_nonneg(peak);
    float low = plo, high = phi;
    lane_tree_pair(low, high, lane);
    // every lane holds the same peak and energies: all of them derive g0 / preflag (no broadcast, no read-back)
    meta = (uint32_t)gain_from_peak(peak) | (uint32_t)(high > __fmul_rn(low, 1.5f) ? 1 : 0) << 17;
    __syncwarp();
  }
  float mx[9], my[9];
  float2 mxy[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) { mxy[j] = reinterpret_cast<const float2 *>(smg[warp])[lane + 32 * j]; mx[j] = mxy[j].x; my[j] = mxy[j].y; }
  uint16_t *bits_out = pb.gc_bits + gslot * kMaxEntries, *bv_out = pb.gc_bv + gslot * kMaxEntries;
  if (ISO && cfg.iso >= 2) { __syncwarp(); continue; }   // level 2: k_outer (iso_psy.cuh) takes it from the magnitudes
  if (ISO) {
    // global_gain search (north_star stage 4).  The bit count falls as the gain rises, so: binary search for the smallest gain
    // that fits the LARGEST budget this frame can have (full reservoir, padded), then the curve gain by gain until the count
    // fits the smallest (no reservoir, no padding); the serial scan, which knows the reservoir, picks the first entry that
    // fits.  If the two are more than 18 steps apart, the last entry is the binary-searched gain of the smallest budget.
    const int f2 = gci >> (chs + 1);
    const int bri = pb.frame_br[(size_t)s * pb.Fc + f2];
    const int mds1 = cfg.frame_base[bri] + 1 - cfg.header_bytes;
    const int hi_bits = min(4095, (mds1 * 8 + (min(511, mds1) * 8 * 9) / 10) >> cfg.channels);
    const int lo_fit = min(lo_bits, 4095);
    const int *sfb = c_sfb_cum[cfg.sfb_index];
    const bool ws = bt_gc != 0;
    // the magnitudes are read from the warp's tile (pair p = lines 2 p, 2 p + 1 as one float2) by the shared evaluation function
    const float2 *mt = reinterpret_cast<const float2 *>(smg[warp]);
    auto eval = [&](int G) { return iso_eval_gain<9>(G, mt, iso_q[warp], lane, iso_len, iso_c[warp], sfb, ws); };   // bits | big_values << 16
    int lo = 0, hi = kIsoGainMax;                     // invariant: the count at `hi` fits (at kIsoGainMax every line quantizes to 0)
    while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)(eval(mid) & 0xFFFFu) <= hi_bits) hi = mid; else lo = mid + 1; }
    const int g_first = hi;
    int n = 0, g_last = g_first, fitted = 0;
    for (int e = 0; e < kMaxEntries - 1 && !fitted; ++e) {
      const int G = min(g_first + e, kIsoGainMax);
      const uint32_t c = eval(G);
      if (lane == 0) { bits_out[e] = (uint16_t)(c & 0xFFFFu); bv_out[e] = (uint16_t)(c >> 16); }
      n = e + 1; g_last = G;
      fitted = (int)(c & 0xFFFFu) <= lo_fit || G == kIsoGainMax;
    }
    if (!fitted) {
      lo = min(g_first + kMaxEntries - 1, kIsoGainMax); hi = kIsoGainMax;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)(eval(mid) & 0xFFFFu) <= lo_fit) hi = mid; else lo = mid + 1; }
      const uint32_t c = eval(hi);
      if (lane == 0) { bits_out[kMaxEntries - 1] = (uint16_t)(c & 0xFFFFu); bv_out[kMaxEntries - 1] = (uint16_t)(c >> 16); }
      n = kMaxEntries; g_last = hi;
    }
    if (lane == 0) pb.gc_meta[gslot] = (uint32_t)g_first | (uint32_t)n << 9 | (uint32_t)g_last << 14;   // 9 + 5 + 9 bits
    __syncwarp();
    continue;
  }
  const int g0 = meta & 255;
  int gain = g0, n = 0, restart = 0;
  for (int it = 0; it < kMaxEntries; ++it) {
    const float inv2 = __fmul_rn(c_inv_step[gain], 2.0f);
    int total = 0, last = 0;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const int idx = pair_index2(mxy[j], inv2);
      total += len31[idx];
      if (idx) last = lane + 32 * j + 1;
    }
    const int bv = warp_max_i(last);                 // pairs up to and including the last non-zero one, SRC:750-763
    const int bits = warp_sum_i(total) - 3 * (288 - bv);   // all-zero pairs beyond big_values are not coded (len15[0][0] = 3)
    if (lane == 0) { bits_out[it] = (uint16_t)bits; bv_out[it] = (uint16_t)bv; }
    n = it + 1;
    if (bv == 0 && it == 0) { restart = 1; gain = max(gain - 40, 0); continue; }   // SRC:758-761
    if (bits <= lo_bits) break;                      // fits every possible budget of this frame
    int next = min(gain + 4, 255);                   // SRC:772-775
    if (next >= 255) break;
    gain = next;
  }
  if (lane == 0) pb.gc_meta[gslot] = meta | (uint32_t)n << 8 | (uint32_t)restart << 16;
  __syncwarp();
  }
}

// ISO mode level 3: block types with one granule of look-ahead (iso_mode.cuh).  One warp per granule of the pass (both
// channels): attack flags of PCM granules g - 2, g - 1, g, the reference's detector on each channel, either channel switches both.
__global__ void __launch_bounds__(128) k_iso_blocktype(Config cfg, PassBuffers pb) {
  const int s = blockIdx.x, lane = threadIdx.x & 31;
  const int g = blockIdx.y * 4 + (threadIdx.x >> 5);
  const int nfr = (int)pb.plan[s].n_frames, ch = cfg.channels;
  if (g >= 2 * nfr) return;
  const PcmView pv = pcm_view(cfg, pb, s);
  bool a[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const int j = g - 2 + d;                                      // PCM granule of the pass (j = -2, -1: the carried frame)
    bool att = false;
    for (int c = 0; c < ch; ++c) {
      float e3[3];
#pragma unroll
      for (int th = 0; th < 3; ++th) {
        float acc = 0.0f;
#pragma unroll
        for (int jj = 0; jj < 6; ++jj) {
          const float x = pv.at((int64_t)cfg.fsc + ((int64_t)576 * j + 192 * th + 32 * jj + lane) * ch + c);
          acc = __fmaf_rn(x, x, acc);
        }
        e3[th] = __fdiv_rn(lane_tree(acc), 192.0f);
      }
      int bt, sbg[3];
      transient_decide(e3, bt, sbg);
      att |= bt != 0;
    }
    a[d] = att;
  }
  const int bt = a[1] ? 2 : a[2] ? (a[0] ? 2 : 1) : a[0] ? 3 : 0;
  if (lane < ch) pb.gc_bt[(size_t)s * pb.GC + (size_t)g * ch + lane] = (uint16_t)bt;
  if (lane == 0 && (g & 1) == 0) pb.ms[(size_t)s * (pb.Fc + 1) + 1 + (g >> 1)] = 0;   // level 3 codes L / R (a frame's samples straddle two input frames)
}

}  // namespace mp3b
#include "iso_psy.cuh"    // ISO mode level 2: psychoacoustic model (k_psy) and scalefactor outer loop (k_outer)
namespace mp3b {

// ------------------------------------------------------------------------------------------------------------
// K_scan: the serial part of a stream (SRC:475-568): padding, reservoir, per-gc gain choice, main_data_begin,
// slot filling bookkeeping.  One thread per stream.
__device__ __forceinline__ void region_counts(const Config &cfg, int big_values, int &r0, int &r1) {  // SRC:856-887
  const int region = big_values * 2;
  const int *b = c_sfb_cum[cfg.sfb_index];
  int region0 = 0;
  for (int i = 0; i < 15; ++i) { if (b[i] <= region) region0 = i; else break; }
  int region1 = 0, start = region0 + 1, lim = min(start + 7, 21);
  for (int i = start; i < lim; ++i) { if (b[i] <= region) region1 = i - region0 - 1; else break; }
  r0 = min(region0, 15); r1 = min(region1, 7);
}

// One warp per stream.  Frames are taken 32 at a time: (1) the warp stages the chunk's curve tables in shared memory
// with coalesced loads, (2) lane 0 walks the 32 frames serially out of shared memory — this is the only truly
// sequential work of the whole encoder —, (3) lane l writes the records of frame l.
struct ScanFrame { int padding, mdb, res_bits, bpg, huff, is_final; uint32_t w_off, e_src, e_take, e_out; };

__global__ void __launch_bounds__(32) k_scan(Config cfg, PassBuffers pb) {
  __shared__ __align__(16) uint16_t sh_bits[32 * 4 * kMaxEntries];
  __shared__ uint32_t sh_meta[32 * 4];
  __shared__ uint8_t sh_bri[32];
  __shared__ __align__(4) uint8_t sh_sel[32 * 4][4];       // chosen entry, gain_out, gain_used, iterations
  __shared__ ScanFrame sh_fr[32];
  __shared__ int sh_state[8];
  __shared__ float sh_pe[32 * 4];                            // ISO mode level 2: perceptual entropy per gc
  __shared__ uint16_t sh_bpg[32 * 4];                        // the budget each gc was given
  __shared__ int16_t sh_mds[32];                             // fast path: main-data bytes of the frame (padding included), -1: final frame
  __shared__ uint32_t sh_ser[32][2];                         // fast path: what the serial loop leaves per frame: huff | avail << 16, W before
  const int s = blockIdx.x, lane = threadIdx.x;
  const StreamPlan plan = pb.plan[s];
  StreamState &st = pb.state[s];
  const int ch = cfg.channels, nf = (int)plan.n_frames, ngc = 2 * ch;
  FrameRec *rec = pb.rec + (size_t)s * (pb.Fc + 1);
  FrameEmit *emit = pb.emit + (size_t)s * (pb.Fc + 1);
  uint16_t *emit_size = pb.emit_size + (size_t)s * (pb.Fc + 1);
  // rec[0] = bufferedFrame of the previous pass
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(&st.buffered);
    uint32_t *dst = reinterpret_cast<uint32_t *>(&rec[0]);
    for (int i = lane; i < (int)(sizeof(FrameRec) / 4); i += 32) dst[i] = src[i];
  }
  // serial state lives in lane 0's registers
  uint32_t out_pos = (plan.flags & 4) ? 0u : st.out_pos;
  const uint32_t B0 = (uint32_t)st.backlog;
  uint32_t R = 0, W = B0, n_emit = 0, frame_count = st.frame_count, total_bytes = st.total_bytes;
  int avail = st.avail_bytes, pad_rem = st.pad_rem, err = 0;
  int prev_slot = st.buffered.valid ? (int)st.buffered.slot : -1;     // slot of the buffered frame, -1 = none
  const int first_emit = prev_slot >= 0 ? 0 : 1;                      // rec index of the first frame this pass can emit
  __syncwarp();
  for (int base = 0; base < nf; base += 32) {
    const int cnt = min(32, nf - base);
    {  // (1) stage
      const size_t g0s = (size_t)s * pb.GC + (size_t)base * ngc;
      const uint4 *src = reinterpret_cast<const uint4 *>(pb.gc_bits + g0s * kMaxEntries);   // 40-byte rows, 16-byte aligned chunk start
      uint4 *dst = reinterpret_cast<uint4 *>(sh_bits);
      const int n16 = cnt * ngc * kMaxEntries * 2 / 16;
      for (int i = lane; i < n16; i += 32) dst[i] = src[i];
      for (int i = n16 * 8 + lane; i < cnt * ngc * kMaxEntries; i += 32) sh_bits[i] = pb.gc_bits[g0s * kMaxEntries + i];
      for (int i = lane; i < cnt * ngc; i += 32) sh_meta[i] = pb.gc_meta[g0s + i];
      if (lane < cnt) sh_bri[lane] = pb.frame_br[(size_t)s * pb.Fc + base + lane];
      if (cfg.iso >= 2) for (int i = lane; i < cnt * ngc; i += 32) sh_pe[i] = pb.gc_psy[(g0s + i) * 24 + 22];
    }
    __syncwarp();
    if (!cfg.iso) {
      // (2) fast path (the reference-compatible mode).  One warp runs the serial chain at its own issue latency — every instruction
      // in the loop costs the stream about six cycles per frame — so the loop keeps only what truly depends on the reservoir:
      // budget, curve look-up (lane j = gc j), bit total, reservoir, FIFO cursors.  Padding / slot sizes come before it as a
      // parallel prefix (lane = frame), the emission records and counters after it, again lane = frame.
      const int ngc_shift = ch == 1 ? 1 : 2;
      int pad_l = 0, mds_l = 0;
      {  // (2a) shouldPad SRC:456-463 for 32 frames at once: rem < sample_rate, so the accumulator wraps at most once per frame and
         // padding = floor(cum / sr) - floor(cum_prev / sr) with cum the running sum of the remainders on top of the carried one
        const int bri = lane < cnt ? sh_bri[lane] : 0;
        int cum = lane < cnt ? cfg.frame_rem[bri] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, cum, d); if (lane >= d) cum += t; }
        cum += pad_rem;
        const int q = cum / cfg.sample_rate;
        int qp = __shfl_up_sync(0xffffffffu, q, 1);
        if (lane == 0) qp = 0;
        pad_l = q - qp;
        mds_l = cfg.frame_base[bri] + pad_l - cfg.header_bytes;       // SRC:497
        pad_rem = __shfl_sync(0xffffffffu, cum - q * cfg.sample_rate, cnt - 1);
        if (lane < cnt) sh_mds[lane] = (int16_t)mds_l;
      }
      __syncwarp();
      const int prev_slot0 = prev_slot;
      const uint32_t R0 = R;
#pragma unroll 2
      for (int l = 0; l < cnt; ++l) {
        const int mds = sh_mds[l];
        const bool is_final = (plan.flags & 1) && base + l == nf - 1;
        const int res_bits = is_final ? 0 : avail * 8;               // SRC:500
        const int bpg = (mds * 8 + (res_bits * 9) / 10) >> ngc_shift;  // SRC:647-650
        int total;
        {
          const int j = lane & (ngc - 1);
          const uint32_t meta = sh_meta[l * ngc + j];
          const int g0 = (int)(meta & 255u), n = (int)((meta >> 8) & 255u), restart = (int)((meta >> 16) & 1u);
          const uint16_t *cb = sh_bits + (l * ngc + j) * kMaxEntries;
          int gain = g0, chosen = n - 1, gain_out = g0, gain_used = g0, iters = n, bits = -1;
          for (int e = 0; e < n; ++e) {                              // quantizeToFitBudget SRC:745-776
            gain_used = gain;
            if (e == 0 && restart) { gain = max(gain - 40, 0); continue; }
            const int be = cb[e];
            if (be <= bpg) { chosen = e; gain_out = gain; iters = e + 1; bits = be; break; }
            int next = min(gain + 4, 255);
            if (next >= 255 || e == kMaxEntries - 1) { chosen = e; gain_out = next; iters = e + 1; bits = be; break; }
            if (e == n - 1) { chosen = e; gain_out = next; err |= 1; bits = be; break; }   // curve ended early: engine bug
            gain = next;
          }
          if (bits < 0) bits = cb[chosen];
          if (lane < ngc)
            *reinterpret_cast<uint32_t *>(sh_sel[l * ngc + j]) = (uint32_t)chosen | (uint32_t)gain_out << 8 | (uint32_t)gain_used << 16 | (uint32_t)iters << 24;
          total = bits;
        }
        total += __shfl_xor_sync(0xffffffffu, total, 1);
        if (ngc == 4) total += __shfl_xor_sync(0xffffffffu, total, 2);
        const int huff = (total + 7) >> 3;                           // padToByte SRC:729
        if (lane < 2) sh_ser[l][lane] = lane == 0 ? ((uint32_t)huff | (uint32_t)avail << 16) : W;
        W += (uint32_t)huff;                                         // appendHuffmanData SRC:511
        if (W > pb.md_stride) { err |= 2; W = (uint32_t)pb.md_stride; }
        const int a = avail + mds - huff;                            // updateReservoir SRC:565, 2125-2128
        avail = a < 0 ? 0 : a > 511 ? 511 : a;
      }
      __syncwarp();
      {  // (2b) lane = frame: FIFO read cursor, emission of the buffered frame (SRC:548-556, fillSlot 2110-2121), counters
        const uint32_t huff = lane < cnt ? (sh_ser[lane][0] & 0xFFFFu) : 0u, avail_l = lane < cnt ? (sh_ser[lane][0] >> 16) : 0u;
        const uint32_t w_before = lane < cnt ? sh_ser[lane][1] : 0u;
        const uint32_t w_after = min(w_before + huff, (uint32_t)pb.md_stride);
        int pslot = __shfl_up_sync(0xffffffffu, mds_l, 1);           // the frame emitted while frame l is encoded is frame l - 1
        if (lane == 0) pslot = prev_slot0;
        // R_l = min(R_{l-1} + slot_{l-1}, W_l): a serial recurrence, but over 32 lanes of registers, not in the loop above
        uint32_t r_before = R0;
        {
          uint32_t r = R0;
          for (int k = 0; k < cnt; ++k) {
            const int ps = __shfl_sync(0xffffffffu, pslot, k);
            const uint32_t wa = __shfl_sync(0xffffffffu, w_after, k);
            if (lane == k) r_before = r;
            if (ps >= 0) r = min(r + (uint32_t)ps, wa);
          }
          R = r;
        }
        const bool emits = lane < cnt && pslot >= 0;
        const uint32_t sz = emits ? (uint32_t)cfg.header_bytes + (uint32_t)pslot : 0u;
        uint32_t pre = sz;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, pre, d); if (lane >= d) pre += t; }
        const uint32_t sum_sz = __shfl_sync(0xffffffffu, pre, 31), n_em = (uint32_t)__popc(__ballot_sync(0xffffffffu, emits));
        if (lane < cnt) {
          const bool is_final = (plan.flags & 1) && base + lane == nf - 1;
          ScanFrame o;
          o.padding = pad_l; o.is_final = is_final; o.huff = (int)huff;
          o.mdb = is_final ? 0 : (int)min(w_before - r_before, 511u);                    // SRC:499, 2099-2101
          o.res_bits = is_final ? 0 : (int)avail_l * 8;
          o.bpg = (mds_l * 8 + (o.res_bits * 9) / 10) >> ngc_shift;
          o.w_off = w_before;
          o.e_take = emits ? min((uint32_t)pslot, w_after - r_before) : 0xFFFFFFFFu;
          o.e_src = emits ? r_before : 0u; o.e_out = emits ? out_pos + pre - sz : 0u;
          sh_fr[lane] = o;
          for (int j = 0; j < ngc; ++j) sh_bpg[lane * ngc + j] = (uint16_t)min(o.bpg, 65535);
        }
        out_pos += sum_sz; total_bytes += sum_sz; frame_count += n_em; n_emit += n_em;
        prev_slot = __shfl_sync(0xffffffffu, mds_l, cnt - 1);
      }
    } else
    {  // (2) the serial chain, SRC:475-568.  Every lane carries the scalar recurrences (padding, reservoir, cursors)
       // redundantly, lane j < ngc additionally walks the curve of gc j: the per-frame critical path is the reservoir
       // update plus one curve look-up plus two shuffles instead of four look-ups in sequence.
      const int ngc_shift = ch == 1 ? 1 : 2;
#pragma unroll 2
      for (int l = 0; l < cnt; ++l) {
        const int f = base + l;
        const bool is_final = (plan.flags & 1) && f == nf - 1;
        const int bri = sh_bri[l];
        int padding = 0;                                             // shouldPad SRC:456-463
        pad_rem += cfg.frame_rem[bri];
        if (pad_rem >= cfg.sample_rate) { pad_rem -= cfg.sample_rate; padding = 1; }
        const int mds = cfg.frame_base[bri] + padding - cfg.header_bytes;   // SRC:497
        // ISO mode: `avail` IS the back pointer (start of this slot minus start of this frame's data; the stuffing below keeps
        // the FIFO and the counter in step), and the final frame keeps its reservoir like any other
        const int mdb = cfg.iso ? avail : is_final ? 0 : (int)min(W - R, 511u);        // SRC:499, 2099-2101
        const int res_bits = (is_final && !cfg.iso) ? 0 : avail * 8; // SRC:500
        const int bpg = (mds * 8 + (res_bits * 9) / 10) >> ngc_shift;  // SRC:647-650: / (2 * channels)
        // every group of ngc lanes mirrors lanes 0..ngc-1 (lane L walks the curve of gc L mod ngc), so the butterfly below
        // leaves the frame's bit total in every lane without a broadcast
        int total;
        {
          const int j = lane & (ngc - 1);
          int my_bpg = bpg;
          if (cfg.iso >= 2) {
            // The granule's share of the reservoir follows its perceptual entropy (ISO 11172-3 C.1.5.4.5 in spirit): demand =
            // 3.1 PE bits beyond the mean; the granule-channels of a frame share at most 60 % of the reservoir in proportion to
            // their demand, what exceeds 80 % of the reservoir's capacity is spent anyway, nobody gets more than twice the mean
            // (k_outer's curve starts there).
            const int mean = (mds * 8) >> ngc_shift;
            const int want = max((int)fminf(sh_pe[l * ngc + j] * 3.1f, 60000.0f) - mean, 0);
            int sumwant = want + __shfl_xor_sync(0xffffffffu, want, 1);
            if (ngc == 4) sumwant += __shfl_xor_sync(0xffffffffu, sumwant, 2);
            const int pool = (res_bits * 6) / 10;
            const int give = sumwant > pool ? (int)((long long)want * pool / sumwant) : want;
            const int over = max(res_bits - (min(511, mds) * 64) / 10 - min(sumwant, pool), 0);
            my_bpg = min(min(mean + give + (over >> ngc_shift), 2 * mean), 4095);
          }
          if (lane < ngc) sh_bpg[l * ngc + j] = (uint16_t)min(my_bpg, 65535);
          const uint32_t meta = sh_meta[l * ngc + j];
          const int g0 = cfg.iso ? (int)(meta & 511u) : (int)(meta & 255u), n = cfg.iso ? (int)((meta >> 9) & 31u) : (int)((meta >> 8) & 255u);
          const int restart = cfg.iso ? 0 : (int)((meta >> 16) & 1u);
          const uint16_t *cb = sh_bits + (l * ngc + j) * kMaxEntries;
          int gain = g0, chosen = n - 1, gain_out = g0, gain_used = g0, iters = n, bits = -1;
          if (cfg.iso) {                                             // first entry whose count fits the budget and the 12-bit field
            const int fit = min(my_bpg, 4095);
            for (int e = 0; e < n; ++e) if ((int)cb[e] <= fit) { chosen = e; break; }
            bits = cb[chosen]; iters = chosen + 1;
            gain_used = chosen == kMaxEntries - 1 ? (int)((meta >> 14) & 511u) : min(g0 + chosen, kIsoGainMax);   // the search gain
            gain_out = gain_used = min(gain_used, 255);              // what the side info can say (iso_mode.cuh)
            if (bits > fit) err |= 1;
          } else
          for (int e = 0; e < n; ++e) {                              // quantizeToFitBudget SRC:745-776
            gain_used = gain;
            if (e == 0 && restart) { gain = max(gain - 40, 0); continue; }
            const int be = cb[e];
            if (be <= bpg) { chosen = e; gain_out = gain; iters = e + 1; bits = be; break; }
            int next = min(gain + 4, 255);
            if (next >= 255 || e == kMaxEntries - 1) { chosen = e; gain_out = next; iters = e + 1; bits = be; break; }
            if (e == n - 1) { chosen = e; gain_out = next; err |= 1; bits = be; break; }   // curve ended early: engine bug
            gain = next;
          }
          if (bits < 0) bits = cb[chosen];
          if (lane < ngc)
            *reinterpret_cast<uint32_t *>(sh_sel[l * ngc + j]) = (uint32_t)chosen | (uint32_t)gain_out << 8 | (uint32_t)gain_used << 16 | (uint32_t)iters << 24;
          total = bits;
        }
        total += __shfl_xor_sync(0xffffffffu, total, 1);
        if (ngc == 4) total += __shfl_xor_sync(0xffffffffu, total, 2);
        const int huff = (total + 7) >> 3;                           // padToByte SRC:729
        ScanFrame o;
        o.padding = padding; o.mdb = mdb; o.res_bits = res_bits; o.bpg = bpg; o.huff = huff; o.is_final = is_final;
        o.w_off = W;
        W += (uint32_t)huff;                                         // appendHuffmanData SRC:511
        if (cfg.iso) {                                               // stuffing: the reservoir may hold 511 bytes (9-bit pointer) and,
          const int over = avail + mds - huff - min(511, mds);       // with the one-frame delay of the FIFO, at most one slot
          if (over > 0) W += (uint32_t)over;                         // (zero bytes after the frame's data: ancillary to a decoder)
        }
        if (W > pb.md_stride) { err |= 2; W = (uint32_t)pb.md_stride; }
        o.e_take = 0xFFFFFFFFu; o.e_src = 0; o.e_out = 0;
        if (prev_slot >= 0) {                                        // emit the buffered frame, SRC:548-556 + fillSlot 2110-2121
          uint32_t take = min((uint32_t)prev_slot, W - R);
          o.e_src = R; o.e_take = take; o.e_out = out_pos;
          R += take;
          uint32_t sz = (uint32_t)cfg.header_bytes + (uint32_t)prev_slot;
          out_pos += sz; frame_count += 1; total_bytes += sz; n_emit += 1;
        }
        if (lane == 0) sh_fr[l] = o;
        prev_slot = mds;
        int a = avail + mds - huff;                                  // updateReservoir SRC:565, 2125-2128
        avail = a < 0 ? 0 : a > 511 ? 511 : a;
        if (cfg.iso) avail = min(avail, mds);
      }
    }
    __syncwarp();
    if (lane < cnt) {  // (3) records of frame f = base + lane
      const int f = base + lane;
      const ScanFrame o = sh_fr[lane];
      const int bri = sh_bri[lane];
      FrameRec fr;
      fr.valid = 1; fr.br_index = (uint8_t)bri; fr.padding = (uint8_t)o.padding;
      fr.ms = pb.ms[(size_t)s * (pb.Fc + 1) + 1 + f];
      const int mds = cfg.frame_base[bri] + o.padding - cfg.header_bytes;
      fr.mdb = (uint16_t)o.mdb; fr.slot = (uint16_t)mds; fr.is_final = (uint8_t)o.is_final; fr.pad0[0] = fr.pad0[1] = fr.pad0[2] = 0;
      fr.reservoir_bits = o.res_bits; fr.huff_bytes = o.huff; fr.frame_energy = pb.frame_energy[(size_t)s * pb.Fc + f];
      int total = 0;
      for (int j = 0; j < ngc; ++j) {
        const int gci = f * ngc + j;
        const size_t gslot = (size_t)s * pb.GC + gci;
        const uint32_t meta = sh_meta[lane * ngc + j];
        const int chosen = sh_sel[lane * ngc + j][0];
        const int bits = sh_bits[(lane * ngc + j) * kMaxEntries + chosen];
        const int bv = min((int)pb.gc_bv[gslot * kMaxEntries + chosen], 288);
        GcSide &g = fr.gc[j];
        g.part23 = (uint16_t)bits; g.big_values = (uint16_t)bv;
        g.global_gain = sh_sel[lane * ngc + j][1]; g.gain_used = sh_sel[lane * ngc + j][2];
        const uint16_t btw = pb.gc_bt[gslot];
        g.block_type = btw & 3; g.sbg[0] = (btw >> 2) & 7; g.sbg[1] = (btw >> 5) & 7; g.sbg[2] = (btw >> 8) & 7;
        int r0 = 0, r1 = 0;
        if (!cfg.iso) region_counts(cfg, bv, r0, r1);                // ISO mode: k_pack_iso fills regions, table_select, count1table
        g.region0 = (uint8_t)r0; g.region1 = (uint8_t)r1; g.preflag = cfg.iso ? 0 : (uint8_t)((meta >> 17) & 1); g.g0 = (uint8_t)(meta & 255);
        g.iterations = sh_sel[lane * ngc + j][3]; g.pad = 0; g.max_bits = sh_bpg[lane * ngc + j]; g.sfc = 0; g.part2 = 0;
        g.tsel[0] = g.tsel[1] = g.tsel[2] = 15; g.c1sel = 0;
        g.energy = pb.gc_energy[(size_t)s * (10 + pb.GC) + 10 + gci];
        if (cfg.iso) {                                                // the search gain (may exceed 255) | big_values << 9
          const int G = chosen == kMaxEntries - 1 ? (int)((meta >> 14) & 511u) : min((int)(meta & 511u) + chosen, kIsoGainMax);
          pb.gc_sel[gslot] = (uint32_t)G | (uint32_t)bv << 9;
          g.g0 = (uint8_t)min((int)(meta & 511u), 255); g.pad = (uint8_t)max(G - 255, 0);   // trace: gain_used + pad = the search gain
        } else pb.gc_sel[gslot] = (uint32_t)g.gain_used | (uint32_t)bv << 8;
        pb.gc_bitoff[gslot] = (uint32_t)total;
        total += bits;
      }
      for (int j = ngc; j < 4; ++j) fr.gc[j] = GcSide{};
      pb.fr_md[((size_t)s * pb.Fc + f) * 2] = o.w_off;
      pb.fr_md[((size_t)s * pb.Fc + f) * 2 + 1] = (uint32_t)o.huff;
      rec[1 + f] = fr;
      FrameEmit em;                                                  // emission of rec[f], decided while frame f was encoded
      em.emit = o.e_take != 0xFFFFFFFFu; em.src_off = em.emit ? o.e_src : 0; em.take = em.emit ? o.e_take : 0; em.out_off = em.emit ? o.e_out : 0;
      emit[f] = em;
    }
    __syncwarp();
  }
  __syncwarp();
  // tail: flush emission, state write-back
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) err |= __shfl_xor_sync(0xffffffffu, err, m);
  if (lane == 0) {
    FrameEmit em; em.emit = 0; em.src_off = em.take = em.out_off = 0;
    if ((plan.flags & 2) && prev_slot >= 0) {                        // flush SRC:335-347
      uint32_t take = min((uint32_t)prev_slot, W - R);
      em.emit = 1; em.src_off = R; em.take = take; em.out_off = out_pos;
      R += take;
      uint32_t sz = (uint32_t)cfg.header_bytes + (uint32_t)prev_slot;
      out_pos += sz; frame_count += 1; total_bytes += sz; n_emit += 1;
      prev_slot = -1;
    }
    emit[nf] = em;
    if (out_pos > pb.out_stride) err |= 4;
    if (W - R > (uint32_t)kMdCarryCap) err |= 8;
    st.out_pos = out_pos; st.frame_count = frame_count; st.total_bytes = total_bytes;
    st.avail_bytes = avail; st.pad_rem = pad_rem; st.error |= err;
    st.backlog = (int32_t)min(W - R, (uint32_t)kMdCarryCap);
    st.frames_total += (uint32_t)nf;
    pb.md_tail[(size_t)s * 4] = R; pb.md_tail[(size_t)s * 4 + 1] = (uint32_t)st.backlog; pb.md_tail[(size_t)s * 4 + 2] = B0;
    pb.emit_n[s] = n_emit;
    sh_state[0] = prev_slot;
  }
  __syncwarp();
  // sizes of the emitted frames, in order: rec[first_emit ...]
  const uint32_t ne = pb.emit_n[s];
  for (uint32_t k = lane; k < ne; k += 32) emit_size[k] = (uint16_t)(cfg.header_bytes + rec[first_emit + k].slot);
  // bufferedFrame for the next pass = last frame of this one (unless flushed)
  {
    const bool keep = sh_state[0] >= 0;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(&rec[nf]);
    uint32_t *dst = reinterpret_cast<uint32_t *>(&st.buffered);
    for (int i = lane; i < (int)(sizeof(FrameRec) / 4); i += 32) dst[i] = keep ? src[i] : 0u;
  }
}

// ------------------------------------------------------------------------------------------------------------
// K5: Huffman table-15 bit packing (SRC:1705-1737, writer semantics SRC:2230-2252).  One warp per frame, four frames per
// CTA; the warp walks the frame's 2 ch granule-channels.  Lane L codes pairs 9L...9L+8, so a warp prefix sum of the lane
// bit counts gives every lane its bit position; codes are OR-ed MSB-first into the warp's own bit buffer and the frame's
// bytes are written once.  Nothing in the frame loop waits for another warp (the previous layout — a warp per
// granule-channel, a CTA walking four frames — paid three block barriers per frame: mono 1.14 -> 0.97 ms, stereo 1.62 -> 1.60).
constexpr int kPackFramesPerCta = 4;
__device__ __forceinline__ void put_bits64(uint32_t *buf, uint32_t pos, unsigned long long v, int len) {   // len in 1...64, MSB first
  const unsigned long long top = v << (64 - len);
  const uint32_t w = pos >> 5, off = pos & 31;
  const uint32_t a = (uint32_t)(top >> (32 + off)), b = (uint32_t)(top >> off), c = off ? (uint32_t)(top << (32 - off)) : 0u;
  if (a) atomicOr(&buf[w], a);
  if (b) atomicOr(&buf[w + 1], b);
  if (c) atomicOr(&buf[w + 2], c);
}
struct PackGc { float2 v[9]; uint32_t sel, bitoff; };
__device__ __forceinline__ void pack_load(const PassBuffers &pb, size_t gslot, int lane, PackGc &g) {
  const float2 *sm2 = reinterpret_cast<const float2 *>(pb.smag + gslot * 576) + 9 * lane;
#pragma unroll
  for (int j = 0; j < 9; ++j) g.v[j] = __ldg(sm2 + j);
  g.sel = pb.gc_sel[gslot]; g.bitoff = pb.gc_bitoff[gslot];
}
template <bool TRACE> __global__ void __launch_bounds__(32 * kPackFramesPerCta, TRACE ? 4 : 8) k_pack(Config cfg, PassBuffers pb) {   // TRACE: also leave ix behind
  __shared__ uint32_t bufs[kPackFramesPerCta][548];
  __shared__ __align__(16) uint16_t tab31[31 * 32];   // code | length << 8, indexed by quant30
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ch = cfg.channels, ngc = 2 * ch;
  const int f = blockIdx.y * kPackFramesPerCta + warp;
  const bool active = f < (int)pb.plan[s].n_frames;
  PackGc cur;
  const size_t gslot0 = (size_t)s * pb.GC + (size_t)(active ? f : 0) * ngc;
  if (active) pack_load(pb, gslot0, lane, cur);    // in flight while the tables are staged
  if (tid < 31 * 32 / 8) reinterpret_cast<uint4 *>(tab31)[tid] = reinterpret_cast<const uint4 *>(tab::kTab31)[tid];   // global, coalesced: a lane-indexed constant-bank read would serialise
  uint32_t *buf = bufs[warp];
  for (int i = lane; i < 548; i += 32) buf[i] = 0;
  __syncthreads();
  if (!active) return;
  for (int g = 0; g < ngc; ++g) {
    const size_t gslot = gslot0 + g;
    const int gain = cur.sel & 255, bv = cur.sel >> 8;
    const float inv = __fmul_rn(c_inv_step[gain], 2.0f);
    int32_t *trix = TRACE ? pb.tr_ix + gslot * 576 : nullptr;
    uint32_t val[9]; int len[9]; int mine = 0;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const int p = 9 * lane + j;
      const float2 v = cur.v[j];
      const int qx = quant30(fabsf(v.x), inv), qy = quant30(fabsf(v.y), inv);     // (u, not q: non-zero exactly when q is; the conversion-free form of k_granule measured slower here: 12.5 vs 12.3 ms)
      if (TRACE) { const int ax = (qx + 1) >> 1, ay = (qy + 1) >> 1; trix[2 * p] = v.x < 0.0f ? -ax : ax; trix[2 * p + 1] = v.y < 0.0f ? -ay : ay; }
      const uint32_t t15 = tab31[qx * 32 + qy];
      uint32_t code = t15 & 255u; int l = (int)(t15 >> 8);
      if (qx) { code = code << 1 | __float_as_uint(v.x) >> 31; ++l; }   // SRC:1729-1736 (the magnitude carries the line's sign)
      if (qy) { code = code << 1 | __float_as_uint(v.y) >> 31; ++l; }
      if (p >= bv) { l = 0; code = 0; }
      val[j] = code; len[j] = l; mine += l;
    }
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    uint32_t pos = cur.bitoff + (uint32_t)(incl - mine);
    // a lane's nine codes are consecutive in the stream: they are concatenated in registers (a code is at most 13 + 2 bits, four
    // of them fit 64) and leave as three pieces instead of nine — each piece costs a dozen instructions and up to three atomics
#pragma unroll
    for (int j0 = 0; j0 < 9; j0 += 4) {
      unsigned long long acc = 0; int al = 0;
#pragma unroll
      for (int j = j0; j < (j0 + 4 < 9 ? j0 + 4 : 9); ++j) { acc = acc << len[j] | val[j]; al += len[j]; }
      if (al) put_bits64(buf, pos, acc, al);
      pos += al;
    }
    if (g + 1 < ngc) pack_load(pb, gslot + 1, lane, cur);   // (holding the next one's magnitudes during the coding above costs more in occupancy than it hides: measured)
  }
  __syncwarp();
  const uint32_t off = pb.fr_md[((size_t)s * pb.Fc + f) * 2], nbytes = pb.fr_md[((size_t)s * pb.Fc + f) * 2 + 1];
  uint8_t *dst = pb.md + (size_t)s * pb.md_stride + off;
  if (off + nbytes <= pb.md_stride)
    for (uint32_t i = lane; i < nbytes; i += 32) dst[i] = (uint8_t)(buf[i >> 2] >> (24 - 8 * (i & 3)));
}

// ------------------------------------------------------------------------------------------------------------
// K5 in ISO mode (north_star stage 5; supersedes the reference's dead HuffmanEncoder.encode / writePair / selectTable,
// SRC:1740-1806): the frame's granule-channels are quantized with the ISO law at the gain the scan chose, partitioned and
// table-selected by the same iso_evaluate the curve used (so the bit count is the one the scan budgeted — checked), then coded:
// big_values pairs with the region's table (+ linbits escapes, + sign bits), count1 quadruples with table A or B.  One warp per
// frame; lane L codes pairs 9L...9L+8 and quadruples 5L...5L+4, bit positions from warp prefix sums.  The choices (regions,
// table_select, count1table_select, big_values) go into the frame record for k_frames' side info.
template <bool TRACE> __global__ void __launch_bounds__(32 * kPackFramesPerCta) k_pack_iso(Config cfg, PassBuffers pb) {
  __shared__ uint32_t bufs[kPackFramesPerCta][552];
  __shared__ uint32_t s_huff[kHuffEntries];
  __shared__ __align__(16) uint8_t s_len[(kHuffEntries + 15) / 16 * 16];
  __shared__ int16_t s_ix[kPackFramesPerCta][576];
  __shared__ uint8_t s_c[kPackFramesPerCta][288];
  __shared__ uint8_t s_band[288];                                  // scalefactor band of pair p (level 2)
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ch = cfg.channels, ngc = 2 * ch;
  const int f = blockIdx.y * kPackFramesPerCta + warp;
  const int nfr = (int)pb.plan[s].n_frames;
  const bool active = f < nfr;
  const bool lv2 = cfg.iso >= 2;
  for (int i = tid; i < kHuffEntries; i += 32 * kPackFramesPerCta) { s_huff[i] = kHuffPacked[i]; s_len[i] = kHuffLenFlat[i]; }
  for (int i = tid; i < 288; i += 32 * kPackFramesPerCta) {
    int b = 0;
    for (int k = 0; k < 21; ++k) b += c_sfb_cum[cfg.sfb_index][k] <= 2 * i;
    s_band[i] = (uint8_t)b;
  }
  uint32_t *buf = bufs[warp];
  for (int i = lane; i < 552; i += 32) buf[i] = 0;
  __syncthreads();
  if (!active) return;
  const int *sfb = c_sfb_cum[cfg.sfb_index];
  FrameRec &fr = pb.rec[(size_t)s * (pb.Fc + 1) + 1 + f];
  for (int g = 0; g < ngc; ++g) {
    const size_t gslot = (size_t)s * pb.GC + (size_t)f * ngc + g;
    const uint32_t sel = pb.gc_sel[gslot], bitoff = pb.gc_bitoff[gslot];
    const float inv = c_inv_step_iso[sel & 511u];
    const float2 *sm2 = reinterpret_cast<const float2 *>(pb.smag + gslot * 576);
    // level 2: the scalefactors k_outer chose amplify |xr|^0.75 band by band (sf[21] = 0: the last band has none)
    const uint8_t *sfp = lv2 ? pb.gc_sf + gslot * 24 : nullptr;
    const int part2 = lv2 ? sfp[22] : 0, sfc = lv2 ? sfp[21] : 0;
    int qx[9], qy[9];
    int16_t *ix = s_ix[warp];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const int p = lane + 32 * j;
      const float2 v = __ldg(sm2 + p);
      float m0 = fabsf(v.x), m1 = fabsf(v.y);
      if (lv2) { const int b = s_band[p]; const float a = c_amp34[b < 21 ? sfp[b] : 0]; m0 = __fmul_rn(m0, a); m1 = __fmul_rn(m1, a); }
      qx[j] = iso_quant(m0, inv); qy[j] = iso_quant(m1, inv);
      ix[2 * p] = (int16_t)(v.x < 0.0f ? -qx[j] : qx[j]); ix[2 * p + 1] = (int16_t)(v.y < 0.0f ? -qy[j] : qy[j]);
      if (TRACE) { pb.tr_ix[gslot * 576 + 2 * p] = ix[2 * p]; pb.tr_ix[gslot * 576 + 2 * p + 1] = ix[2 * p + 1]; }
    }
    const IsoChoice c = iso_evaluate(qx, qy, lane, s_len, s_c[warp], sfb, (pb.gc_bt[gslot] & 3) != 0);
    __syncwarp();
    if (lane == 0) {
      GcSide &gs = fr.gc[g];
      if (c.bits + part2 != (int)gs.part23) atomicOr(&pb.state[s].error, 16);  // the curve and the packer must agree bit for bit
      gs.sfc = (uint8_t)sfc; gs.part2 = (uint8_t)part2;
      gs.big_values = (uint16_t)c.bv; gs.region0 = (uint8_t)c.r0; gs.region1 = (uint8_t)c.r1;
      gs.tsel[0] = (uint8_t)c.tsel[0]; gs.tsel[1] = (uint8_t)c.tsel[1]; gs.tsel[2] = (uint8_t)c.tsel[2]; gs.c1sel = (uint8_t)c.c1sel;
      if (f == nfr - 1 && pb.state[s].buffered.valid) {             // the scan has already parked this frame as bufferedFrame
        GcSide &bs = pb.state[s].buffered.gc[g];
        bs.big_values = gs.big_values; bs.region0 = gs.region0; bs.region1 = gs.region1;
        bs.tsel[0] = gs.tsel[0]; bs.tsel[1] = gs.tsel[1]; bs.tsel[2] = gs.tsel[2]; bs.c1sel = gs.c1sel;
        bs.sfc = gs.sfc; bs.part2 = gs.part2;
      }
    }
    if (part2 && lane < 21) {                                       // part 2: 11 scalefactors of slen1 bits, 10 of slen2 (scfsi = 0)
      const int l1 = c_slen1[sfc], l2 = c_slen2[sfc];
      const int len = lane < 11 ? l1 : l2;
      if (len) put_bits64(buf, bitoff + (uint32_t)(lane < 11 ? lane * l1 : 11 * l1 + (lane - 11) * l2), (unsigned long long)sfp[lane], len);
    }
    // ---- big_values: lane L codes pairs 9L ... 9L+8
    const uint32_t d0 = iso_desc(c.tsel[0]), d1 = iso_desc(c.tsel[1]), d2 = iso_desc(c.tsel[2]);
    unsigned long long val[9]; int len[9]; int mine = 0;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const int p = 9 * lane + j;
      val[j] = 0ull; len[j] = 0;
      if (p < c.bv) {
        const int x = ix[2 * p], y = ix[2 * p + 1], ax = abs(x), ay = abs(y);
        const uint32_t dk = 2 * p < c.a1 ? d0 : 2 * p < c.a2 ? d1 : d2;
        const int dim = (int)((dk >> 16) & 255u), lb = (int)((dk >> 24) & 255u);
        if (dim) {                                                  // (table 0: the region is all zero, nothing is coded)
          const uint32_t e = s_huff[(dk & 0xFFFFu) + min(ax, 15) * dim + min(ay, 15)];
          unsigned long long w = e & 0xFFFFFFu; int l = (int)(e >> 24);
          if (lb && ax >= 15) { w = w << lb | (unsigned long long)(ax - 15); l += lb; }
          if (ax) { w = w << 1 | (unsigned long long)(x < 0); ++l; }
          if (lb && ay >= 15) { w = w << lb | (unsigned long long)(ay - 15); l += lb; }
          if (ay) { w = w << 1 | (unsigned long long)(y < 0); ++l; }
          val[j] = w; len[j] = l;
        }
      }
      mine += len[j];
    }
    int incl = mine;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, dlt); if (lane >= dlt) incl += t; }
    const int big_total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t pos = bitoff + (uint32_t)part2 + (uint32_t)(incl - mine);
#pragma unroll
    for (int j = 0; j < 9; ++j) if (len[j]) { put_bits64(buf, pos, val[j], len[j]); pos += len[j]; }
    // ---- count1: lane L codes quadruples 5L ... 5L+4 (at most 144 of them)
    uint32_t qv[5]; int ql[5]; int qmine = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int k = 5 * lane + i;
      qv[i] = 0; ql[i] = 0;
      if (k < c.c1) {
        const int16_t *q4 = ix + 2 * c.bv + 4 * k;
        const int idx = (q4[0] != 0) << 3 | (q4[1] != 0) << 2 | (q4[2] != 0) << 1 | (q4[3] != 0);
        uint32_t w = c.c1sel ? (uint32_t)(15 - idx) : (uint32_t)((kQuadCodeAPacked >> (4 * idx)) & 15ull);
        int l = c.c1sel ? 4 : quad_len_a(idx);
#pragma unroll
        for (int m = 0; m < 4; ++m) if (q4[m]) { w = w << 1 | (uint32_t)(q4[m] < 0); ++l; }
        qv[i] = w; ql[i] = l;
      }
      qmine += ql[i];
    }
    int qincl = qmine;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) { int t = __shfl_up_sync(0xffffffffu, qincl, dlt); if (lane >= dlt) qincl += t; }
    pos = bitoff + (uint32_t)part2 + (uint32_t)big_total + (uint32_t)(qincl - qmine);
#pragma unroll
    for (int i = 0; i < 5; ++i) if (ql[i]) { put_bits64(buf, pos, qv[i], ql[i]); pos += ql[i]; }
    __syncwarp();
  }
  __syncwarp();
  const uint32_t off = pb.fr_md[((size_t)s * pb.Fc + f) * 2], nbytes = pb.fr_md[((size_t)s * pb.Fc + f) * 2 + 1];
  uint8_t *dst = pb.md + (size_t)s * pb.md_stride + off;
  if (off + nbytes <= pb.md_stride && nbytes <= 552 * 4 - 8)
    for (uint32_t i = lane; i < nbytes; i += 32) dst[i] = (uint8_t)(buf[i >> 2] >> (24 - 8 * (i & 3)));
}

// ------------------------------------------------------------------------------------------------------------
// K_frames: header (SRC:522-544) + side info (SRC:571-625) + slot fill (SRC:2110-2121).  One warp per frame slot.
struct BitW {   // MSB-first writer into a byte array (BitstreamWriter SRC:2230-2252), single thread
  uint8_t *p; int n; uint32_t acc; int nb;
  __device__ void put(uint32_t bits, int count) {
    acc = (acc << count) | (bits & ((1u << count) - 1u)); nb += count;
    while (nb >= 8) { nb -= 8; p[n++] = (uint8_t)(acc >> nb); }
  }
  __device__ void pad() { if (nb > 0) { p[n++] = (uint8_t)(acc << (8 - nb)); nb = 0; } acc = 0; }
};

__device__ inline uint16_t crc16_mpeg(const uint8_t *p, int n) {  // SRC:2190-2215 (poly 0x8005, init 0xFFFF)
  uint32_t crc = 0xFFFF;
  for (int i = 0; i < n; ++i) {
    crc ^= (uint32_t)p[i] << 8;
    for (int b = 0; b < 8; ++b) crc = (crc & 0x8000) ? ((crc << 1) ^ 0x8005) & 0xFFFF : (crc << 1) & 0xFFFF;
  }
  return (uint16_t)crc;
}

__global__ void __launch_bounds__(128) k_frames(Config cfg, PassBuffers pb) {
  __shared__ uint32_t hwords[4][12];               // per warp: header + CRC + side info as big-endian words (<= 38 bytes)
  const int s = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.y * 4 + warp;
  if (r > (int)pb.plan[s].n_frames) return;
  const FrameRec &fr = pb.rec[(size_t)s * (pb.Fc + 1) + r];
  const FrameEmit em = pb.emit[(size_t)s * (pb.Fc + 1) + r];
  if (!fr.valid || !em.emit) return;
  const int ch = cfg.channels;
  uint32_t *hw = hwords[warp];
  if (lane < 12) hw[lane] = 0;
  __syncwarp();
  // The bit layout is fixed, so the fields are assembled in parallel: lane 4 the header (+ CRC), lane 5 main_data_begin /
  // private bits / scfsi, lanes 0..2ch-1 one 59-bit granule record each; every lane ORs its bits into the word buffer.
  {
    unsigned long long v = 0; int c = 0, o = 0;
    const int crc_bits = cfg.crc ? 16 : 0, prefix_bits = 9 + (ch == 1 ? 5 : 3) + 4 * ch;
    if (lane < 2 * ch) {                                                              // buildSideInfo SRC:586-624
      const GcSide &g = fr.gc[lane];
      const int ws = g.block_type != 0;
      v = (unsigned long long)(g.part23 & 0xFFF) << 47 | (unsigned long long)(g.big_values & 0x1FF) << 38 |
          (unsigned long long)g.global_gain << 30 | (unsigned long long)(g.sfc & 15) << 26 | (unsigned long long)ws << 25;   // scalefac_compress = 0 outside ISO mode level 2 (SRC:722)
      unsigned long long mid;                                                         // 22 bits at 3
      if (ws && cfg.iso) mid = (unsigned long long)(g.block_type & 3) << 20 | (unsigned long long)(g.tsel[0] & 31) << 14 | (unsigned long long)(g.tsel[1] & 31) << 9;   // ISO mode level 3: mixed_block_flag = 0, subblock_gain = 0
      else if (ws) mid = (unsigned long long)(g.block_type & 3) << 20 | (unsigned long long)(g.block_type == 1) << 19 | 15ull << 14 | 15ull << 9 |
                    (unsigned long long)(g.sbg[0] & 7) << 6 | (unsigned long long)(g.sbg[1] & 7) << 3 | (unsigned long long)(g.sbg[2] & 7);
      else mid = (unsigned long long)(g.tsel[0] & 31) << 17 | (unsigned long long)(g.tsel[1] & 31) << 12 | (unsigned long long)(g.tsel[2] & 31) << 7 |
                 (unsigned long long)(g.region0 & 15) << 3 | (unsigned long long)(g.region1 & 7);   // table_select = 15, 15, 15 outside ISO mode (SRC:717)
      v |= mid << 3 | (unsigned long long)(g.preflag & 1) << 2 | (unsigned long long)(g.c1sel & 1);   // scalefac_scale = 0; count1table_select = 0 outside ISO mode
      c = 59; o = 32 + crc_bits + prefix_bits + 59 * lane;
    } else if (lane == 4) {                                                           // header SRC:522-536, CRC SRC:538-543
      uint32_t h = 0x7FFu;
      h = h << 2 | 3u; h = h << 2 | 1u; h = h << 1 | (cfg.crc ? 0u : 1u); h = h << 4 | (fr.br_index & 15u); h = h << 2 | (uint32_t)cfg.sr_index;
      // mode_extension: the reference writes 0b10 on every joint-stereo frame; ISO mode signals M/S per frame
      const uint32_t mode_ext = cfg.iso ? (cfg.mode == 2 && fr.ms ? 2u : 0u) : (uint32_t)cfg.mode_ext;
      h = h << 1 | (fr.padding & 1u); h = h << 1; h = h << 2 | (uint32_t)cfg.mode_bits; h = h << 2 | mode_ext;
      h = h << 1 | (cfg.copyright ? 1u : 0u); h = h << 1 | (cfg.original ? 1u : 0u); h = h << 2;
      v = h; c = 32; o = 0;
      if (cfg.crc) {
        const uint8_t hb[4] = {(uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h};
        v = (unsigned long long)h << 16 | crc16_mpeg(hb, 4); c = 48;
      }
    } else if (lane == 5) {                                                           // SRC:577-584: mdb, private bits, scfsi = 0
      v = (unsigned long long)min((int)fr.mdb, 511) << (prefix_bits - 9); c = prefix_bits; o = 32 + crc_bits;
    }
    if (c) {
      const unsigned long long hi = v << (64 - c);                                    // left aligned
      const int w = o >> 5, sft = o & 31;
      const uint32_t a = (uint32_t)(hi >> (32 + sft)), b2 = (uint32_t)((hi << (32 - sft)) >> 32);
      const uint32_t d = sft ? (uint32_t)((hi << (64 - sft)) >> 32) : 0u;
      if (a) atomicOr(&hw[w], a);
      if (b2) atomicOr(&hw[w + 1], b2);
      if (d) atomicOr(&hw[w + 2], d);
    }
  }
  __syncwarp();
  if (cfg.iso && cfg.crc && lane == 0) {                         // ISO 11172-3 CRC: header bytes 2-3 and the side info (the reference: the 4 header bytes)
    uint8_t bytes[2 + 32];
    bytes[0] = (uint8_t)(hw[0] >> 8); bytes[1] = (uint8_t)hw[0];
    for (int i = 0; i < cfg.side_bytes; ++i) { const int k = 6 + i; bytes[2 + i] = (uint8_t)(hw[k >> 2] >> (24 - 8 * (k & 3))); }
    const uint32_t crc = crc16_mpeg(bytes, 2 + cfg.side_bytes);
    hw[1] = (hw[1] & 0x0000FFFFu) | crc << 16;
  }
  __syncwarp();
  uint8_t *dst = pb.out + (size_t)s * pb.out_stride + em.out_off;
  if ((size_t)em.out_off + cfg.header_bytes + fr.slot > pb.out_stride) return;
  for (int i = lane; i < cfg.header_bytes; i += 32) dst[i] = (uint8_t)(hw[i >> 2] >> (24 - 8 * (i & 3)));
  dst += cfg.header_bytes;
  const uint32_t B0 = pb.md_tail[(size_t)s * 4 + 2];
  const uint8_t *carry = pb.md_carry + (size_t)s * kMdCarryCap, *md = pb.md + (size_t)s * pb.md_stride;
  // fillSlot: `take` bytes from the FIFO, zero padding up to the slot size.  When the bytes all come from this pass's part
  // of the FIFO, the middle of the slot is copied as aligned words (two aligned source words funnel-shifted per word).
  auto byte_at = [&](uint32_t i) -> uint8_t {
    if (i >= em.take) return 0;
    const uint32_t o = em.src_off + i;
    return o < B0 ? carry[o] : md[o];
  };
  const uint32_t slot = fr.slot;
  if (em.src_off >= B0 || em.take == 0) {
    const uint32_t head = min((4u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u, slot), words = (slot - head) >> 2;
    if (lane < head) dst[lane] = byte_at(lane);
    for (uint32_t i = head + 4 * words + lane; i < slot; i += 32) dst[i] = byte_at(i);
    uint32_t *dw = reinterpret_cast<uint32_t *>(dst + head);
    const uint32_t *mw = reinterpret_cast<const uint32_t *>(md);
    for (uint32_t k = lane; k < words; k += 32) {
      const uint32_t i0 = head + 4 * k;
      uint32_t w = 0;
      if (i0 < em.take) {
        const uint32_t o = em.src_off + i0;
        w = __funnelshift_r(mw[o >> 2], mw[(o >> 2) + 1], 8 * (o & 3));
        const uint32_t left = em.take - i0;                       // bytes of this word that exist
        if (left < 4) w &= (1u << (8 * left)) - 1u;
      }
      dw[k] = w;
    }
  } else {
    for (uint32_t i = lane; i < slot; i += 32) dst[i] = byte_at(i);
  }
}

// ------------------------------------------------------------------------------------------------------------
// K_carry: state that crosses the pass boundary: last frame + pending partial PCM, reservoir backlog bytes, VBR
// history, stereo decision of the carried frame.  One CTA per stream.
// The pass boundary in two halves, so that the tail of a pass (scan, pack, frames) can run beside the head of the next one
// (filterbank, granule) on a second stream: what the NEXT HEAD needs is known once k_granule is done — the carried PCM frame and
// pending samples, the MDCT overlap rows, the VBR energy history and the stereo decision of the last frame (k_carry_head) —,
// what the next SCAN needs comes out of this pass's scan — the main-data backlog (k_carry_tail).
__global__ void __launch_bounds__(256) k_carry_head(Config cfg, PassBuffers pb) {
  const int s = blockIdx.x, tid = threadIdx.x;
  const StreamPlan plan = pb.plan[s];
  const PcmView pv = pcm_view(cfg, pb, s);
  const int64_t total = (int64_t)plan.head_n + plan.cur_n;
  const int64_t start = (int64_t)plan.n_frames * cfg.fsc;
  int64_t left = total - start - cfg.fsc; if (left < 0) left = 0;      // pending partial after the carried frame
  const int keep = cfg.fsc + (int)left;
  float *dst = pb.head_out + (size_t)s * 2 * cfg.fsc;
  for (int i = tid; i < keep && i < 2 * cfg.fsc; i += 256) dst[i] = pv.at(start + i);
  // MDCT overlap (SRC:1534-1535): the last granule's subband rows become rows 0..17 of the next pass
  if (plan.n_frames) {
    const int ngr = 2 * (int)plan.n_frames;
    for (int c = 0; c < cfg.channels; ++c) {
      float *base = pb.sub + (size_t)(s * cfg.channels + c) * pb.sub_rows * 32;
      for (int i = tid; i < 576; i += 256) base[i] = base[(size_t)18 * ngr * 32 + i];
    }
  }
  StreamState &st = pb.state[s];
  const int ngc = (int)plan.n_frames * 2 * cfg.channels;
  if (tid == 0 && plan.n_frames) {
    const float *hist = pb.gc_energy + (size_t)s * (10 + pb.GC);
    int count = min(10, st.vbr_n + ngc);
    float h[10];
    for (int i = 0; i < count; ++i) h[i] = hist[10 + ngc - count + i];
    for (int i = 0; i < count; ++i) st.vbr_hist[i] = h[i];
    st.vbr_n = count;
    st.ms_prev = pb.ms[(size_t)s * (pb.Fc + 1) + plan.n_frames];
  }
}
__global__ void __launch_bounds__(256) k_carry_tail(Config cfg, PassBuffers pb) {
  __shared__ uint8_t tail[kMdCarryCap];
  const int s = blockIdx.x, tid = threadIdx.x;
  const uint32_t R = pb.md_tail[(size_t)s * 4], len = pb.md_tail[(size_t)s * 4 + 1], B0 = pb.md_tail[(size_t)s * 4 + 2];
  uint8_t *carry = pb.md_carry + (size_t)s * kMdCarryCap;
  const uint8_t *md = pb.md + (size_t)s * pb.md_stride;
  for (uint32_t i = tid; i < len; i += 256) { uint32_t o = R + i; tail[i] = o < B0 ? carry[o] : md[o]; }
  __syncthreads();
  for (uint32_t i = tid; i < len; i += 256) carry[i] = tail[i];
}

// ------------------------------------------------------------------------------------------------------------
// PsychoacousticModel.maskingThresholds SRC:1983-2013 — dead output in the reference (SRC:737); trace only.
__global__ void __launch_bounds__(256) k_thresholds(Config cfg, PassBuffers pb) {
  const int s = blockIdx.x, lane = threadIdx.x & 31;
  const int gci = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (gci >= (int)pb.plan[s].n_frames * 2 * cfg.channels || !pb.tr_thr) return;
  const size_t gslot = (size_t)s * pb.GC + gci;
  const float *x = pb.spec + gslot * 576;
  float *thr = pb.tr_thr + gslot * 576;
  double qd = (double)(10 - cfg.quality) / 10.0;
  const float quality_scale = (float)(qd > 0.1 ? qd : 0.1);
  for (int i = lane; i < 576; i += 32) thr[i] = 0.0001f;
  __syncwarp();
  int cursor = 0;
  for (int b = 0; b < 21; ++b) {
    int end = min(c_sfb_cum[cfg.sfb_index][b], 576), size = end - cursor;
    if (size > 0) {
      float p = 0.0f;
      for (int i = lane; i < size; i += 32) { float v = x[cursor + i]; p = __fmaf_rn(v, v, p); }
      float t = fmaxf(__fmul_rn(__fdiv_rn(lane_tree(p), (float)size), quality_scale), 0.0001f);
      for (int i = lane; i < size; i += 32) thr[cursor + i] = t;
    }
    cursor = end;
    if (cursor >= 576) break;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Output compaction: offsets[s] = 16-byte aligned exclusive prefix sum of the streams' output lengths.
__global__ void __launch_bounds__(1024) k_offsets(Config cfg, PassBuffers pb, uint64_t *offsets) {
  __shared__ uint64_t part[1024];
  const int tid = threadIdx.x, S = cfg.n_streams;
  const int per = (S + 1023) / 1024, b = tid * per, e = min(b + per, S);
  uint64_t sum = 0;
  for (int i = b; i < e; ++i) sum += ((uint64_t)pb.state[i].out_pos + 15) & ~15ull;
  part[tid] = sum;
  __syncthreads();
  if (tid == 0) { uint64_t run = 0; for (int i = 0; i < 1024; ++i) { uint64_t t = part[i]; part[i] = run; run += t; } offsets[S] = run; }
  __syncthreads();
  uint64_t run = part[tid];
  for (int i = b; i < e; ++i) { offsets[i] = run; run += ((uint64_t)pb.state[i].out_pos + 15) & ~15ull; }
}
__global__ void __launch_bounds__(256) k_gather(Config cfg, PassBuffers pb, const uint64_t *offsets, uint8_t *compact) {
  const int s = blockIdx.x;
  const uint32_t len = pb.state[s].out_pos, n16 = (len + 15) >> 4;
  const uint4 *src = reinterpret_cast<const uint4 *>(pb.out + (size_t)s * pb.out_stride);
  uint4 *dst = reinterpret_cast<uint4 *>(compact + offsets[s]);
  for (uint32_t i = blockIdx.y * 256 + threadIdx.x; i < n16; i += gridDim.y * 256) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------------------
// Synthetic PCM (bench / tests), BASELINE C1/C4 recipe: a*sin(2 pi f t) + noise*N(0,1), clipped to [-1, 1].
// Counter-based and bit-reproducible on any IEEE-754 machine: every operation is an explicitly rounded double operation
// (no libm, no fast-math intrinsics), so the CPU twin in oracle/mp3_oracle.c (orc_synth_fill) produces the same floats and
// the CPU arm of the bench encodes exactly the inputs the GPU arm does (tests/test_gpu_parity.py::test_synth_twin).
//   h1 = splitmix64(seed * 0x100000001B3 + i), h2 = splitmix64(h1)
//   u1 = ((h1 >> 11) + 1) / 2^53 in (0, 1], u2 = (h2 >> 11) / 2^53 in [0, 1)
//   Box-Muller: rad = sqrt(-2 ln u1), (gL, gR) = rad * (cos, sin)(2 pi u2)
//   tone_c = sin(2 pi * frac(frac(f_c * i / sr) + phase_c)), phase_L = 0, phase_R = 0.3 / (2 pi)
//   sample = clip((float)(amp * tone_c + noise * g_c))
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// ln(k * 2^-53), k in [1, 2^53]: k = m * 2^e with m in [sqrt(1/2), sqrt(2)), ln m = 2 atanh((m - 1) / (m + 1)) as an 11-term
// odd series (|s| <= 0.172: the first dropped term is below 1e-18)
__device__ __forceinline__ double synth_ln_u(uint64_t k) {
  int e = 63 - __clzll((long long)k);                              // k = 2^e * [1, 2)
  double m = e == 53 ? 1.0 : __longlong_as_double((long long)((k << ((52 - e) & 63)) & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ll);   // exact: k has at most 53 bits
  if (m > 1.4142135623730951) { m = __dmul_rn(m, 0.5); e += 1; }
  const double sq = __ddiv_rn(__dadd_rn(m, -1.0), __dadd_rn(m, 1.0)), z = __dmul_rn(sq, sq);
  double p = 1.0 / 21.0;
  p = __fma_rn(p, z, 1.0 / 19.0); p = __fma_rn(p, z, 1.0 / 17.0); p = __fma_rn(p, z, 1.0 / 15.0); p = __fma_rn(p, z, 1.0 / 13.0);
  p = __fma_rn(p, z, 1.0 / 11.0); p = __fma_rn(p, z, 1.0 / 9.0); p = __fma_rn(p, z, 1.0 / 7.0); p = __fma_rn(p, z, 1.0 / 5.0);
  p = __fma_rn(p, z, 1.0 / 3.0); p = __fma_rn(p, z, 1.0);
  return __fma_rn((double)(e - 53), 0.6931471805599453, __dmul_rn(__dmul_rn(2.0, sq), p));
}
// (sin, cos)(2 pi t), t in [0, 1): quadrant from t * 4 (exact), Taylor polynomials on [0, pi / 2)
__device__ __forceinline__ void synth_sincos_turn(double t, double &sn, double &cs) {
  const double q4 = floor(__dmul_rn(t, 4.0));
  const double x = __dmul_rn(6.283185307179586, __dadd_rn(t, -__dmul_rn(q4, 0.25))), z = __dmul_rn(x, x);
  double ps = -1.0 / 25852016738884976640000.0;                    // -1/23!
  ps = __fma_rn(ps, z, 1.0 / 51090942171709440000.0); ps = __fma_rn(ps, z, -1.0 / 121645100408832000.0); ps = __fma_rn(ps, z, 1.0 / 355687428096000.0);
  ps = __fma_rn(ps, z, -1.0 / 1307674368000.0); ps = __fma_rn(ps, z, 1.0 / 6227020800.0); ps = __fma_rn(ps, z, -1.0 / 39916800.0);
  ps = __fma_rn(ps, z, 1.0 / 362880.0); ps = __fma_rn(ps, z, -1.0 / 5040.0); ps = __fma_rn(ps, z, 1.0 / 120.0);
  ps = __fma_rn(ps, z, -1.0 / 6.0); ps = __fma_rn(ps, z, 1.0);
  const double s0 = __dmul_rn(x, ps);
  double pc = 1.0 / 620448401733239439360000.0;                    // 1/24!
  pc = __fma_rn(pc, z, -1.0 / 1124000727777607680000.0); pc = __fma_rn(pc, z, 1.0 / 2432902008176640000.0); pc = __fma_rn(pc, z, -1.0 / 6402373705728000.0);
  pc = __fma_rn(pc, z, 1.0 / 20922789888000.0); pc = __fma_rn(pc, z, -1.0 / 87178291200.0); pc = __fma_rn(pc, z, 1.0 / 479001600.0);
  pc = __fma_rn(pc, z, -1.0 / 3628800.0); pc = __fma_rn(pc, z, 1.0 / 40320.0); pc = __fma_rn(pc, z, -1.0 / 720.0);
  pc = __fma_rn(pc, z, 1.0 / 24.0); pc = __fma_rn(pc, z, -0.5); pc = __fma_rn(pc, z, 1.0);
  const int q = (int)q4;
  sn = q == 0 ? s0 : q == 1 ? pc : q == 2 ? -s0 : -pc;
  cs = q == 0 ? pc : q == 1 ? -s0 : q == 2 ? -pc : s0;
}
__global__ void k_synth(float *pcm, size_t n, int channels, int sample_rate, float f_left, float f_right, float amp,
                        float noise, uint64_t seed) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t h1 = splitmix64(seed * 0x100000001B3ull + i), h2 = splitmix64(h1);
  const double rad = __dsqrt_rn(__dmul_rn(-2.0, synth_ln_u((h1 >> 11) + 1)));
  double gs, gc;
  synth_sincos_turn(__dmul_rn((double)(h2 >> 11), 1.1102230246251565e-16), gs, gc);   // * 2^-53, exact
  for (int c = 0; c < channels; ++c) {
    const double cyc = __ddiv_rn(__dmul_rn((double)(c == 0 ? f_left : f_right), (double)i), (double)sample_rate);
    double t = __dadd_rn(cyc, -floor(cyc));
    if (c == 1) { t = __dadd_rn(t, 0.0477464829275686); if (t >= 1.0) t = __dadd_rn(t, -1.0); }   // + 0.3 rad
    double sn, cs;
    synth_sincos_turn(t, sn, cs);
    const double v = __dadd_rn(__dmul_rn((double)amp, sn), __dmul_rn((double)noise, __dmul_rn(rad, c == 0 ? gc : gs)));
    pcm[i * channels + c] = fminf(fmaxf(__double2float_rn(v), -1.0f), 1.0f);
  }
}

// ------------------------------------------------------------------------------------------------------------
// 16-bit PCM input (mp3b_batch_encode_i16): sample = Float(s) / 32768, exact in FP32.  Runs on the copy stream right after
// the upload of a pass, so the PCIe traffic of 16-bit sources is halved and nothing else changes.
__global__ void __launch_bounds__(256) k_widen_i16(const int16_t *in, float *out, size_t stride, const StreamPlan *plan) {
  const int s = blockIdx.y;
  const uint32_t n = plan[s].cur_n;
  const short4 *in4 = reinterpret_cast<const short4 *>(in + (size_t)s * stride);
  float4 *out4 = reinterpret_cast<float4 *>(out + (size_t)s * stride);
  const float k = 1.0f / 32768.0f;
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n / 4; i += gridDim.x * 256) {
    const short4 v = in4[i];
    out4[i] = make_float4(__fmul_rn((float)v.x, k), __fmul_rn((float)v.y, k), __fmul_rn((float)v.z, k), __fmul_rn((float)v.w, k));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3u)) {
    const size_t i = (size_t)s * stride + (n & ~3u) + threadIdx.x;
    out[i] = __fmul_rn((float)in[i], k);
  }
}
int launch_widen_i16(const int16_t *in, float *out, size_t stride, const StreamPlan *d_plan, int n_streams, cudaStream_t st) {
  dim3 grid(64, n_streams);
  k_widen_i16<<<grid, 256, 0, st>>>(in, out, stride, d_plan);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -(int)e;
}

// ------------------------------------------------------------------------------------------------------------
// Self-test (tests only): pow34 against its definition on every finite float >= 1e-10, div_exact against the IEEE
// division for every finite float.  mismatch[0] / [1] / [2] count pow34, /9 and /3 disagreements.
__global__ void k_selftest(unsigned long long *mismatch) {
  const uint32_t stride = gridDim.x * blockDim.x;
  unsigned long long bad0 = 0, bad1 = 0, bad2 = 0;
  for (uint64_t u = blockIdx.x * blockDim.x + threadIdx.x; u < (1ull << 32); u += stride) {
    const float a = __uint_as_float((uint32_t)u);
    if (a >= 1e-10f && a <= 3.402823466e38f) bad0 += __float_as_uint(pow34(a)) != __float_as_uint(pow34_reference(a));
    if ((((uint32_t)u >> 23) & 255) != 255) {
      bad1 += __float_as_uint(div_exact(a, 9.0f, 1.0f / 9.0f)) != __float_as_uint(__fdiv_rn(a, 9.0f));
      bad2 += __float_as_uint(div_exact(a, 3.0f, 1.0f / 3.0f)) != __float_as_uint(__fdiv_rn(a, 3.0f));
    }
  }
  if (bad0) atomicAdd(mismatch, bad0);
  if (bad1) atomicAdd(mismatch + 1, bad1);
  if (bad2) atomicAdd(mismatch + 2, bad2);
}
int launch_selftest(unsigned long long *d_mismatch, cudaStream_t st) {
  k_selftest<<<148 * 8, 256, 0, st>>>(d_mismatch);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -(int)e;
}

// ------------------------------------------------------------------------------------------------------------
// launchers
static inline int check(int launched) { cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? launched : -(int)e; }

int launch_prepass(const Config &cfg, const PassBuffers &pb, cudaStream_t st) {
  if (pb.max_frames <= 0) return 0;
  dim3 grid(cfg.n_streams, (pb.max_frames + 3) / 4);
  k_prepass<<<grid, 128, 0, st>>>(cfg, pb);
  int n = 1;
  if (cfg.vbr) { dim3 g2(cfg.n_streams, (pb.max_frames + 63) / 64); k_bitrate<<<g2, 64, 0, st>>>(cfg, pb); ++n; }
  return check(n);
}
int launch_spectrum(const Config &cfg, const PassBuffers &pb, cudaStream_t st) {
  if (pb.max_frames <= 0) return 0;
  static bool attr_set[64] = {};
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 64 && !attr_set[dev]) {
    cudaFuncSetAttribute(k_filterbank, cudaFuncAttributeMaxDynamicSharedMemorySize, kFbSmemBytes);
    cudaFuncSetAttribute(k_filterbank, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr_set[dev] = true;
  }
  // run length: 128 granules (2304 steps = 9 full tiles) when that still gives every SM several CTAs, shorter runs for
  // small batches (64 -> 4.5 tiles, 28 -> 1.97, 14 -> 0.98)
  const int ngr = 2 * pb.max_frames;
  int R = 128;
  while (R > 14 && (long long)cfg.n_streams * cfg.channels * ((ngr + R - 1) / R) < 148 * 6) R = R > 64 ? 64 : R > 28 ? 28 : 14;
  if (R > ngr) R = ngr;
  dim3 grid(cfg.channels, cfg.n_streams, (ngr + R - 1) / R);
  static int dbg = -1;
  if (dbg < 0) { const char *v = getenv("MP3B_FB_DEBUG"); dbg = v ? atoi(v) : 0; }
  if (pb.tc_b) {                                                   // opt-in: matrixing on the tensor cores (filterbank_tc.cuh)
    static bool tc_attr[64] = {};
    if (dev < 64 && !tc_attr[dev]) {
      cudaFuncSetAttribute(k_filterbank_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes(1));
      cudaFuncSetAttribute(k_filterbank_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes(2));
      cudaFuncSetAttribute(k_filterbank_tc<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      cudaFuncSetAttribute(k_filterbank_tc<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      tc_attr[dev] = true;
    }
    if (cfg.channels == 2) k_filterbank_tc<2><<<grid, kTcThreads, kTcSmemBytes(2), st>>>(cfg, pb, R | dbg << 16);
    else k_filterbank_tc<1><<<grid, kTcThreads, kTcSmemBytes(1), st>>>(cfg, pb, R | dbg << 16);
    return check(1);
  }
  k_filterbank<<<grid, kFbThreads, kFbSmemBytes, st>>>(cfg, pb, R | dbg << 16);
  return check(1);
}
int launch_curve(const Config &cfg, const PassBuffers &pb, cudaStream_t st, bool fused_prepass) {
  if (pb.max_frames <= 0) return 0;
  dim3 grid(cfg.n_streams, (pb.max_frames * 2 * cfg.channels + kGrWarps * kGranulePerWarp - 1) / (kGrWarps * kGranulePerWarp));
  if (cfg.iso) { if (pb.spec) k_granule<true, false, true><<<grid, 32 * kGrWarps, 0, st>>>(cfg, pb); else k_granule<false, false, true><<<grid, 32 * kGrWarps, 0, st>>>(cfg, pb); }
  else if (pb.spec) k_granule<true, false, false><<<grid, 32 * kGrWarps, 0, st>>>(cfg, pb);
  else if (fused_prepass) { if (cfg.channels == 2) k_granule<false, true, false, 2><<<grid, 32 * kGrWarps, 0, st>>>(cfg, pb); else k_granule<false, true, false, 1><<<grid, 32 * kGrWarps, 0, st>>>(cfg, pb); }
  else k_granule<false, false, false><<<grid, 32 * kGrWarps, 0, st>>>(cfg, pb);
  return check(1);
}
int launch_blocktype(const Config &cfg, const PassBuffers &pb, cudaStream_t st) {
  if (pb.max_frames <= 0) return 0;
  dim3 grid(cfg.n_streams, (2 * pb.max_frames + 3) / 4);
  k_iso_blocktype<<<grid, 128, 0, st>>>(cfg, pb);
  return check(1);
}
int launch_psy(const Config &cfg, const PassBuffers &pb, cudaStream_t st) {
  if (pb.max_frames <= 0) return 0;
  dim3 grid(cfg.n_streams, (pb.max_frames * 2 * cfg.channels + kPsyWarps - 1) / kPsyWarps);
  static bool psy_attr[64];
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 64 && !psy_attr[dev]) { cudaFuncSetAttribute(k_psy, cudaFuncAttributeMaxDynamicSharedMemorySize, kPsySmemBytes); psy_attr[dev] = true; }
  k_psy<<<grid, 32 * kPsyWarps, kPsySmemBytes, st>>>(cfg, pb);
  return check(1);
}
int launch_outer(const Config &cfg, const PassBuffers &pb, cudaStream_t st) {
  if (pb.max_frames <= 0) return 0;
  dim3 grid(cfg.n_streams, (pb.max_frames * 2 * cfg.channels + kOuterWarps - 1) / kOuterWarps);
  k_outer<<<grid, 32 * kOuterWarps, 0, st>>>(cfg, pb);
  return check(1);
}
int launch_scan(const Config &cfg, const PassBuffers &pb, cudaStream_t st) {
  k_scan<<<cfg.n_streams, 32, 0, st>>>(cfg, pb);
  return check(1);
}
int launch_clear_md(const Config &cfg, const PassBuffers &pb, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(pb.md, 0, (size_t)cfg.n_streams * pb.md_stride + 16, st);
  return e == cudaSuccess ? 0 : -(int)e;
}
int launch_pack(const Config &cfg, const PassBuffers &pb, cudaStream_t st) {
  if (pb.max_frames <= 0) return 0;
  dim3 grid(cfg.n_streams, (pb.max_frames + kPackFramesPerCta - 1) / kPackFramesPerCta);
  const int nt = 32 * kPackFramesPerCta;
  if (cfg.iso) {
    if (pb.tr_ix) k_pack_iso<true><<<grid, nt, 0, st>>>(cfg, pb); else k_pack_iso<false><<<grid, nt, 0, st>>>(cfg, pb);
    return check(1);
  }
  if (pb.tr_ix) k_pack<true><<<grid, nt, 0, st>>>(cfg, pb); else k_pack<false><<<grid, nt, 0, st>>>(cfg, pb);
  return check(1);
}
int launch_frames(const Config &cfg, const PassBuffers &pb, cudaStream_t st) {
  dim3 grid(cfg.n_streams, (pb.max_frames + 1 + 3) / 4);
  k_frames<<<grid, 128, 0, st>>>(cfg, pb);
  return check(1);
}
int launch_carry_head(const Config &cfg, const PassBuffers &pb, cudaStream_t st) {
  k_carry_head<<<cfg.n_streams, 256, 0, st>>>(cfg, pb);
  return check(1);
}
int launch_carry_tail(const Config &cfg, const PassBuffers &pb, cudaStream_t st) {
  k_carry_tail<<<cfg.n_streams, 256, 0, st>>>(cfg, pb);
  return check(1);
}
int launch_thresholds(const Config &cfg, const PassBuffers &pb, cudaStream_t st) {
  if (pb.max_frames <= 0) return 0;
  dim3 grid(cfg.n_streams, (pb.max_frames * 2 * cfg.channels + 7) / 8);
  k_thresholds<<<grid, 256, 0, st>>>(cfg, pb);
  return check(1);
}
int launch_compact(const Config &cfg, const PassBuffers &pb, uint64_t *offsets, uint8_t *compact, int gather, cudaStream_t st) {
  if (!gather) { k_offsets<<<1, 1024, 0, st>>>(cfg, pb, offsets); return check(1); }
  dim3 grid(cfg.n_streams, 8);
  k_gather<<<grid, 256, 0, st>>>(cfg, pb, offsets, compact);
  return check(1);
}
int launch_synth(float *d_pcm, size_t n_per_channel, int channels, int sample_rate, float f_left, float f_right, float amp,
                 float noise, uint64_t seed, cudaStream_t st) {
  if (!n_per_channel) return 0;
  k_synth<<<(unsigned)((n_per_channel + 255) / 256), 256, 0, st>>>(d_pcm, n_per_channel, channels, sample_rate, f_left, f_right,
                                                                   amp, noise, seed);
  return check(1);
}

}  // namespace mp3b
