// ISO mode, level 2 (north_star stages (3) and (4)): the psychoacoustic model and the scalefactor
// outer loop the reference only has dead stubs for (ScaleFactorBands.scale SRC:1831-1876, ScaleFactorCompression SRC:2017-2037;
// its live "model", SRC:1983-2013, is a band mean whose output nobody reads).  There is no reference behaviour to be identical
// to, so the model is DEFINED here (and restated in numpy by tests/psymodel.py): the structure of ISO 11172-3 psychoacoustic
// model 2 made stateless, so that every granule-channel is an independent unit of work (batched across granules and channels):
//   * one 1024-point FFT (Hann) centred on the granule's MDCT window gives the line energies; three 256-point FFTs (Hann, hop
//     192) give the unpredictability: the third spectrum predicted from the first two (r^ = 2 r1 - r0, phi^ = 2 phi1 - phi0) —
//     ISO predicts from the two PREVIOUS granules, which would chain the granules together;
//   * partitions of 1/3 Bark, the ISO spreading function, tonality tb = -0.299 - 0.43 ln(cb / eb) in [0, 1], SNR =
//     max(minval, 29 tb + 6 (1 - tb)) dB, threshold = max(absolute threshold, spread energy * norm * 10^(-SNR / 10));
//   * perceptual entropy PE = sum n_lines ln((eb + 1) / (thr + 1)) — the serial scan turns it into the granule's share of the
//     bit reservoir; thresholds mapped onto the 22 long scalefactor bands as ratio = threshold / energy.
// k_outer then runs the ISO outer loop, one warp per granule-channel: quantize at the smallest global_gain that fits the
// granule's nominal budget, measure the quantization noise per band (warp-reduced), amplify every band whose noise exceeds
// ratio * band energy (by one scalefactor step per factor of two it is over, at most three), repeat until no band is over, all are amplified or a scalefactor would exceed
// its field; the best set (fewest bands over) is kept, scalefac_compress chosen, and the bits-vs-gain curve is produced with
// those scalefactors, exactly as in level 1.  Long blocks only.  Included by kernels.cu.
#pragma once

namespace mp3b {

constexpr int kPsyWarps = 4;
constexpr int kPsySmemBytes = kPsyWarps * 1280 * 8;
constexpr int kGainMaxIso = 319;                   // = kIsoGainMax
__constant__ float c_step_iso[kGainMaxIso + 1];    // 2^((G - 210) / 4): the decoder's step
__constant__ float c_amp34[16];                    // 2^(0.375 sf): what a scalefactor does to |xr|^0.75 (scalefac_scale = 0)
__constant__ float c_ampinv[16];                   // 2^(-sf / 2)

cudaError_t upload_psy_constants() {
  float step[kGainMaxIso + 1], a34[16], ainv[16];
  for (int g = 0; g <= kGainMaxIso; ++g) step[g] = (float)exp2((g - 210) / 4.0);
  for (int i = 0; i < 16; ++i) { a34[i] = (float)exp2(0.375 * i); ainv[i] = (float)exp2(-0.5 * i); }
  cudaError_t e;
  if ((e = cudaMemcpyToSymbol(c_step_iso, step, sizeof step))) return e;
  if ((e = cudaMemcpyToSymbol(c_amp34, a34, sizeof a34))) return e;
  return cudaMemcpyToSymbol(c_ampinv, ainv, sizeof ainv);
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// In-place radix-4 decimation-in-frequency FFT of N = 4^k points by one warp; tw[j] = exp(-2 pi i j / 1024).  X[k] ends up at
// the base-4 digit reversal of k.
// Element i lives at x[fpad(i)]: one element of padding after every four.  The last two passes (butterfly spans 4 and 1) touch
// elements 16 resp. 4 apart per lane — unpadded that is 4 of the 32 banks (four times the shared-memory wavefronts needed),
// padded a half-warp covers all of them.  (The array base must be a multiple of 4 elements.)
__device__ __forceinline__ int fpad(int i) { return i + (i >> 2); }
template <int N> __device__ __forceinline__ void warp_fft4(float2 *x, const float2 *tw, int lane) {
#pragma unroll 1
  for (int L = N; L >= 4; L >>= 2) {
    const int q = L >> 2, ts = 1024 / L;
#pragma unroll 2
    for (int t = lane; t < N / 4; t += 32) {
      const int pos = t & (q - 1), i0 = ((t - pos) << 2) + pos;
      const int ja = fpad(i0), jb = fpad(i0 + q), jc = fpad(i0 + 2 * q), jd = fpad(i0 + 3 * q);
      const float2 a = x[ja], b = x[jb], c = x[jc], d = x[jd];
      const float2 apc = make_float2(a.x + c.x, a.y + c.y), amc = make_float2(a.x - c.x, a.y - c.y);
      const float2 bpd = make_float2(b.x + d.x, b.y + d.y), bmd = make_float2(b.x - d.x, b.y - d.y);
      x[ja] = make_float2(apc.x + bpd.x, apc.y + bpd.y);
      x[jb] = cmul(make_float2(amc.x + bmd.y, amc.y - bmd.x), tw[pos * ts]);
      x[jc] = cmul(make_float2(apc.x - bpd.x, apc.y - bpd.y), tw[2 * pos * ts]);
      x[jd] = cmul(make_float2(amc.x - bmd.y, amc.y + bmd.x), tw[3 * pos * ts]);
    }
    __syncwarp();
  }
}
__device__ __forceinline__ int rev4_1024(int k) { return (k & 3) << 8 | ((k >> 2) & 3) << 6 | ((k >> 4) & 3) << 4 | ((k >> 6) & 3) << 2 | ((k >> 8) & 3); }
__device__ __forceinline__ int rev4_256(int k) { return (k & 3) << 6 | ((k >> 2) & 3) << 4 | ((k >> 4) & 3) << 2 | ((k >> 6) & 3); }
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// K3 (north_star stage 3): one warp per granule-channel.  Output: gc_psy[gslot][0..21] = threshold / energy per long
// scalefactor band, [22] = perceptual entropy, [23] = mean tonality (trace).
__global__ void __launch_bounds__(32 * kPsyWarps) k_psy(Config cfg, PassBuffers pb) {
  extern __shared__ __align__(16) float2 s_x_dyn[];               // [kPsyWarps][1280]: 1024 elements per warp, padded (fpad); dynamic: with it the CTA needs 52 KB
  __shared__ __align__(16) float2 s_tw[768];
  __shared__ float s_cw[kPsyWarps][132];
  __shared__ float s_part[kPsyWarps][3][kPsyMaxPart];
  const PsyTab &T = *pb.psy;
  const int s = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 768; i += 32 * kPsyWarps) s_tw[i] = T.tw[i];
  __syncthreads();
  const int ch = cfg.channels, chs = ch - 1;
  const int gci = blockIdx.y * kPsyWarps + warp;
  if (gci >= (int)pb.plan[s].n_frames * 2 * ch) return;
  const size_t gslot = (size_t)s * pb.GC + gci;
  const int g = gci >> chs, c = gci & chs, f = g >> 1;
  const bool ms = cfg.mode == 2 && pb.ms[(size_t)s * (pb.Fc + 1) + 1 + f];
  // the 1024 samples centred on the granule's MDCT window (the filterbank delays by 256): [576 g - 768, 576 g + 256) of the
  // pass, on the decoder's scale; lane holds n = lane + 32 i
  const PcmView pv = pcm_view(cfg, pb, s);
  float v[32];
  {
    // (level 3 codes the signal one granule late; its window is the last 1024 samples the carried frame reaches back to)
    const int64_t m0 = (int64_t)576 * g - (cfg.iso >= 3 ? 1152 : 768) + lane;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int64_t q = (int64_t)cfg.fsc + (m0 + 32 * i) * ch;
      float x;
      if (ch == 1) x = pv.at(q);
      else if (!ms) x = pv.at(q + c);
      else { const float l = pv.at(q), r = pv.at(q + 1); x = (c == 0 ? l + r : l - r) * cfg.ms_scale; }
      v[i] = x * 32768.0f;
    }
  }
  float2 *x = s_x_dyn + warp * 1280;
  // ---- unpredictability from three 256-point FFTs: windows start at 192, 384, 576 of the 1024
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int n = lane + 32 * i;
#pragma unroll
    for (int w = 0; w < 3; ++w) {
      const int k = n - 192 - 192 * w;
      if (k >= 0 && k < 256) x[fpad(256 * w + k)] = make_float2(v[i] * T.hann256[k], 0.0f);
    }
  }
  __syncwarp();
  warp_fft4<256>(x, s_tw, lane); warp_fft4<256>(x + fpad(256), s_tw, lane); warp_fft4<256>(x + fpad(512), s_tw, lane);
  for (int j = lane; j <= 128; j += 32) {
    const int r = rev4_256(j & 255);
    const float2 a0 = x[fpad(r)], a1 = x[fpad(256 + r)], a2 = x[fpad(512 + r)];
    const float r0 = sqrtf(a0.x * a0.x + a0.y * a0.y), r1 = sqrtf(a1.x * a1.x + a1.y * a1.y), r2 = sqrtf(a2.x * a2.x + a2.y * a2.y);
    // unit vectors; e^(i (2 phi1 - phi0)) = u1^2 conj(u0)
    const float2 u0 = r0 > 0.0f ? make_float2(a0.x / r0, a0.y / r0) : make_float2(1.0f, 0.0f);
    const float2 u1 = r1 > 0.0f ? make_float2(a1.x / r1, a1.y / r1) : make_float2(1.0f, 0.0f);
    const float2 up = cmul(cmul(u1, u1), make_float2(u0.x, -u0.y));
    const float rp = 2.0f * r1 - r0;
    const float dx = a2.x - rp * up.x, dy = a2.y - rp * up.y;
    const float den = r2 + fabsf(rp);
    s_cw[warp][j] = den > 0.0f ? sqrtf(dx * dx + dy * dy) / den : 0.0f;
  }
  __syncwarp();
  // ---- line energies from the 1024-point FFT
#pragma unroll
  for (int i = 0; i < 32; ++i) x[fpad(lane + 32 * i)] = make_float2(v[i] * T.hann1024[lane + 32 * i], 0.0f);
  __syncwarp();
  warp_fft4<1024>(x, s_tw, lane);
  float e[16], cw[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const int k = lane + 32 * t;
    const float2 a = x[fpad(rev4_1024(k))];
    e[t] = a.x * a.x + a.y * a.y;
    cw[t] = e[t] * (k < 206 ? s_cw[warp][(k + 2) >> 2] : 0.4f);
  }
  __syncwarp();
  float *se = reinterpret_cast<float *>(x);                       // [0, 512): e, [512, 1024): e * cw
#pragma unroll
  for (int t = 0; t < 16; ++t) { se[lane + 32 * t] = e[t]; se[512 + lane + 32 * t] = cw[t]; }
  __syncwarp();
  // ---- partitions: lane = partition b, b + 32, b + 64
  const int np = T.n_part;
  for (int b = lane; b < np; b += 32) {
    const int lo = T.part_lo[b], n = T.part_n[b];
    float eb = 0.0f, cb = 0.0f;
    for (int k = lo; k < lo + n; ++k) { eb += se[k]; cb += se[512 + k]; }
    s_part[warp][0][b] = eb; s_part[warp][1][b] = cb;
  }
  __syncwarp();
  float pe = 0.0f, tsum = 0.0f;
  for (int b = lane; b < np; b += 32) {
    float ecb = 0.0f, ctb = 0.0f;
    for (int j = 0; j < np; ++j) { const float w = T.s3t[j * kPsyMaxPart + b]; ecb += w * s_part[warp][0][j]; ctb += w * s_part[warp][1][j]; }
    const float cbb = ecb > 0.0f ? ctb / ecb : 0.0f;
    const float tb = cbb > 0.0f ? fminf(fmaxf(-0.299f - 0.43f * logf(cbb), 0.0f), 1.0f) : 1.0f;
    const float snr = fmaxf(T.minval[b], 29.0f * tb + 6.0f * (1.0f - tb));
    const float nb = ecb * T.rnorm[b] * exp10f(-0.1f * snr);
    const float thr = fmaxf(T.qthr[b], nb);
    const float eb = s_part[warp][0][b];
    pe += (float)T.part_n[b] * logf((eb + 1.0f) / (thr + 1.0f));
    tsum += tb;
    s_part[warp][2][b] = thr / (float)T.part_n[b];
  }
  pe = warp_sum_f(pe); tsum = warp_sum_f(tsum);
  __syncwarp();
  float *out = pb.gc_psy + gslot * 24;
  if (lane < 22) {
    float en = 0.0f, th = 0.0f;
    for (int k = T.sfb_line[lane]; k < T.sfb_line[lane + 1]; ++k) { en += se[k]; th += s_part[warp][2][T.line_part[k]]; }
    out[lane] = th / fmaxf(en, 1e-20f);
  } else if (lane == 22) out[22] = fmaxf(pe, 0.0f);
  else if (lane == 23) out[23] = tsum / (float)np;
}

// x^(4/3) for the noise estimate (not bit-critical: it steers the scalefactors, the bitstream's own consistency is checked elsewhere)
__device__ __forceinline__ float pow43_fast(float x) { return x > 0.0f ? exp2f(log2f(x) * (4.0f / 3.0f)) : 0.0f; }

constexpr int kOuterWarps = 4;
constexpr int kOuterMaxIter = 8;
__device__ __forceinline__ int slen_need(int m) { return m == 0 ? 0 : m < 2 ? 1 : m < 4 ? 2 : m < 8 ? 3 : 4; }
// scalefac_compress -> (slen1, slen2), ISO 11172-3 2.4.2.7
__constant__ uint8_t c_slen1[16] = {0, 0, 0, 0, 3, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4};
__constant__ uint8_t c_slen2[16] = {0, 1, 2, 3, 0, 1, 2, 3, 1, 2, 3, 1, 2, 3, 2, 3};

// K4 in ISO mode level 2 (north_star stage 4): the outer loop.  One warp per granule-channel.
__global__ void __launch_bounds__(32 * kOuterWarps, 8) k_outer(Config cfg, PassBuffers pb) {
  // The kernel is a chain of searches, each step one call of the shared evaluation function: latency-bound, so it wants resident
  // warps, i.e. few registers — everything per-line lives in shared memory (amplified magnitudes, |xr|, band of a pair), the
  // magnitudes themselves are re-read from HBM / L2 whenever the scalefactors change (at most eight times).
  __shared__ __align__(16) uint8_t s_len[(kHuffEntries + 15) / 16 * 16];
  __shared__ uint8_t s_c[kOuterWarps][288];
  __shared__ uint8_t s_bnd[288];                                  // long scalefactor band of pair p
  __shared__ float s_val[kOuterWarps][288];
  __shared__ __align__(8) float2 s_mag[kOuterWarps][288];         // magnitudes |xr|^0.75 amplified by the current scalefactors
  __shared__ __align__(8) float2 s_ax[kOuterWarps][288];          // |xr| on the decoder's scale
  __shared__ int s_sf[kOuterWarps][2][24];                        // current and best scalefactors
  __shared__ float s_p43[256];                                    // ix^(4/3) for the values nearly all lines quantize to
  const int s = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int *sfb = c_sfb_cum[cfg.sfb_index];
  for (int i = threadIdx.x; i < kHuffEntries; i += 32 * kOuterWarps) s_len[i] = kHuffLenFlat[i];
  for (int i = threadIdx.x; i < 256; i += 32 * kOuterWarps) s_p43[i] = pow43_fast((float)i);
  for (int p = threadIdx.x; p < 288; p += 32 * kOuterWarps) {
    int b = 0;
#pragma unroll
    for (int i = 0; i < 21; ++i) b += sfb[i] <= 2 * p;
    s_bnd[p] = (uint8_t)b;
  }
  __syncthreads();
  const int ch = cfg.channels, chs = ch - 1;
  const int gci = blockIdx.y * kOuterWarps + warp;
  if (gci >= (int)pb.plan[s].n_frames * 2 * ch) return;
  const size_t gslot = (size_t)s * pb.GC + gci;
  const int bt = pb.gc_bt[gslot] & 3;                            // level 3: 1 start, 2 short, 3 stop
  const bool ws = bt != 0;
  const int f = gci >> (chs + 1);
  const int bri = pb.frame_br[(size_t)s * pb.Fc + f];
  const int lo_bits = min(lo_bits_of(cfg, bri), 4095);
  const int mds1 = cfg.frame_base[bri] + 1 - cfg.header_bytes;
  const int hi_bits = min(4095, (mds1 * 8 + min(511, mds1) * 8) >> cfg.channels);
  // pairs p = lane + 32 j: magnitudes |xr|^0.75 (k_granule left them behind), |xr| on the decoder's scale
  const float2 *sm2 = reinterpret_cast<const float2 *>(pb.smag + gslot * 576);
  float2 *s_m = s_mag[warp], *s_a = s_ax[warp];
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    const float2 v = __ldg(sm2 + lane + 32 * j);
    const float ax = 32768.0f * pow43_fast(fabsf(v.x)), ay = 32768.0f * pow43_fast(fabsf(v.y));
    s_a[lane + 32 * j] = make_float2(ax, ay);
    s_val[warp][lane + 32 * j] = ax * ax + ay * ay;
  }
  // band sums of a per-pair quantity: lane b adds up the pairs of band b
  const int p_lo = lane == 0 ? 0 : lane < 22 ? sfb[lane - 1] >> 1 : 288, p_hi = lane < 21 ? sfb[lane] >> 1 : 288;
  auto band_sum = [&]() { float a = 0.0f; for (int p = p_lo; p < p_hi; ++p) a += s_val[warp][p]; return a; };
  __syncwarp();
  const float xmin = lane < 22 ? pb.gc_psy[gslot * 24 + lane] * band_sum() : 0.0f;
  __syncwarp();
  int *sf = s_sf[warp][0], *best = s_sf[warp][1];
  if (lane < 24) { sf[lane] = 0; best[lane] = 0; }
  __syncwarp();
  // the amplified magnitudes of the current scalefactors (a lane reads back only what it wrote)
  auto load_amp = [&](const int *q) {
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const float2 v = __ldg(sm2 + lane + 32 * j);
      const float a34 = c_amp34[q[s_bnd[lane + 32 * j]]];
      s_m[lane + 32 * j] = make_float2(__fmul_rn(fabsf(v.x), a34), __fmul_rn(fabsf(v.y), a34));
    }
  };
  auto eval = [&](int G) { return iso_eval_gain<3>(G, s_m, reinterpret_cast<uint32_t *>(s_val[warp]), lane, s_len, s_c[warp], sfb, ws); };   // (s_val is free during the searches)   // min(bits, 65535) | big_values << 16
  auto bits_of = [](uint32_t c) { return (int)(c & 0xFFFFu); };
  auto part2_of = [&](const int *q, int &sfc) {                  // cheapest scalefac_compress that holds the scalefactors
    int m1 = 0, m2 = 0;
    for (int i = 0; i < 11; ++i) m1 = max(m1, q[i]);
    for (int i = 11; i < 21; ++i) m2 = max(m2, q[i]);
    const int n1 = slen_need(m1), n2 = slen_need(m2);
    int bits = 1 << 30; sfc = 15;
    for (int k = 0; k < 16; ++k)
      if (c_slen1[k] >= n1 && c_slen2[k] >= n2 && 11 * c_slen1[k] + 10 * c_slen2[k] < bits) { bits = 11 * c_slen1[k] + 10 * c_slen2[k]; sfc = k; }
    return bits;
  };
  // smallest gain whose count fits `budget` (the count falls as the gain rises).  Every evaluation costs several hundred
  // instructions, so the searches start where the answer is expected: search_up gallops upwards from a known lower bound (the
  // previous outer iteration's gain: amplifying bands only adds bits), search_down downwards from a gain known to fit (the curve
  // starts below the outer loop's gain: its budget is larger), and only the very first search bisects the whole range.
  auto bisect = [&](int lo, int hi, int budget) {                  // invariant: gains < lo do not fit, hi fits
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (bits_of(eval(mid)) <= budget) hi = mid; else lo = mid + 1; }
    return hi;
  };
  auto search_up = [&](int from, int budget) {
    int lo = from, hi = kIsoGainMax;
    for (int stp = 1; lo + stp < hi; stp <<= 1) {
      if (bits_of(eval(lo + stp - 1)) <= budget) { hi = lo + stp - 1; break; }
      lo = lo + stp;
    }
    return bisect(lo, hi, budget);
  };
  auto search_down = [&](int fits, int budget) {
    int hi = fits, lo = 0;
    for (int stp = 1; hi - stp > 0; stp <<= 1) {
      if (bits_of(eval(hi - stp)) > budget) { lo = hi - stp + 1; break; }
      hi = hi - stp;
    }
    return bisect(lo, hi, budget);
  };
  // ---- outer loop at the granule's nominal budget
  int G = 0, best_over = 99, n_iter = 0, best_G = 0;
  // short blocks keep their scalefactors at 0: their lines are ordered by short scalefactor band and window, the thresholds are
  // per long band; one gain search gives best_G for the curve
  for (int it = 0; it < (bt == 2 ? 1 : kOuterMaxIter); ++it) {
    load_amp(sf);
    int sfc;
    const int part2 = part2_of(sf, sfc);
    G = it == 0 ? bisect(0, kIsoGainMax, max(lo_bits - part2, 0)) : search_up(G, max(lo_bits - part2, 0));
    const float inv = c_inv_step_iso[G], step = c_step_iso[G];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const float2 m = s_m[lane + 32 * j], a = s_a[lane + 32 * j];
      const int qx = iso_quant(m.x, inv), qy = iso_quant(m.y, inv);
      const float back = step * c_ampinv[sf[s_bnd[lane + 32 * j]]];
      const float px = qx < 256 ? s_p43[qx] : pow43_fast((float)qx), py = qy < 256 ? s_p43[qy] : pow43_fast((float)qy);
      const float dx = a.x - px * back, dy = a.y - py * back;
      s_val[warp][lane + 32 * j] = dx * dx + dy * dy;
    }
    __syncwarp();
    const float noise = band_sum();
    __syncwarp();
    const unsigned over = __ballot_sync(0xffffffffu, lane < 21 && noise > xmin);
    const int n_over = __popc(over) + (int)((__ballot_sync(0xffffffffu, lane == 21 && noise > xmin) >> 21) & 1u);
    n_iter = it + 1;
    if (n_over < best_over) { best_over = n_over; best_G = G; if (lane < 21) best[lane] = sf[lane]; __syncwarp(); }
    if (over == 0u || bt == 2) break;
    // amplify the bands over their threshold — by one step per factor of two the noise is above it (a step takes 1.5 dB off the
    // band's noise at equal gain), at most three at once; stop when a scalefactor would leave its field or every band is amplified
    const bool mine = lane < 21 && ((over >> lane) & 1u);
    const int lim = lane < 11 ? 15 : 7;
    int add = 1;
    if (mine) { const float r = noise / fmaxf(xmin, 1e-30f); add = r >= 4.0f ? 3 : r >= 2.0f ? 2 : 1; add = min(add, max(lim - sf[lane], 1)); }
    if (__ballot_sync(0xffffffffu, mine && sf[lane] + add > lim)) break;
    if (mine) sf[lane] += add;
    __syncwarp();
    if (__ballot_sync(0xffffffffu, lane < 21 && sf[lane] > 0) == 0x1FFFFFu) break;
  }
  // ---- the curve with the best scalefactors (as k_granule's ISO branch, counts including part2)
  load_amp(best);
  int sfc;
  const int part2 = part2_of(best, sfc);
  uint16_t *bits_out = pb.gc_bits + gslot * kMaxEntries, *bv_out = pb.gc_bv + gslot * kMaxEntries;
  const int g_first = search_down(max(best_G, 1), max(hi_bits - part2, 0));   // best_G fits the smaller budget lo_bits - part2, so it fits this one
  int n = 0, g_last = g_first, fitted = 0;
  for (int e = 0; e < kMaxEntries - 1 && !fitted; ++e) {
    const int Ge = min(g_first + e, kIsoGainMax);
    const uint32_t c = eval(Ge);
    if (lane == 0) { bits_out[e] = (uint16_t)min(bits_of(c) + part2, 65535); bv_out[e] = (uint16_t)(c >> 16); }
    n = e + 1; g_last = Ge;
    fitted = bits_of(c) + part2 <= lo_bits || Ge == kIsoGainMax;
  }
  if (!fitted) {
    const int Gl = search_up(min(g_first + kMaxEntries - 1, kIsoGainMax), max(lo_bits - part2, 0));
    const uint32_t c = eval(Gl);
    if (lane == 0) { bits_out[kMaxEntries - 1] = (uint16_t)min(bits_of(c) + part2, 65535); bv_out[kMaxEntries - 1] = (uint16_t)(c >> 16); }
    n = kMaxEntries; g_last = Gl;
  }
  if (lane == 0) pb.gc_meta[gslot] = (uint32_t)g_first | (uint32_t)n << 9 | (uint32_t)g_last << 14;
  uint8_t *o = pb.gc_sf + gslot * 24;
  if (lane < 21) o[lane] = (uint8_t)best[lane];
  else if (lane == 21) o[21] = (uint8_t)sfc;
  else if (lane == 22) o[22] = (uint8_t)part2;
  else if (lane == 23) o[23] = (uint8_t)(min(best_over, 15) | min(n_iter, 15) << 4);
}

}  // namespace mp3b
