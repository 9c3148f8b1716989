// K1 on the tensor cores (opt-in: mp3b_batch_set_matrixing(b, 1)).  north_star stage (1): "its 32x64 cosine matrixing runs on
// tensor cores only if 3xTF32 split-precision stays inside tolerance, and on FP32 FMA otherwise" — the default stays FP32 FMA
// (bit-exact with the oracle); this kernel is the measured alternative (tests/test_gpu_parity.py::test_tensor_core_matrixing,
// profiles/).  Same CTA shape as k_filterbank (a run of granules of one (stream, channel), PCM rows staged by cp.async), the
// windowing is the same FFMA2 code with the same two roundings, but the 32 x 64 matrix-vector products of 128 steps become
// tcgen05.mma tiles:
//   D[128 steps x 96] (TMEM, FP32) += A[128 steps x 8 n] (shared memory, K-major, 128-byte swizzle) * B[96 x 8 n]
// with every FP32 operand split exactly into three TF32 terms (x = hi + mid + lo, 11 + 11 + 2 mantissa bits):
//   A = Yhi against B = [Mhi | Mmid | Mlo]  (N = 96)     columns  0-31: Y M for the hi-hi, mid-hi, lo-hi products
//   A = Ymid against B = [Mhi | Mmid]       (N = 64)     columns 32-63: hi-mid, mid-mid
//   A = Ylo against B = [Mhi]               (N = 32)     columns 64-95: hi-lo
// i.e. the six products whose weight is above 2^-24 of the result; the epilogue (tcgen05.ld, one TMEM lane = one step per
// thread) adds the three column groups smallest first.  The split analysis matrix arrives pre-swizzled with one cp.async.bulk
// (TMA unit) per CTA.  n is processed in two halves of 32 (one 128-byte swizzle row each), so the A tile is 3 x 16 KB.
#pragma once

namespace mp3b {

constexpr int kTcTile = 128;                        // filterbank steps per tile = M of the MMA
constexpr int kTcPRows = kTcTile + kLook;
constexpr int kTcABytes = 3 * kTcTile * 128;        // three split terms of one n half: [128 rows][32 floats], 1024-byte swizzle atoms
constexpr int kTcBBytes = 2 * 96 * 128;             // two n halves of [96 rows = hi | mid | lo of the 32 subbands][32 floats]
constexpr int kTcThreads = 256, kTcWarps = kTcThreads / 32;
__host__ __device__ constexpr int kTcPBytes(int ch) { return kTcPRows * 128 * ch; }          // PCM rows stay interleaved: 32 * ch floats per row
__host__ __device__ constexpr int kTcSmemBytes(int ch) { return kTcABytes + kTcBBytes + kTcPBytes(ch) + 64 + 1024; }   // + barriers, + alignment slack
// true: window products and sums fused (one rounding per tap, 8 FFMA2 per step pair instead of 15).  The tensor-core path is not
// bit-exact with the oracle anyway; what this costs in changed granules is part of the measured flip rate.
#ifndef MP3B_TC_FUSED_WINDOW
#define MP3B_TC_FUSED_WINDOW 0
#endif

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row atoms 1024 bytes apart (sm_100 descriptor version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (uint64_t)(1024 >> 4) << 32 | (uint64_t)1 << 46 | (uint64_t)2 << 61;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc_tf32(int n) { return 1u << 4 | 2u << 7 | 2u << 10 | (uint32_t)(n >> 3) << 17 | (uint32_t)(128 >> 4) << 24; }
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                 "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
      "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// exact three-term TF32 split: hi and mid carry 11 significant bits each (the low 13 bits of a TF32 operand are not read), lo the rest
__device__ __forceinline__ void split3(float y, float &hi, float &mid, float &lo) {
  hi = __uint_as_float(__float_as_uint(y) & 0xFFFFE000u);
  const float r = __fsub_rn(y, hi);
  mid = __uint_as_float(__float_as_uint(r) & 0xFFFFE000u);
  lo = __fsub_rn(r, mid);
}

template <int CH> __global__ void __launch_bounds__(kTcThreads, 2) k_filterbank_tc(Config cfg, PassBuffers pb, int Rdbg) {
  const int R = Rdbg & 0xFFFF, dbg = Rdbg >> 16;     // dbg: timing experiments only (tools/stage_times.py), 0 in production
  extern __shared__ uint8_t tc_raw[];
  uint8_t *base = tc_raw + ((1024u - (smem_u32(tc_raw) & 1023u)) & 1023u);   // (pointer arithmetic, so that the accesses stay LDS / STS)
  uint8_t *sA = base;                               // [3 terms][128 rows][128 bytes], swizzled
  uint8_t *sB = sA + kTcABytes;                     // [2 halves][96 rows][128 bytes], swizzled (as it lies in pb.tc_b)
  float *P = reinterpret_cast<float *>(sB + kTcBBytes);   // [143][32 * CH] PCM rows of the tile, interleaved as in the input
  uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(P) + kTcPBytes(CH));   // [0] MMA done, [1] B arrived, [2] PCM arrived
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3);
  const int c = blockIdx.x, s = blockIdx.y, run = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const StreamPlan &plan = pb.plan[s];
  const int ngr = 2 * (int)plan.n_frames;
  const int g_begin = run * R;
  if (g_begin >= ngr) return;
  const int g_cnt = min(R, ngr - g_begin);
  const int U = 18 * g_cnt;
  const int n_tiles = (U + kTcTile - 1) / kTcTile;
  const int rows_total = kLook + U;
  const PcmView pv = pcm_view(cfg, pb, s);
  const uint8_t *msrow = pb.ms + (size_t)s * (pb.Fc + 1);
  const uint32_t ms_prev = pb.state[s].ms_prev;
  const bool joint = cfg.mode == 2;
  const int n_start = 576 * g_begin - 480 - cfg.iso_delay;
  float *out = pb.sub + ((size_t)(s * CH + c) * pb.sub_rows + 18 * (g_begin + 1)) * 32;
  const uint32_t bar_mma = smem_u32(bars), bar_b = smem_u32(bars + 1), bar_p = smem_u32(bars + 2);

  // ---- one-time setup: barriers, TMEM (128 columns: 96 used), the split matrix by one bulk copy
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(bar_mma, 1); mbar_init(bar_b, 1); mbar_init(bar_p, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_b), "r"(kTcBBytes));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sB)), "l"(pb.tc_b), "r"(kTcBBytes), "r"(bar_b) : "memory");
  }
  // PCM rows [ra, rb) of the run -> P rows slot0 ...  The usual case — the rows are plain PCM of this pass, 16-byte aligned — is
  // ONE bulk copy by the TMA unit (both channels of a stereo stream come along: the windowing reads with a stride of CH);
  // otherwise (carried head, zero padding, mid / side frames, odd alignment) row by row through the lanes.
  // Returns whether bar_p has to be waited for.
  auto load_rows = [&](int ra, int rb, int slot0) -> bool {
    rb = (dbg & 4) ? ra : min(rb, rows_total);
    if (ra >= rb) return false;
    const int64_t rel0 = (int64_t)(n_start + 32 * ra + 1152) * CH - (int64_t)pv.head_n;
    if (!joint && rel0 >= 0 && rel0 + (int64_t)(rb - ra) * 32 * CH <= (int64_t)pv.cur_n && ((reinterpret_cast<uintptr_t>(pv.cur + rel0) & 15) == 0)) {
      if (tid == 0) {
        const uint32_t bytes = (uint32_t)(rb - ra) * 128u * CH;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_p), "r"(bytes));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(P + slot0 * 32 * CH)), "l"(pv.cur + rel0), "r"(bytes), "r"(bar_p) : "memory");
      }
      return true;
    }
    const int lane_off = CH == 1 ? lane : 2 * lane + c;
    for (int r = ra + warp; r < rb; r += kTcWarps) {
      const int nrow = n_start + 32 * r;
      const int64_t q = (int64_t)(nrow + 1152) * CH, rel = q - (int64_t)pv.head_n;
      float *dstp = P + (slot0 + r - ra) * 32 * CH + lane_off;
      const float *src = nullptr;
      if (!joint) {                                  // a row that is plain PCM of this pass or of the carried head: 4-byte cp.async
        if (rel >= 0 && rel + 32 * CH <= (int64_t)pv.cur_n) src = pv.cur + rel + lane_off;
        else if (q >= 0 && q + 32 * CH <= (int64_t)pv.head_n) src = pv.head + q + lane_off;
      }
      if (src) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dstp)), "l"(src));
      } else {
        float v;
        if (CH == 1) v = pv.at(q + lane);
        else {
          const float l = pv.at(q + 2 * lane), rr = pv.at(q + 2 * lane + 1);
          const int fr = nrow >= 0 ? nrow / 1152 : -1;
          const bool ms = joint && (fr < 0 ? ms_prev != 0 : msrow[1 + fr] != 0);
          if (!ms) v = c == 0 ? l : rr;
          else v = c == 0 ? __fmul_rn(__fadd_rn(l, rr), cfg.ms_scale) : __fmul_rn(__fsub_rn(l, rr), cfg.ms_scale);
        }
        *dstp = v;
      }
    }
    asm volatile("cp.async.commit_group;");
    return false;
  };
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_d = *tmem_slot;
  bool p_wait = load_rows(0, kTcPRows, 0);
  uint32_t p_phase = 0;

  float wc[2][8];                                  // C[32 H + lane + 64 i]
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int i = 0; i < 8; ++i) wc[h][i] = __ldg(tab::kWindow + 32 * h + lane + 64 * i);
  const float2 neg0 = make_float2(cfg.f_neg0, cfg.f_neg0), one = make_float2(cfg.f_one, cfg.f_one);
  const uint32_t idesc96 = umma_idesc_tf32(96), idesc64 = umma_idesc_tf32(64), idesc32 = umma_idesc_tf32(32);
  uint32_t mma_phase = 0;
  bool b_ready = false;

  for (int tile = 0; tile < n_tiles; ++tile) {
    const int valid = min(kTcTile, U - kTcTile * tile);
    if (p_wait) { mbar_wait(bar_p, p_phase); p_phase ^= 1; }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                               // P has arrived; the epilogue of the previous tile is done with sA and with D
#pragma unroll
    for (int H = 0; H < 2; ++H) {
      // ---- windowing (SRC:1386-1399) of n = 32 H + lane, steps 16 warp ... 16 warp + 15: k_filterbank's loop over interleaved rows
      if (16 * warp < valid && !(dbg & 1)) {
        const float *Pc = P + ((16 * warp + 1 - H) * 32 + (31 - lane)) * CH + (CH == 2 ? c : 0);
        float2 q[8];
#pragma unroll
        for (int k = 1; k <= 7; ++k) { q[k].x = Pc[(2 * k - 2) * 32 * CH]; q[k].y = Pc[(2 * k - 1) * 32 * CH]; }
        Pc += 14 * 32 * CH;
        float2 ypair[8];
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) {
          q[ii].x = Pc[(2 * ii) * 32 * CH]; q[ii].y = Pc[(2 * ii + 1) * 32 * CH];
          float2 y;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 w2 = make_float2(wc[H][i], wc[H][i]);
            if (MP3B_TC_FUSED_WINDOW) y = i == 0 ? __ffma2_rn(q[(ii - i) & 7], w2, neg0) : __ffma2_rn(q[(ii - i) & 7], w2, y);
            else {
              const float2 z = __ffma2_rn(q[(ii - i) & 7], w2, neg0);
              y = i == 0 ? z : __ffma2_rn(y, one, z);
            }
          }
          ypair[ii] = y;
        }
        // Y[t][n] -> the three A terms: row t = 16 warp + 2 ii (+ 1), column n = lane of the 128-byte swizzle row
        uint8_t *arow = sA + 2 * warp * 1024 + (lane & 3) * 4;
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int tl = 2 * ii + e;               // row inside the warp's two 8-row atoms
            float hi, mid, lo;
            split3(e ? ypair[ii].y : ypair[ii].x, hi, mid, lo);
            uint8_t *p = arow + (tl >> 3) * 1024 + (tl & 7) * 128 + (((lane >> 2) ^ (tl & 7)) << 4);
            *reinterpret_cast<float *>(p) = hi;
            *reinterpret_cast<float *>(p + kTcTile * 128) = mid;
            *reinterpret_cast<float *>(p + 2 * kTcTile * 128) = lo;
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the A tile was written through the generic proxy, the MMA reads it through the async proxy
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncthreads();
      if (!(dbg & 2)) {
        if (tid == 0) {
          if (!b_ready) { mbar_wait(bar_b, 0); b_ready = true; }
          asm volatile("tcgen05.fence::after_thread_sync;");
          const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB) + H * (96 * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {              // K = 8 per instruction: 32 bytes along the swizzle row
            umma_tf32(tmem_d, umma_desc(a0 + 32 * k), umma_desc(b0 + 32 * k), idesc96, (H | k) != 0);
            umma_tf32(tmem_d, umma_desc(a0 + kTcTile * 128 + 32 * k), umma_desc(b0 + 32 * k), idesc64, 1);
            umma_tf32(tmem_d, umma_desc(a0 + 2 * kTcTile * 128 + 32 * k), umma_desc(b0 + 32 * k), idesc32, 1);
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_mma) : "memory");
        }
      }
      if (H == 1 && tile + 1 < n_tiles) {
        // P is free: the look-back of the next tile = the last 15 rows of this one; then the next tile's rows start to arrive
        // under the MMAs and the epilogue
        for (int i = tid; i < kLook * 32 * CH; i += kTcThreads) P[i] = P[kTcTile * 32 * CH + i];
        __syncthreads();
        p_wait = load_rows(kTcTile * (tile + 1) + kLook, kTcTile * (tile + 2) + kLook, kLook);
      }
      if (!(dbg & 2)) { mbar_wait(bar_mma, mma_phase); mma_phase ^= 1; }   // the A tile is reused by the next half
    }
    // ---- epilogue: lane of TMEM = step 32 (warp & 3) + lane, warps 0-3 take subbands 0-15, warps 4-7 subbands 16-31; the three
    // column groups are added smallest first; the warp's 32 x 16 block is transposed through 2 KB of (now idle) sA so that the
    // subband rows leave as 64-byte pieces of whole lines
    asm volatile("tcgen05.fence::after_thread_sync;");
    const int w4 = warp & 3, hs = warp >> 2;
    if (32 * w4 < valid && !(dbg & 16)) {
      float d0[16], d1[16];
      const uint32_t taddr = tmem_d + ((uint32_t)(32 * w4) << 16) + 16 * hs;
      tmem_ld16(taddr + 64, d0); tmem_ld16(taddr + 32, d1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) d0[i] = __fadd_rn(d0[i], d1[i]);
      tmem_ld16(taddr, d1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) d0[i] = __fadd_rn(d0[i], d1[i]);
      float4 *stg = reinterpret_cast<float4 *>(sA + warp * 2048);
#pragma unroll
      for (int j = 0; j < 4; ++j) stg[lane * 4 + (j ^ ((lane >> 1) & 3))] = make_float4(d0[4 * j], d0[4 * j + 1], d0[4 * j + 2], d0[4 * j + 3]);
      __syncwarp();
      const int rows = min(32, valid - 32 * w4), l3 = lane >> 2, cidx = lane & 3;
      float4 *dst = reinterpret_cast<float4 *>(out + (size_t)(kTcTile * tile + 32 * w4) * 32) + 4 * hs;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int row = 8 * it + l3;
        if (row < rows && !(dbg & 8)) dst[row * 8 + cidx] = stg[row * 4 + (cidx ^ ((row >> 1) & 3))];
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_d));
}


}  // namespace mp3b
