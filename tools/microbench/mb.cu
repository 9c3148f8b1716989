// Micro-benchmarks behind the k_spectrum2 design (run on a B200 through gpurun).
//   A  FFMA2 fed by uniform-register operands (LDCU from the constant bank) + one LDS.32 per 16 FFMA2
//   B  FFMA2 fed by two LDS.128 per 8 FFMA2 (4x4 register tile out of shared memory)
//   C  FFMA2 only (register operands): the pipe's own ceiling
#include <cstdio>
#include <cuda_runtime.h>
__constant__ float cM[64 * 32];
constexpr int kT = 128;
__global__ void __launch_bounds__(256, 2) mbA(const float *in, float *out, int reps) {
  __shared__ float Y[64 * kT];
  for (int i = threadIdx.x; i < 64 * kT; i += 256) Y[i] = in[i];
  __syncthreads();
  float2 acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = make_float2(0.f, 0.f);
  const float *y = Y + (threadIdx.x & 127);
  for (int r = 0; r < reps; ++r) {
#pragma unroll 2
    for (int n = 0; n < 64; ++n) {
      float yv = y[n * kT];
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[k] = __ffma2_rn(make_float2(yv, yv), make_float2(cM[n * 32 + 2 * k], cM[n * 32 + 2 * k + 1]), acc[k]);
    }
  }
  float s = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += acc[k].x + acc[k].y;
  out[blockIdx.x * 256 + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256, 2) mbB(const float *in, float *out, int reps) {
  __shared__ __align__(16) float Y[64 * kT];
  __shared__ __align__(16) float M[64 * 32];
  for (int i = threadIdx.x; i < 64 * kT; i += 256) Y[i] = in[i];
  for (int i = threadIdx.x; i < 64 * 32; i += 256) M[i] = in[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tg = lane & 3, kg = lane >> 2;
  float2 acc[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) { acc[j][0] = make_float2(0.f, 0.f); acc[j][1] = make_float2(0.f, 0.f); }
  const float4 *mrow = reinterpret_cast<const float4 *>(M + kg * 4);
  const float4 *yrow = reinterpret_cast<const float4 *>(Y + (4 * warp + tg) * 4);
  for (int r = 0; r < reps; ++r) {
#pragma unroll 8
    for (int n = 0; n < 64; ++n) {
      const float4 m = mrow[n * 8], y = yrow[n * (kT / 4)];
      const float yv[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j][0] = __ffma2_rn(make_float2(yv[j], yv[j]), make_float2(m.x, m.y), acc[j][0]);
        acc[j][1] = __ffma2_rn(make_float2(yv[j], yv[j]), make_float2(m.z, m.w), acc[j][1]);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) s += acc[j][0].x + acc[j][0].y + acc[j][1].x + acc[j][1].y;
  out[blockIdx.x * 256 + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256, 2) mbC(const float *in, float *out, int reps) {
  float2 acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = make_float2(0.f, 0.f);
  float2 a = make_float2(in[threadIdx.x], in[threadIdx.x + 1]), b = make_float2(in[threadIdx.x + 2], in[threadIdx.x + 3]);
  for (int r = 0; r < reps; ++r) {
#pragma unroll 4
    for (int n = 0; n < 64; ++n) {
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[k] = __ffma2_rn(a, b, acc[k]);
    }
  }
  float s = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += acc[k].x + acc[k].y;
  out[blockIdx.x * 256 + threadIdx.x] = s;
}

// E: 8k x 8t register tile (64 accumulators), lanes = 4 kg x 8 tg; E2: 16k x 4t, lanes = 2 kg x 16 tg.  128 threads, 2 CTAs / SM
// (dynamic shared memory sized to force that), Y holds 32 n x 256 t.
template <int VAR> __global__ void __launch_bounds__(128, 2) mbE(const float *in, float *out, int reps) {
  extern __shared__ __align__(16) float sm[];
  float *Y = sm, *M = sm + 32 * 256;
  for (int i = threadIdx.x; i < 32 * 256; i += 128) Y[i] = in[i];
  for (int i = threadIdx.x; i < 64 * 32; i += 128) M[i] = in[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float s = 0;
  if (VAR == 0) {
    const int tg = lane & 7, kg = lane >> 3;
    float2 acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = make_float2(0.f, 0.f);
    const float4 *mrow = reinterpret_cast<const float4 *>(M + kg * 8);
    const float4 *yrow = reinterpret_cast<const float4 *>(Y + (8 * warp + tg) * 8);
    for (int r = 0; r < reps; ++r) {
#pragma unroll 4
      for (int n = 0; n < 32; ++n) {
        const float4 ma = mrow[n * 8], mb = mrow[n * 8 + 1], ya = yrow[n * 64], yb = yrow[n * 64 + 1];
        const float yv[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 yy = make_float2(yv[j], yv[j]);
          acc[j][0] = __ffma2_rn(yy, make_float2(ma.x, ma.y), acc[j][0]);
          acc[j][1] = __ffma2_rn(yy, make_float2(ma.z, ma.w), acc[j][1]);
          acc[j][2] = __ffma2_rn(yy, make_float2(mb.x, mb.y), acc[j][2]);
          acc[j][3] = __ffma2_rn(yy, make_float2(mb.z, mb.w), acc[j][3]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) s += acc[j][i].x + acc[j][i].y;
  } else {
    const int tg = lane & 15, kg = lane >> 4;
    float2 acc[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[j][i] = make_float2(0.f, 0.f);
    const float4 *mrow = reinterpret_cast<const float4 *>(M + kg * 16);
    const float4 *yrow = reinterpret_cast<const float4 *>(Y + (16 * warp + tg) * 4);
    for (int r = 0; r < reps; ++r) {
#pragma unroll 4
      for (int n = 0; n < 32; ++n) {
        const float4 m0 = mrow[n * 8], m1 = mrow[n * 8 + 1], m2 = mrow[n * 8 + 2], m3 = mrow[n * 8 + 3], y = yrow[n * 64];
        const float yv[4] = {y.x, y.y, y.z, y.w};
        const float mv[16] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w, m2.x, m2.y, m2.z, m2.w, m3.x, m3.y, m3.z, m3.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[j][i] = __ffma2_rn(make_float2(yv[j], yv[j]), make_float2(mv[2 * i], mv[2 * i + 1]), acc[j][i]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) s += acc[j][i].x + acc[j][i].y;
  }
  out[blockIdx.x * 128 + threadIdx.x] = s;
}

// F: 32k x 2t register tile (64 accumulators); the matrix comes from the constant bank as 128-bit uniform loads (LDCU.128 ->
// uniform registers -> FFMA2 UR operand), Y is one LDS.64 per n.  128 threads, 3 CTAs / SM.
__constant__ float4 cM4[64 * 8];
__global__ void __launch_bounds__(128, 3) mbF(const float *in, float *out, int reps) {
  extern __shared__ __align__(16) float sm[];
  float *Y = sm;
  for (int i = threadIdx.x; i < 32 * 256; i += 128) Y[i] = in[i];
  __syncthreads();
  float2 acc[2][16];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[j][i] = make_float2(0.f, 0.f);
  const float2 *yrow = reinterpret_cast<const float2 *>(Y) + threadIdx.x;
  for (int r = 0; r < reps; ++r) {
#pragma unroll 2
    for (int n = 0; n < 32; ++n) {
      const float2 y = yrow[n * 128];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 m = cM4[n * 8 + q];
        acc[0][2 * q] = __ffma2_rn(make_float2(y.x, y.x), make_float2(m.x, m.y), acc[0][2 * q]);
        acc[0][2 * q + 1] = __ffma2_rn(make_float2(y.x, y.x), make_float2(m.z, m.w), acc[0][2 * q + 1]);
        acc[1][2 * q] = __ffma2_rn(make_float2(y.y, y.y), make_float2(m.x, m.y), acc[1][2 * q]);
        acc[1][2 * q + 1] = __ffma2_rn(make_float2(y.y, y.y), make_float2(m.z, m.w), acc[1][2 * q + 1]);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[j][i].x + acc[j][i].y;
  out[blockIdx.x * 128 + threadIdx.x] = s;
}
static double g_warps = 148.0 * 2 * 8;
template <typename F> static void run(const char *name, F launch, double ffma2_per_thread) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaDeviceSynchronize();
  cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double winstr = g_warps * ffma2_per_thread;       // warp-level FFMA2 instructions
  const double peak = 148.0 * 4 * 0.5 * 1.965e9;                                // FFMA2 warp-instr / s at one per 2 clk per SMSP
  printf("%s: %.3f ms  %.1f%% of FFMA2 peak (%s)\n", name, ms, 100.0 * winstr / (ms * 1e-3) / peak, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  float *in, *out; cudaMalloc(&in, 64 * kT * 4 + 64);   // 32 KB + slack, also read as 32 x 256 by E
  cudaMalloc(&out, 148 * 2 * 256 * 4);
  cudaMemset(in, 0, 64 * kT * 4 + 64);
  float h[64 * 32]; for (int i = 0; i < 64 * 32; ++i) h[i] = 1.0f / (1 + i); cudaMemcpyToSymbol(cM, h, sizeof(h));
  const int reps = 400;
  run("A ldcu  (16 FFMA2 + 1 LDS.32 + uniform constant operands per n)", [&] { mbA<<<148 * 2, 256>>>(in, out, reps); }, 16.0 * 64 * reps);
  run("B lds128 (8 FFMA2 + 2 LDS.128 per n, 4x4 tile)", [&] { mbB<<<148 * 2, 256>>>(in, out, reps); }, 8.0 * 64 * reps);
  run("C registers only (16 FFMA2 per n)", [&] { mbC<<<148 * 2, 256>>>(in, out, reps); }, 16.0 * 64 * reps);
  g_warps = 148.0 * 2 * 4;
  const int smE = 100 * 1024;
  cudaFuncSetAttribute(mbE<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smE);
  cudaFuncSetAttribute(mbE<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smE);
  run("E  8k x 8t tile, 128 thr x 2 CTA/SM (32 FFMA2 + 4 LDS.128 per n)", [&] { mbE<0><<<148 * 2, 128, smE>>>(in, out, reps); }, 32.0 * 32 * reps);
  run("E2 16k x 4t tile, 128 thr x 2 CTA/SM (32 FFMA2 + 5 LDS.128 per n)", [&] { mbE<1><<<148 * 2, 128, smE>>>(in, out, reps); }, 32.0 * 32 * reps);
  g_warps = 148.0 * 3 * 4;
  const int smF = 70 * 1024;
  cudaFuncSetAttribute(mbF, cudaFuncAttributeMaxDynamicSharedMemorySize, smF);
  cudaMemcpyToSymbol(cM4, h, sizeof(h));
  run("F  32k x 2t tile, LDCU.128 matrix, 128 thr x 3 CTA/SM (32 FFMA2 + 8 LDCU.128 + 1 LDS.64 per n)", [&] { mbF<<<148 * 3, 128, smF>>>(in, out, reps); }, 32.0 * 32 * reps);
  g_warps = 148.0 * 3 * 4;
  cudaFuncSetAttribute(mbE<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smF);
  run("E2 again at 3 CTA/SM", [&] { mbE<1><<<148 * 3, 128, smF>>>(in, out, reps); }, 32.0 * 32 * reps);
  return 0;
}
