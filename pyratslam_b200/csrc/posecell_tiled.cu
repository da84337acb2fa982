// Tiled pose-cell update for large grids (float32), e.g. BASELINE config 3: one 256x256x72 network.
//
// The state of such a grid (18.9 MB) lives in L2, not in one SM's shared memory, so the update
// (ratslam/posecell_network.py:326-353) is four tiled kernels that stream it through L2 once each:
//
//   k_tl_theta      theta pass of the separable DoG: one thread per (x,y) line, a 7-register window slides
//                   along theta, every cell is read once;            P -> (E,I) float2        12 B / cell
//   k_tl_yx         y pass then x pass on a 32x32 tile (+3 halo) in shared memory, packed FFMA2 on (E,I),
//                   A = aE*E - aI*I, inhibition (:339-340), per-tile partial sums (:343)      8.5 + 4 B / cell
//   k_tl_2d         7x7 correlate (:273-274, convolution.py:320-340) on a 64x32 tile whose halo load is
//                   displaced by the plane's integer origin (the shift is free), 2x4 outputs per thread,
//                   rows arrive by LDS.128; * 1/total, clamp (:300)                            5.3 + 4 B / cell
//   k_tl_theta_fin  shifted 7-tap theta pass (convolution.py:344-359), clamp (:314), block arg-max  8 B / cell
//
// Four launches per update: the decisions (k_plan's arithmetic) run inside k_tl_theta, and the grid-wide sum
// and arg-max are finished by the last block of k_tl_yx / k_tl_theta_fin to retire (atomicInc counters that
// wrap back to zero, partials added in a fixed order).  Any X, Y, Th >= 3 (edge tiles are masked, halos
// wrap by modulo); chosen by the plan for float32 grids of at least 1024 cells per plane that the fused
// SMEM-resident kernel does not cover.
#include "common.cuh"

namespace {

constexpr int kT = 256;

// v in [-n, 2n): one conditional add/sub instead of an integer division (the kernels below are issue bound)
__device__ __forceinline__ int wrap_near(int v, int n) {
  v += v < 0 ? n : 0;
  v -= v >= n ? n : 0;
  return v;
}

// ------------------------------------------------------------------------------------------------
constexpr int kTK = 8;  // theta cells per thread in the two theta kernels (7-register window + 8 new loads)

struct PlanArgs {  // what prs_plan_cell needs; the theta kernel runs it for its network in one of its blocks
  const double *odom, *cos_th, *sin_th;
  double vtrans_scale, vrot_scale;
  int* shift;
  unsigned char* fsel;
  int *ogi, *err;
  int minXY;
};

__global__ void __launch_bounds__(kT) k_tl_theta(const float* __restrict__ P, float2* __restrict__ EI, int XY, int Th,
                                                 PcTables<float> tab, PlanArgs pa) {
  if (blockIdx.x == 0 && blockIdx.z == 0) {  // decisions of this update, needed from k_tl_2d on
    for (int k = threadIdx.x; k < Th; k += kT)
      prs_plan_cell(blockIdx.y, k, Th, pa.minXY, pa.odom, pa.cos_th, pa.sin_th, pa.vtrans_scale, pa.vrot_scale, pa.shift,
                    pa.fsel, pa.ogi, pa.err);
  }
  const int p = blockIdx.x * kT + threadIdx.x;
  if (p >= XY) return;
  const size_t base = (size_t)blockIdx.y * Th * XY + p;
  const float* Pb = P + base;
  float2* Eb = EI + base;
  const int k_lo = blockIdx.z * kTK;
  const bool near = Th >= kTK + 3;  // k_lo - 3 + j stays within [-Th, 2 Th)
  const float e0 = tab.ge[3], e1 = tab.ge[2], e2 = tab.ge[1], e3 = tab.ge[0];
  const float i0 = tab.gi[3], i1 = tab.gi[2], i2 = tab.gi[1], i3 = tab.gi[0];
  float w[kTK + 6];  // planes k_lo-3 .. k_lo+kTK+2, all loads issued before the first use
#pragma unroll
  for (int j = 0; j < kTK + 6; ++j) {
    const int kq = near ? wrap_near(k_lo - 3 + j, Th) : modp(k_lo - 3 + j, Th);
    w[j] = Pb[kq * XY];
  }
#pragma unroll
  for (int kk = 0; kk < kTK; ++kk) {
    const int k = k_lo + kk;
    if (k < Th) {
      const float s1 = w[kk + 2] + w[kk + 4], s2 = w[kk + 1] + w[kk + 5], s3 = w[kk] + w[kk + 6];
      const float e = fmaf(e0, w[kk + 3], fmaf(e1, s1, fmaf(e2, s2, e3 * s3)));
      const float i = fmaf(i0, w[kk + 3], fmaf(i1, s1, fmaf(i2, s2, i3 * s3)));
      Eb[k * XY] = make_float2(e, i);
    }
  }
}

// ------------------------------------------------------------------------------------------------
constexpr int kYXx = 56, kYXy = 32;      // outputs per tile (x rows, y columns)
constexpr int kYXrows = kYXx + 6;        // 62 halo rows
constexpr int kYXcols = kYXy + 6;        // 38 halo columns
constexpr int kInStride = kYXcols + 1;   // 39: odd, so that lanes walking down rows hit distinct banks
constexpr int kMidStride = kYXy + 1;     // 33

__global__ void __launch_bounds__(kT) k_tl_yx(const float2* __restrict__ EI, float* __restrict__ A,
                                              const float* __restrict__ gi, int X, int Y, int Th, PcTables<float> tab,
                                              float* __restrict__ part, unsigned* __restrict__ done_ctr,
                                              float* __restrict__ total, float* __restrict__ inv_total) {
  __shared__ float2 s_in[kYXrows * kInStride];
  __shared__ int s_last;
  __shared__ float2 s_mid[kYXrows * kMidStride];
  __shared__ float s_red[kT / 32];
  const int XY = X * Y;
  const int x0 = blockIdx.x * kYXx, y0 = blockIdx.y * kYXy;
  const int plane = blockIdx.z;  // b * Th + k
  const float2* src = EI + (size_t)plane * XY;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // halo tile: one warp per row, the row's wrapped x index is computed once, lanes walk along y
  {
    const int gy0 = modp(y0 - 3, Y);
    int gya = gy0 + lane, gyb = gy0 + lane + 32;
    gya = gya >= Y ? gya % Y : gya;
    gyb = gyb >= Y ? gyb % Y : gyb;
    constexpr int NR = (kYXrows + kT / 32 - 1) / (kT / 32);
    const bool xnear = X >= kYXrows;
    float2 va[NR], vb[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int r = wid + i * (kT / 32);
      if (r < kYXrows) {
        const int gx = xnear ? wrap_near(x0 - 3 + r, X) : modp(x0 - 3 + r, X);
        const float2* row = src + gx * Y;
        va[i] = row[gya];
        if (lane + 32 < kYXcols) vb[i] = row[gyb];
      }
    }
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int r = wid + i * (kT / 32);
      if (r < kYXrows) {
        s_in[r * kInStride + lane] = va[i];
        if (lane + 32 < kYXcols) s_in[r * kInStride + lane + 32] = vb[i];
      }
    }
  }
  __syncthreads();
  // y pass: item = (segment of 8 outputs, halo row); lanes walk down the rows (248 of 256 threads busy)
  if (tid < 4 * kYXrows) {
    const int seg = tid / kYXrows, r = tid - seg * kYXrows;
    float2 cf[7];
#pragma unroll
    for (int t = 0; t < 7; ++t) cf[t] = make_float2(tab.ge[t], tab.gi[t]);
    float2 in[14];
#pragma unroll
    for (int j = 0; j < 14; ++j) in[j] = s_in[r * kInStride + seg * 8 + j];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int t = 0; t < 7; ++t) acc = __ffma2_rn(in[jj + t], cf[t], acc);
      s_mid[r * kMidStride + seg * 8 + jj] = acc;
    }
  }
  __syncthreads();
  // x pass: item = (segment of 8 outputs, column); lanes walk along y (224 of 256 threads busy)
  float psum = 0.f;
  if (tid < (kYXx / 8) * kYXy) {
    const int seg = tid / kYXy, y = tid - seg * kYXy;
    float2 cf[7];
#pragma unroll
    for (int t = 0; t < 7; ++t) cf[t] = make_float2(tab.gex[t], tab.gix[t]);
    float2 in[14];
#pragma unroll
    for (int j = 0; j < 14; ++j) in[j] = s_mid[(seg * 8 + j) * kMidStride + y];
    const float g = gi[plane / Th];
    const int gy = y0 + y;
    float* Ap = A + (size_t)plane * XY;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int t = 0; t < 7; ++t) acc = __ffma2_rn(in[jj + t], cf[t], acc);
      float a = acc.x - acc.y;
      a = (a < g) ? 0.f : a - g;  // posecell_network.py:339-340
      const int gx = x0 + seg * 8 + jj;
      if (gx < X && gy < Y) {
        Ap[gx * Y + gy] = a;
        psum += a;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
  if (lane == 0) s_red[wid] = psum;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kT / 32; ++w) s += s_red[w];
    const int ntiles = gridDim.x * gridDim.y;
    const int k = plane % Th, b = plane / Th;
    part[(size_t)b * Th * ntiles + (size_t)k * ntiles + blockIdx.y * gridDim.x + blockIdx.x] = s;
    __threadfence();
    // the last block of this network to finish adds the partial sums up (fixed order: deterministic)
    s_last = (atomicInc(&done_ctr[b], (unsigned)(Th * ntiles - 1)) == (unsigned)(Th * ntiles - 1)) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    const int ntiles = gridDim.x * gridDim.y, np = Th * ntiles, b = plane / Th;
    const volatile float* pp = part + (size_t)b * np;
    float acc = 0.f;
    for (int i = tid; i < np; i += kT) acc += pp[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __syncthreads();
    if (lane == 0) s_red[wid] = acc;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kT / 32; ++w) t += s_red[w];
      total[b] = t;
      inv_total[b] = (t != 0.f) ? 1.f / t : 1.f;  // posecell_network.py:344-345
    }
  }
}

// ------------------------------------------------------------------------------------------------
constexpr int k2X = 64, k2Y = 32;          // outputs per tile
constexpr int k2XH = k2X + 6, k2YH = k2Y + 6;
constexpr int k2Stride = 40;               // floats per halo row, a multiple of 4 for LDS.128

__global__ void __launch_bounds__(kT) k_tl_2d(const float* __restrict__ A, float* __restrict__ Bp,
                                              const int* __restrict__ shift, const unsigned char* __restrict__ fsel,
                                              const float* __restrict__ inv_total, int X, int Y, int Th,
                                              PcTables<float> tab) {
  __shared__ __align__(16) float s_a[k2XH * k2Stride];
  const int XY = X * Y;
  const int x0 = blockIdx.x * k2X, y0 = blockIdx.y * k2Y;
  const int plane = blockIdx.z;
  const float* src = A + (size_t)plane * XY;
  const int tid = threadIdx.x;
  // the plane's integer origin displaces the tile that is loaded (convolution.py:329-331)
  const int gx0 = modp(x0 + shift[2 * plane] - 3, X), gy0 = modp(y0 + shift[2 * plane + 1] - 3, Y);
  {
    const int lane = tid & 31, wid = tid >> 5;
    int gya = gy0 + lane, gyb = gy0 + lane + 32;
    gya = gya >= Y ? gya % Y : gya;
    gyb = gyb >= Y ? gyb % Y : gyb;
    constexpr int NR = (k2XH + kT / 32 - 1) / (kT / 32);
    const bool xnear = X >= k2XH;
    float va[NR], vb[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int r = wid + i * (kT / 32);
      if (r < k2XH) {
        int gx = gx0 + r;
        gx = xnear ? (gx >= X ? gx - X : gx) : gx % X;
        const float* row = src + gx * Y;
        va[i] = row[gya];
        if (lane + 32 < k2YH) vb[i] = row[gyb];
      }
    }
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int r = wid + i * (kT / 32);
      if (r < k2XH) {
        s_a[r * k2Stride + lane] = va[i];
        if (lane + 32 < k2YH) s_a[r * k2Stride + lane + 32] = vb[i];
      }
    }
  }
  float F[49];
  {
    const float* f = tab.f2d[fsel[plane]];
#pragma unroll
    for (int i = 0; i < 49; ++i) F[i] = f[i];
  }
  __syncthreads();
  const int xb = tid >> 3, yb = tid & 7;
  const int x = 2 * xb, y = 4 * yb;
  float acc[2][4];
#pragma unroll
  for (int d = 0; d < 2; ++d)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[d][j] = 0.f;
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) {
    const float4* rp = reinterpret_cast<const float4*>(s_a + (x + rr) * k2Stride + y);
    const float4 v0 = rp[0], v1 = rp[1], v2 = rp[2];
    const float in[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
    if (rr <= 6) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 7; ++q) acc[0][j] = fmaf(in[j + q], F[rr * 7 + q], acc[0][j]);
    }
    if (rr >= 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 7; ++q) acc[1][j] = fmaf(in[j + q], F[(rr - 1) * 7 + q], acc[1][j]);
    }
  }
  const float inv = inv_total[plane / Th];
  const int gy = y0 + y;
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const int gx = x0 + x + d;
    if (gx >= X) continue;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float v = acc[d][j] * inv;
      o[j] = (v < 0.f) ? 0.f : v;  // posecell_network.py:300
    }
    float* dst = Bp + (size_t)plane * XY + (size_t)gx * Y + gy;
    if ((Y & 3) == 0 && gy + 3 < Y) {
      *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (gy + j < Y) dst[j] = o[j];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The same 7x7 stage on TWO adjacent theta planes at once with packed FFMA2 (the kernel above is bound by
// instruction issue: 49 FFMA per cell).  The two planes' tiles -- each displaced by its own origin -- are
// interleaved as float2 in shared memory, the coefficient pairs (F_k0[a][q], F_k1[a][q]) are read by LDS.128.
constexpr int k2pStride = 40;  // float2 per halo row

__global__ void __launch_bounds__(kT) k_tl_2d_pair(const float* __restrict__ A, float* __restrict__ Bp,
                                                   const int* __restrict__ shift, const unsigned char* __restrict__ fsel,
                                                   const float* __restrict__ inv_total, int X, int Y, int Th, int NPh,
                                                   PcTables<float> tab) {
  __shared__ __align__(16) float2 s_a[k2XH * k2pStride];
  __shared__ __align__(16) float2 s_cf[7 * 8];
  const int XY = X * Y;
  const int x0 = blockIdx.x * k2X, y0 = blockIdx.y * k2Y;
  const int b = blockIdx.z / NPh, kp = blockIdx.z - b * NPh;
  const int k0 = 2 * kp, k1 = (2 * kp + 1 < Th) ? 2 * kp + 1 : 2 * kp;
  const int pl0 = b * Th + k0, pl1 = b * Th + k1;
  const float* src0 = A + (size_t)pl0 * XY;
  const float* src1 = A + (size_t)pl1 * XY;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid < 56) {
    const int a = tid >> 3, q = tid & 7;
    s_cf[tid] = q < 7 ? make_float2(tab.f2d[fsel[pl0]][a * 7 + q], tab.f2d[fsel[pl1]][a * 7 + q]) : make_float2(0.f, 0.f);
  }
  {
    // each plane's integer origin displaces the tile that is loaded (convolution.py:329-331)
    const int gx00 = modp(x0 + shift[2 * pl0] - 3, X), gy00 = modp(y0 + shift[2 * pl0 + 1] - 3, Y);
    const int gx01 = modp(x0 + shift[2 * pl1] - 3, X), gy01 = modp(y0 + shift[2 * pl1 + 1] - 3, Y);
    int ya0 = gy00 + lane, yb0 = gy00 + lane + 32, ya1 = gy01 + lane, yb1 = gy01 + lane + 32;
    ya0 = ya0 >= Y ? ya0 % Y : ya0;
    yb0 = yb0 >= Y ? yb0 % Y : yb0;
    ya1 = ya1 >= Y ? ya1 % Y : ya1;
    yb1 = yb1 >= Y ? yb1 % Y : yb1;
    constexpr int NR = (k2XH + kT / 32 - 1) / (kT / 32);
    const bool xnear = X >= k2XH;
    float2 va[NR], vb[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int r = wid + i * (kT / 32);
      if (r < k2XH) {
        int g0 = gx00 + r, g1 = gx01 + r;
        g0 = xnear ? (g0 >= X ? g0 - X : g0) : g0 % X;
        g1 = xnear ? (g1 >= X ? g1 - X : g1) : g1 % X;
        const float* r0 = src0 + g0 * Y;
        const float* r1 = src1 + g1 * Y;
        va[i] = make_float2(r0[ya0], r1[ya1]);
        if (lane + 32 < k2YH) vb[i] = make_float2(r0[yb0], r1[yb1]);
      }
    }
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int r = wid + i * (kT / 32);
      if (r < k2XH) {
        s_a[r * k2pStride + lane] = va[i];
        if (lane + 32 < k2YH) s_a[r * k2pStride + lane + 32] = vb[i];
      }
    }
  }
  __syncthreads();
  const int xb = tid >> 3, yb = tid & 7;
  const int x = 2 * xb, y = 4 * yb;
  float2 acc[2][4];
#pragma unroll
  for (int d = 0; d < 2; ++d)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[d][j] = make_float2(0.f, 0.f);
  float2 cprev[7];
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) {
    const float4* rp = reinterpret_cast<const float4*>(s_a + (x + rr) * k2pStride + y);
    float2 in[12];
#pragma unroll
    for (int h = 0; h < 6; ++h) {
      const float4 v = rp[h];
      in[2 * h] = make_float2(v.x, v.y);
      in[2 * h + 1] = make_float2(v.z, v.w);
    }
    float2 ccur[7];
    if (rr <= 6) {
      const float4* cp = reinterpret_cast<const float4*>(s_cf + rr * 8);
      const float4 c01 = cp[0], c23 = cp[1], c45 = cp[2], c6x = cp[3];
      ccur[0] = make_float2(c01.x, c01.y);
      ccur[1] = make_float2(c01.z, c01.w);
      ccur[2] = make_float2(c23.x, c23.y);
      ccur[3] = make_float2(c23.z, c23.w);
      ccur[4] = make_float2(c45.x, c45.y);
      ccur[5] = make_float2(c45.z, c45.w);
      ccur[6] = make_float2(c6x.x, c6x.y);
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 7; ++q) acc[0][j] = __ffma2_rn(in[j + q], ccur[q], acc[0][j]);
    }
    if (rr >= 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 7; ++q) acc[1][j] = __ffma2_rn(in[j + q], cprev[q], acc[1][j]);
    }
#pragma unroll
    for (int q = 0; q < 7; ++q) cprev[q] = ccur[q];
  }
  const float inv = inv_total[b];
  const int gy = y0 + y;
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const int gx = x0 + x + d;
    if (gx >= X) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && k1 == k0) continue;  // odd Th: the last pair has one plane
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float v = (h == 0 ? acc[d][j].x : acc[d][j].y) * inv;
        o[j] = fmaxf(v, 0.f);  // posecell_network.py:300
      }
      float* dst = Bp + (size_t)(h == 0 ? pl0 : pl1) * XY + gx * Y + gy;
      if ((Y & 3) == 0 && gy + 3 < Y) {
        *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (gy + j < Y) dst[j] = o[j];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT) k_tl_theta_fin(const float* __restrict__ Bp, float* __restrict__ S,
                                                     const int* __restrict__ ogi, int XY, int Th, PcTables<float> tab,
                                                     float* __restrict__ part_val, long long* __restrict__ part_idx,
                                                     unsigned* __restrict__ done_ctr, long long* __restrict__ argmax) {
  __shared__ float s_v[kT / 32];
  __shared__ long long s_i[kT / 32];
  const int p = blockIdx.x * kT + threadIdx.x;
  float best = -INFINITY;
  int bidx = 0x7fffffff;  // the plan guarantees X*Y*Th < 2^31
  if (p < XY) {
    const size_t base = (size_t)blockIdx.y * Th * XY + p;
    const float* Bb = Bp + base;
    float* Sb = S + base;
    const float* f = tab.f1d[ogi[blockIdx.y]];
    const float f0 = f[0], f1 = f[1], f2 = f[2], f3 = f[3], f4 = f[4], f5 = f[5], f6 = f[6];
    const int k_lo = blockIdx.z * kTK;
    const bool near = Th >= kTK + 3;
    float w[kTK + 6];
#pragma unroll
    for (int j = 0; j < kTK + 6; ++j) {
      const int kq = near ? wrap_near(k_lo - 3 + j, Th) : modp(k_lo - 3 + j, Th);
      w[j] = Bb[kq * XY];
    }
    const int flat0 = p * Th + k_lo;
#pragma unroll
    for (int kk = 0; kk < kTK; ++kk) {
      const int k = k_lo + kk;
      if (k < Th) {
        float c = fmaf(f0, w[kk], fmaf(f1, w[kk + 1], fmaf(f2, w[kk + 2], fmaf(f3, w[kk + 3],
                  fmaf(f4, w[kk + 4], fmaf(f5, w[kk + 5], f6 * w[kk + 6]))))));
        c = fmaxf(c, 0.f);  // posecell_network.py:314
        Sb[k * XY] = c;
        if (c > best) {  // theta ascending: strict '>' keeps the lowest flat index of this line
          best = c;
          bidx = flat0 + kk;
        }
      }
    }
  }
  // block arg-max (first maximum in [x][y][th] order): values are >= 0, so their bit patterns order like the
  // values -- REDUX.MAX on the bits, then REDUX.MIN on the index among the lanes that hold the maximum
  {
    const unsigned vb = best >= 0.f ? __float_as_uint(best) : 0u;
    const int ib = best >= 0.f ? bidx : 0x7fffffff;
    const unsigned wmax = __reduce_max_sync(0xffffffffu, vb);
    const int widx = __reduce_min_sync(0xffffffffu, vb == wmax ? ib : 0x7fffffff);
    if ((threadIdx.x & 31) == 0) {
      s_v[threadIdx.x >> 5] = __uint_as_float(wmax);
      s_i[threadIdx.x >> 5] = widx;
    }
  }
  __syncthreads();
  __shared__ int s_last;
  const int nslots = gridDim.x * gridDim.z;
  if (threadIdx.x == 0) {
    float bv = s_v[0];
    long long bi = s_i[0];
#pragma unroll
    for (int w2 = 1; w2 < kT / 32; ++w2)
      if (s_v[w2] > bv || (s_v[w2] == bv && s_i[w2] < bi)) {
        bv = s_v[w2];
        bi = s_i[w2];
      }
    const size_t slot = (size_t)blockIdx.y * nslots + (size_t)blockIdx.z * gridDim.x + blockIdx.x;
    part_val[slot] = bv;
    part_idx[slot] = bi;
    __threadfence();
    s_last = (atomicInc(&done_ctr[blockIdx.y], (unsigned)(nslots - 1)) == (unsigned)(nslots - 1)) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {  // the last block of this network: final arg-max over the per-block candidates
    const volatile float* pv = part_val + (size_t)blockIdx.y * nslots;
    const volatile long long* pi = part_idx + (size_t)blockIdx.y * nslots;
    float bv = -1.f;
    long long bi = 0x7fffffffffffffffLL;
    for (int i = threadIdx.x; i < nslots; i += kT) {
      const float v = pv[i];
      const long long ix = pi[i];
      if (v > bv || (v == bv && ix < bi)) {
        bv = v;
        bi = ix;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float v2 = __shfl_xor_sync(0xffffffffu, bv, o);
      const long long i2 = __shfl_xor_sync(0xffffffffu, bi, o);
      if (v2 > bv || (v2 == bv && i2 < bi)) {
        bv = v2;
        bi = i2;
      }
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
      s_v[threadIdx.x >> 5] = bv;
      s_i[threadIdx.x >> 5] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int w2 = 1; w2 < kT / 32; ++w2)
        if (s_v[w2] > bv || (s_v[w2] == bv && s_i[w2] < bi)) {
          bv = s_v[w2];
          bi = s_i[w2];
        }
      argmax[blockIdx.y] = bi;
    }
  }
}

}  // namespace

int prs_pc_tiled_supported(const prs_pc_plan* p) {
  if (p->dtype != PRS_F32) return 0;
  if ((long long)p->B * p->Th > 65535) return 0;  // gridDim.z of the tile kernels
  return (p->X * p->Y >= 1024) ? 1 : 0;
}

int prs_pc_tiled_step(prs_pc_plan* p, float* state, const double* odom, const float* gi, long long* argmax, float* total,
                      int* err, cudaStream_t st) {
  const int X = p->X, Y = p->Y, Th = p->Th, B = p->B, XY = X * Y;
  float2* EI = (float2*)p->s1;  // s1|s2 are contiguous: 2*B*N floats
  float* A = (float*)p->s3;
  float* Bp = (float*)p->s4;
  const int nline = (XY + kT - 1) / kT;
  const int nchunk = (Th + kTK - 1) / kTK;
  const PlanArgs pa{odom, p->cos_th, p->sin_th, p->vtrans_scale, p->vrot_scale, p->shift, p->fsel, p->ogi, err,
                    X < Y ? X : Y};
  k_tl_theta<<<dim3(nline, B, nchunk), kT, 0, st>>>(state, EI, XY, Th, p->tf, pa);
  const dim3 g2((X + kYXx - 1) / kYXx, (Y + kYXy - 1) / kYXy, B * Th);
  k_tl_yx<<<g2, kT, 0, st>>>(EI, A, gi, X, Y, Th, p->tf, (float*)p->part_val, p->done_ctr, total, (float*)p->inv_total);
  const int NPh = (Th + 1) / 2;
  const dim3 g3p((X + k2X - 1) / k2X, (Y + k2Y - 1) / k2Y, B * NPh);
  if ((long long)g3p.x * g3p.y * g3p.z >= 2 * 148) {  // enough plane pairs to fill the chip: packed FFMA2 variant
    k_tl_2d_pair<<<g3p, kT, 0, st>>>(A, Bp, p->shift, p->fsel, (const float*)p->inv_total, X, Y, Th, NPh, p->tf);
  } else {
    const dim3 g3(g3p.x, g3p.y, B * Th);
    k_tl_2d<<<g3, kT, 0, st>>>(A, Bp, p->shift, p->fsel, (const float*)p->inv_total, X, Y, Th, p->tf);
  }
  k_tl_theta_fin<<<dim3(nline, B, nchunk), kT, 0, st>>>(Bp, state, p->ogi, XY, Th, p->tf, (float*)p->part_val,
                                                        p->part_idx, p->done_ctr + B, argmax);
  PRS_CUDA(cudaGetLastError());
  return PRS_OK;
}
