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
// plus the shared k_plan / k_sum_final / k_argmax_final.  Any X, Y, Th >= 3 (edge tiles are masked, halos
// wrap by modulo); chosen by the plan for float32 grids of at least 1024 cells per plane that the fused
// SMEM-resident kernel does not cover.
#include "common.cuh"

namespace {

constexpr int kT = 256;

// ------------------------------------------------------------------------------------------------
constexpr int kTK = 8;  // theta cells per thread in the two theta kernels (7-register window + 8 new loads)

__global__ void __launch_bounds__(kT) k_tl_theta(const float* __restrict__ P, float2* __restrict__ EI, int XY, int Th,
                                                 PcTables<float> tab) {
  const int p = blockIdx.x * kT + threadIdx.x;
  if (p >= XY) return;
  const size_t base = (size_t)blockIdx.y * Th * XY + p;
  const int k_lo = blockIdx.z * kTK, k_hi = min(Th, k_lo + kTK);
  const float e0 = tab.ge[3], e1 = tab.ge[2], e2 = tab.ge[1], e3 = tab.ge[0];
  const float i0 = tab.gi[3], i1 = tab.gi[2], i2 = tab.gi[1], i3 = tab.gi[0];
  float w0 = P[base + (size_t)modp(k_lo - 3, Th) * XY], w1 = P[base + (size_t)modp(k_lo - 2, Th) * XY];
  float w2 = P[base + (size_t)modp(k_lo - 1, Th) * XY], w3 = P[base + (size_t)k_lo * XY];
  float w4 = P[base + (size_t)((k_lo + 1) % Th) * XY], w5 = P[base + (size_t)((k_lo + 2) % Th) * XY];
#pragma unroll
  for (int kk = 0; kk < kTK; ++kk) {
    const int k = k_lo + kk;
    if (k >= k_hi) break;
    int kn = k + 3;
    kn -= kn >= Th ? Th : 0;
    const float w6 = P[base + (size_t)kn * XY];
    const float s1 = w2 + w4, s2 = w1 + w5, s3 = w0 + w6;
    const float e = fmaf(e0, w3, fmaf(e1, s1, fmaf(e2, s2, e3 * s3)));
    const float i = fmaf(i0, w3, fmaf(i1, s1, fmaf(i2, s2, i3 * s3)));
    EI[base + (size_t)k * XY] = make_float2(e, i);
    w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; w5 = w6;
  }
}

// ------------------------------------------------------------------------------------------------
constexpr int kYX = 32;           // tile edge
constexpr int kYXH = kYX + 6;     // with halo
constexpr int kInStride = kYXH + 1;   // 39: odd, so that lanes walking down rows hit distinct banks
constexpr int kMidStride = kYX + 1;   // 33

__global__ void __launch_bounds__(kT) k_tl_yx(const float2* __restrict__ EI, float* __restrict__ A,
                                              const float* __restrict__ gi, int X, int Y, int Th, PcTables<float> tab,
                                              float* __restrict__ part) {
  __shared__ float2 s_in[kYXH * kInStride];
  __shared__ float2 s_mid[kYXH * kMidStride];
  __shared__ float s_red[kT / 32];
  const int XY = X * Y;
  const int x0 = blockIdx.x * kYX, y0 = blockIdx.y * kYX;
  const int plane = blockIdx.z;  // b * Th + k
  const float2* src = EI + (size_t)plane * XY;
  const int tid = threadIdx.x;
  const int gx0 = modp(x0 - 3, X), gy0 = modp(y0 - 3, Y);
  const bool nowrapdiv = (X >= kYXH && Y >= kYXH);  // one conditional subtract wraps the index
  for (int i = tid; i < kYXH * kYXH; i += kT) {
    const int r = i / kYXH, c = i - r * kYXH;
    int gx = gx0 + r, gy = gy0 + c;
    if (nowrapdiv) {
      gx -= gx >= X ? X : 0;
      gy -= gy >= Y ? Y : 0;
    } else {
      gx %= X;
      gy %= Y;
    }
    s_in[r * kInStride + c] = src[(size_t)gx * Y + gy];
  }
  __syncthreads();
  // y pass: item = (segment of 8 outputs, halo row); lanes walk down the rows
  if (tid < 4 * kYXH) {
    const int seg = tid / kYXH, r = tid - seg * kYXH;
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
  // x pass: item = (segment of 8 outputs, column); lanes walk along y
  float psum = 0.f;
  if (tid < 4 * kYX) {
    const int seg = tid / kYX, y = tid - seg * kYX;
    float2 cf[7];
#pragma unroll
    for (int t = 0; t < 7; ++t) cf[t] = make_float2(tab.gex[t], tab.gix[t]);
    float2 in[14];
#pragma unroll
    for (int j = 0; j < 14; ++j) in[j] = s_mid[(seg * 8 + j) * kMidStride + y];
    const float g = gi[plane / Th];
    const int gy = y0 + y;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int t = 0; t < 7; ++t) acc = __ffma2_rn(in[jj + t], cf[t], acc);
      float a = acc.x - acc.y;
      a = (a < g) ? 0.f : a - g;  // posecell_network.py:339-340
      const int gx = x0 + seg * 8 + jj;
      if (gx < X && gy < Y) {
        A[(size_t)plane * XY + (size_t)gx * Y + gy] = a;
        psum += a;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
  if ((tid & 31) == 0) s_red[tid >> 5] = psum;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kT / 32; ++w) s += s_red[w];
    const int ntiles = gridDim.x * gridDim.y;
    const int k = plane % Th, b = plane / Th;
    part[(size_t)b * Th * ntiles + (size_t)k * ntiles + blockIdx.y * gridDim.x + blockIdx.x] = s;
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
  const bool nowrapdiv = (X >= k2XH && Y >= k2YH);
  for (int i = tid; i < k2XH * k2YH; i += kT) {
    const int r = i / k2YH, c = i - r * k2YH;
    int gx = gx0 + r, gy = gy0 + c;
    if (nowrapdiv) {
      gx -= gx >= X ? X : 0;
      gy -= gy >= Y ? Y : 0;
    } else {
      gx %= X;
      gy %= Y;
    }
    s_a[r * k2Stride + c] = src[(size_t)gx * Y + gy];
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
__global__ void __launch_bounds__(kT) k_tl_theta_fin(const float* __restrict__ Bp, float* __restrict__ S,
                                                     const int* __restrict__ ogi, int XY, int Th, PcTables<float> tab,
                                                     float* __restrict__ part_val, long long* __restrict__ part_idx) {
  __shared__ float s_v[kT / 32];
  __shared__ long long s_i[kT / 32];
  const int p = blockIdx.x * kT + threadIdx.x;
  float best = -INFINITY;
  long long bidx = 0x7fffffffffffffffLL;
  if (p < XY) {
    const size_t base = (size_t)blockIdx.y * Th * XY + p;
    const float* f = tab.f1d[ogi[blockIdx.y]];
    const float f0 = f[0], f1 = f[1], f2 = f[2], f3 = f[3], f4 = f[4], f5 = f[5], f6 = f[6];
    const int k_lo = blockIdx.z * kTK, k_hi = min(Th, k_lo + kTK);
    float w0 = Bp[base + (size_t)modp(k_lo - 3, Th) * XY], w1 = Bp[base + (size_t)modp(k_lo - 2, Th) * XY];
    float w2 = Bp[base + (size_t)modp(k_lo - 1, Th) * XY], w3 = Bp[base + (size_t)k_lo * XY];
    float w4 = Bp[base + (size_t)((k_lo + 1) % Th) * XY], w5 = Bp[base + (size_t)((k_lo + 2) % Th) * XY];
#pragma unroll
    for (int kk = 0; kk < kTK; ++kk) {
      const int k = k_lo + kk;
      if (k >= k_hi) break;
      int kn = k + 3;
      kn -= kn >= Th ? Th : 0;
      const float w6 = Bp[base + (size_t)kn * XY];
      float c = fmaf(f0, w0, fmaf(f1, w1, fmaf(f2, w2, fmaf(f3, w3, fmaf(f4, w4, fmaf(f5, w5, f6 * w6))))));
      c = (c < 0.f) ? 0.f : c;  // posecell_network.py:314
      S[base + (size_t)k * XY] = c;
      if (c > best) {  // theta ascending: strict '>' keeps the lowest flat index of this line
        best = c;
        bidx = (long long)p * Th + k;
      }
      w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; w5 = w6;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float v2 = __shfl_xor_sync(0xffffffffu, best, o);
    const long long i2 = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (v2 > best || (v2 == best && i2 < bidx)) {
      best = v2;
      bidx = i2;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    s_v[threadIdx.x >> 5] = best;
    s_i[threadIdx.x >> 5] = bidx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < kT / 32; ++w)
      if (s_v[w] > best || (s_v[w] == best && s_i[w] < bidx)) {
        best = s_v[w];
        bidx = s_i[w];
      }
    const size_t slot = ((size_t)blockIdx.y * gridDim.z + blockIdx.z) * gridDim.x + blockIdx.x;
    part_val[slot] = best;
    part_idx[slot] = bidx;
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
  int rc = prs_pc_launch_plan(p, odom, err, st);
  if (rc != PRS_OK) return rc;
  const int nline = (XY + kT - 1) / kT;
  const int nchunk = (Th + kTK - 1) / kTK;
  k_tl_theta<<<dim3(nline, B, nchunk), kT, 0, st>>>(state, EI, XY, Th, p->tf);
  const dim3 g2((X + kYX - 1) / kYX, (Y + kYX - 1) / kYX, B * Th);
  k_tl_yx<<<g2, kT, 0, st>>>(EI, A, gi, X, Y, Th, p->tf, (float*)p->part_val);
  rc = prs_pc_launch_sum_final_f32(p, Th * g2.x * g2.y, total, st);
  if (rc != PRS_OK) return rc;
  const dim3 g3((X + k2X - 1) / k2X, (Y + k2Y - 1) / k2Y, B * Th);
  k_tl_2d<<<g3, kT, 0, st>>>(A, Bp, p->shift, p->fsel, (const float*)p->inv_total, X, Y, Th, p->tf);
  k_tl_theta_fin<<<dim3(nline, B, nchunk), kT, 0, st>>>(Bp, state, p->ogi, XY, Th, p->tf, (float*)p->part_val,
                                                        p->part_idx);
  rc = prs_pc_launch_argmax_final_f32(p, nline * nchunk, argmax, st);
  if (rc != PRS_OK) return rc;
  PRS_CUDA(cudaGetLastError());
  return PRS_OK;
}
