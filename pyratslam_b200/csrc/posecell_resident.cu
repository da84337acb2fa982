// Fused, SMEM-resident pose-cell update for reference-size networks (float32, sm_100a).
//
// One CTA owns one network at a time and performs the whole PoseCellNetwork.update()
// (ratslam/posecell_network.py:326-353) -- 7x7x7 DoG correlate, global inhibition, normalisation,
// per-heading shifted 7x7 correlate, 7-tap theta correlate, arg-max -- with the state held in shared
// memory between the stages; HBM sees one read and one write of the state per update (8 B / cell).
// The grid is persistent: min(B, #SM) CTAs stride over the B networks of an ensemble, and the next
// network's state is fetched by a TMA bulk copy (cp.async.bulk -> mbarrier) while the current one
// is being computed.
//
// Shared memory (N = X*Y*Th cells, NP = ceil(Th/2) plane pairs):
//   stage[N] float        the next network's state as it lies in HBM, theta-major [th][x][y]   4N bytes
//   buf2[N]  float2       (E, I) pairs of the separable DoG, [th][x][y]                        8N bytes
//     after the x pass the same bytes are reused as two plane-pair-interleaved tensors
//       A2[NP][x][y] float2 = (A'[kp], A'[kp+NP])   inhibited activity, already moved by the integer
//                                                   (x, y) origin of its plane
//       B2[NP][x][y] float2 = (B[kp],  B[kp+NP])    after the 2-D correlate
// Stages (a __syncthreads between each):
//   1 theta pass   thread = one (x,y) line of Th cells in registers, stage -> buf2         11 op / cell
//   2 y pass       thread = one (th,x) line, in place, packed FFMA2 on (E,I)               7 FFMA2 / cell
//   3 x pass       thread = one (th,y) line, A = aE*E - aI*I, inhibit, block sum; the store applies
//                  the plane's integer origin (a rotation of the x and y indices), so that ...
//   4 7x7 stage    ... is a plain periodic correlate: thread = one x-row of a plane PAIR, packed
//                  FFMA2 over the pair, rows and coefficient pairs arrive by LDS.64 / LDS.128   24.5 FFMA2 / cell
//   5 theta pass   thread = one (x,y) line, packed over the plane pair, clamp, arg-max, -> global   3.5 FFMA2 / cell
// The FMA-heavy stages use the sm_100 packed instruction (fma.rn.f32x2, SASS FFMA2): same FP32 pipe
// rate as scalar FFMA (measured, bench_tools/microbench.cu) for half the issue slots, which is what
// lets the shared-memory loads issue alongside.
#include "common.cuh"

namespace {

template <int X, int Y, int T>
struct ResLayout {
  static constexpr int XY = X * Y;
  static constexpr int N = XY * T;
  static constexpr int NP = (T + 1) / 2;
  static constexpr size_t kStageOff = 0;                                   // float[N]
  static constexpr size_t kBufOff = ((size_t)4 * N + 15) / 16 * 16;        // float2[max(N, 2*NP*XY)]
  static constexpr size_t kBufBytes = (size_t)8 * (2 * NP * XY > N ? 2 * NP * XY : N);
  static constexpr size_t kTabOff = (kBufOff + kBufBytes + 15) / 16 * 16;  // PcTables<float>, 1 KiB slot
  static constexpr size_t kPairOff = kTabOff + 1024;                       // float2[4][7][8] paired 2-D coefficients
  static constexpr size_t kCfOff = kPairOff + 4 * 7 * 8 * 8;               // float2[7] (ge,gi), float2[7] (gex,gix)
  static constexpr size_t kIntOff = kCfOff + 16 * 8;                       // int ox[T], oy[T], fsel[T], misc[4]
  static constexpr size_t kRedOff = (kIntOff + (3 * T + 4) * 4 + 15) / 16 * 16;  // long long[32], float[32], float[4]
  static constexpr size_t kBarOff = kRedOff + 32 * 8 + 32 * 4 + 16;        // mbarrier (8 bytes)
  static constexpr size_t kBytes = kBarOff + 16;
};
static_assert(sizeof(PcTables<float>) <= 1024, "tables must fit their shared-memory slot");

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// ---- mbarrier / TMA bulk copy (global -> shared), PTX ISA "cp.async.bulk"
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_%=;\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int NT>
__device__ __forceinline__ float block_sum_bcast(float v, float* red, float* out_slot) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float s = l < (NT + 31) / 32 ? red[l] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (l == 0) *out_slot = s;
  }
  __syncthreads();
  return *out_slot;
}

template <int X, int Y, int T, int NT>
__global__ void __launch_bounds__(NT, 1)
    k_pc_resident(float* state, const double* __restrict__ odom, int n_steps, const float* __restrict__ gi,
                  long long* __restrict__ argmax, float* __restrict__ total, int* __restrict__ err,
                  const double* __restrict__ cos_th, const double* __restrict__ sin_th, double vtrans_scale,
                  double vrot_scale, int B, const PcTables<float>* __restrict__ tab_g) {
  using L = ResLayout<X, Y, T>;
  constexpr int XY = L::XY, N = L::N, NP = L::NP;
  static_assert(X >= 7 && Y >= 7 && T >= 3, "resident kernel needs X, Y >= 7");
  static_assert((N * 4) % 16 == 0, "bulk copies move multiples of 16 bytes");
  extern __shared__ __align__(128) unsigned char smem[];
  float* stage = reinterpret_cast<float*>(smem + L::kStageOff);
  float2* buf2 = reinterpret_cast<float2*>(smem + L::kBufOff);
  float2* A2 = buf2;
  float2* B2 = buf2 + NP * XY;
  const PcTables<float>* tab = reinterpret_cast<const PcTables<float>*>(smem + L::kTabOff);
  float2* s_f2p = reinterpret_cast<float2*>(smem + L::kPairOff);  // [(fs0*2+fs1)*7 + a][8]
  float2* s_cf_ty = reinterpret_cast<float2*>(smem + L::kCfOff);
  float2* s_cf_x = s_cf_ty + 7;
  int* s_ox = reinterpret_cast<int*>(smem + L::kIntOff);
  int* s_oy = s_ox + T;
  int* s_fs = s_oy + T;
  int* s_misc = s_fs + T;
  long long* red_i = reinterpret_cast<long long*>(smem + L::kRedOff);
  float* red_f = reinterpret_cast<float*>(smem + L::kRedOff + 32 * 8);
  float* s_val = red_f + 32;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + L::kBarOff);
  const int tid = threadIdx.x;

  // ---- one-time set-up: tables, coefficient pairs, mbarrier, first prefetch
  for (int i = tid; i < (int)(sizeof(PcTables<float>) / 4); i += NT)
    reinterpret_cast<float*>(smem + L::kTabOff)[i] = reinterpret_cast<const float*>(tab_g)[i];
  if (tid < 7) {
    s_cf_ty[tid] = make_float2(tab_g->ge[tid], tab_g->gi[tid]);
    s_cf_x[tid] = make_float2(tab_g->gex[tid], tab_g->gix[tid]);
  }
  for (int i = tid; i < 4 * 7 * 8; i += NT) {
    const int q = i & 7, a = (i >> 3) % 7, combo = i / 56;
    s_f2p[i] = q < 7 ? make_float2(tab_g->f2d[combo >> 1][a * 7 + q], tab_g->f2d[combo & 1][a * 7 + q])
                     : make_float2(0.f, 0.f);
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_proxy_async();
    if ((int)blockIdx.x < B) {
      mbar_expect_tx(bar, N * 4);
      bulk_g2s(stage, state + (size_t)blockIdx.x * N, N * 4, bar);
    }
  }
  __syncthreads();
  uint32_t parity = 0;

  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    float* gst = state + (size_t)b * N;
    const float g_inh = gi[b];
    for (int step = 0; step < n_steps; ++step) {
      const double* od = odom + ((size_t)step * B + b) * 2;
      // ---- decisions of this step, float64 exactly as numpy computes them on the host
      //      (posecell_network.py:252-267,249,304)
      if (tid < T) {
        const double vt = __ddiv_rn(od[0], vtrans_scale);
        const double ex = __dmul_rn(vt, cos_th[tid]);
        const double ey = __dmul_rn(vt, sin_th[tid]);
        const double oxd = rint(ex), oyd = rint(ey);
        const int key = (int)__dmul_rn(__dsub_rn(ex, oxd), 10.0);
        s_ox[tid] = modp((int)oxd, X);
        s_oy[tid] = modp((int)oyd, Y);
        s_fs[tid] = key < 0 ? 1 : 0;
        int e = key >= 5 ? PRS_ERR_LUT_KEY : 0;
        if (tid == 0) {
          if (!(3.0 + ceil(fabs(vt)) <= (double)(X < Y ? X : Y))) e |= PRS_ERR_RADIUS;
          const double og = floor(__dadd_rn(__ddiv_rn(od[1], vrot_scale), 0.5));
          if (!(fabs(og) <= 64.0)) e |= PRS_ERR_THETA;
          const int ogc = og < -(double)PRS_OG_RANGE ? -PRS_OG_RANGE : (og > (double)PRS_OG_RANGE ? PRS_OG_RANGE : (int)og);
          s_misc[0] = ogc + PRS_OG_RANGE;
        }
        if (e) atomicOr(&err[b], e);
      }

      // ---- 1. theta pass of the separable DoG: stage (or global on later steps) -> (E, I) pairs
      if (step == 0) {
        mbar_wait(bar, parity);
        parity ^= 1;
      }
      {
        const float e0 = tab->ge[3], e1 = tab->ge[2], e2 = tab->ge[1], e3 = tab->ge[0];
        const float i0 = tab->gi[3], i1 = tab->gi[2], i2 = tab->gi[1], i3 = tab->gi[0];
        for (int p = tid; p < XY; p += NT) {
          float in[T];
          if (step == 0) {
#pragma unroll
            for (int k = 0; k < T; ++k) in[k] = stage[k * XY + p];
          } else {
#pragma unroll
            for (int k = 0; k < T; ++k) in[k] = gst[k * XY + p];
          }
#pragma unroll
          for (int k = 0; k < T; ++k) {
            const float c = in[k];
            const float s1 = in[(k + 1) % T] + in[(k + T - 1) % T];
            const float s2 = in[(k + 2) % T] + in[(k + T - 2) % T];
            const float s3 = in[(k + 3) % T] + in[(k + T - 3) % T];
            const float e = fmaf(e0, c, fmaf(e1, s1, fmaf(e2, s2, e3 * s3)));
            const float i = fmaf(i0, c, fmaf(i1, s1, fmaf(i2, s2, i3 * s3)));
            buf2[k * XY + p] = make_float2(e, i);
          }
        }
      }
      __syncthreads();
      // the staging buffer is free again: fetch the next network while this one is computed
      if (tid == 0 && step == 0) {
        const int nb = b + gridDim.x;
        if (nb < B) {
          fence_proxy_async();
          mbar_expect_tx(bar, N * 4);
          bulk_g2s(stage, state + (size_t)nb * N, N * 4, bar);
        }
      }

      // ---- 2. y pass, in place on each (theta, x) line
      {
        float2 cf[7];
#pragma unroll
        for (int t = 0; t < 7; ++t) cf[t] = s_cf_ty[t];
        constexpr int IT2 = (T * X + NT - 1) / NT;
#pragma unroll
        for (int it = 0; it < IT2; ++it) {
          const int ln = tid + it * NT;
          if (ln < T * X) {
            float2* line = buf2 + ln * Y;
            float2 in[Y];
#pragma unroll
            for (int y = 0; y < Y; ++y) in[y] = line[y];
#pragma unroll
            for (int y = 0; y < Y; ++y) {
              float2 acc = make_float2(0.f, 0.f);
#pragma unroll
              for (int t = 0; t < 7; ++t) acc = ffma2(in[(y + t + Y - 3) % Y], cf[t], acc);
              line[y] = acc;
            }
          }
        }
      }
      __syncthreads();

      // ---- 3. x pass + global inhibition (posecell_network.py:339-340) + sum (:343).
      //      Lines are enumerated theta-fastest so that a warp's loads stride by one plane (odd stride:
      //      no bank conflicts).
      constexpr int IT3 = (T * Y + NT - 1) / NT;
      float keep[IT3][X];
      float psum = 0.f;
      {
        float2 cf[7];
#pragma unroll
        for (int t = 0; t < 7; ++t) cf[t] = s_cf_x[t];
#pragma unroll
        for (int it = 0; it < IT3; ++it) {
          const int ln = tid + it * NT;
          if (ln < T * Y) {
            const int y = ln / T, k = ln - y * T;
            const float2* col = buf2 + k * XY + y;
            float2 in[X];
#pragma unroll
            for (int x = 0; x < X; ++x) in[x] = col[x * Y];
#pragma unroll
            for (int x = 0; x < X; ++x) {
              float2 acc = make_float2(0.f, 0.f);
#pragma unroll
              for (int t = 0; t < 7; ++t) acc = ffma2(in[(x + t + X - 3) % X], cf[t], acc);
              float a = acc.x - acc.y;
              a = (a < g_inh) ? 0.f : a - g_inh;
              keep[it][x] = a;
              psum += a;
            }
          }
        }
      }
      __syncthreads();  // every (E, I) pair has been consumed: the bytes become A2 / B2
      // Store A moved by the plane's integer origin: A'[x'][y'] = A[(x'+ox) % X][(y'+oy) % Y]
      // (convolution.py:320-340 adds origin_x[k], origin_y[k] to the read index; here it is subtracted
      // from the write index once), interleaved with the partner plane k +- NP.
#pragma unroll
      for (int it = 0; it < IT3; ++it) {
        const int ln = tid + it * NT;
        if (ln < T * Y) {
          const int y = ln / T, k = ln - y * T;
          const int half = k >= NP ? 1 : 0;
          const int kp = k - half * NP;
          int ys = y - s_oy[k];
          ys += ys < 0 ? Y : 0;
          int xs = X - s_ox[k];  // (0 - ox) mod X, in 1..X
          xs -= xs >= X ? X : 0;
          float* dst = reinterpret_cast<float*>(A2 + kp * XY + ys) + half;
#pragma unroll
          for (int x = 0; x < X; ++x) {
            dst[xs * (2 * Y)] = keep[it][x];
            xs = (xs + 1 == X) ? 0 : xs + 1;
          }
        }
      }
      const float tot = block_sum_bcast<NT>(psum, red_f, s_val);  // contains the barrier that publishes A2
      const float inv = (tot != 0.f) ? 1.f / tot : 1.f;           // posecell_network.py:344-345

      // ---- 4. 7x7 periodic correlate of both planes of a pair at once (posecell_network.py:273-274,300)
      {
        for (int item = tid; item < NP * X; item += NT) {
          const int kp = item / X, x = item - kp * X;
          const int k1 = (kp + NP < T) ? kp + NP : kp;
          const float2* ctab = s_f2p + (s_fs[kp] * 2 + s_fs[k1]) * 56;
          const float2* plane = A2 + kp * XY;
          float2 acc[Y];
#pragma unroll
          for (int j = 0; j < Y; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll
          for (int a = 0; a < 7; ++a) {
            int xr = x + a - 3;
            xr += xr < 0 ? X : 0;
            xr -= xr >= X ? X : 0;
            const float2* r = plane + xr * Y;
            float2 row[Y];
#pragma unroll
            for (int j = 0; j < Y; ++j) row[j] = r[j];
            const float4* cp = reinterpret_cast<const float4*>(ctab + a * 8);
            const float4 c01 = cp[0], c23 = cp[1], c45 = cp[2], c6x = cp[3];
            const float2 cf[7] = {make_float2(c01.x, c01.y), make_float2(c01.z, c01.w), make_float2(c23.x, c23.y),
                                  make_float2(c23.z, c23.w), make_float2(c45.x, c45.y), make_float2(c45.z, c45.w),
                                  make_float2(c6x.x, c6x.y)};
#pragma unroll
            for (int j = 0; j < Y; ++j) {
#pragma unroll
              for (int q = 0; q < 7; ++q) acc[j] = ffma2(row[(j + q + Y - 3) % Y], cf[q], acc[j]);
            }
          }
          float2* o = B2 + kp * XY + x * Y;
#pragma unroll
          for (int j = 0; j < Y; ++j) {
            const float v0 = acc[j].x * inv, v1 = acc[j].y * inv;
            o[j] = make_float2((v0 < 0.f) ? 0.f : v0, (v1 < 0.f) ? 0.f : v1);  // posecell_network.py:300
          }
        }
      }
      __syncthreads();

      // ---- 5. theta pass (convolution.py:344-359), clamp, arg-max, registers -> global
      float best = -INFINITY;
      long long bidx = 0x7fffffffffffffffLL;
      {
        float fc[7];
        const float* f1 = tab->f1d[s_misc[0]];
#pragma unroll
        for (int t = 0; t < 7; ++t) fc[t] = f1[t];
        for (int p = tid; p < XY; p += NT) {
          float2 pin[NP];
#pragma unroll
          for (int kk = 0; kk < NP; ++kk) pin[kk] = B2[kk * XY + p];
          if constexpr (T % 2 == 0) {
            // planes kk and kk+NP advance together; a tap that leaves [0, NP) lands in the partner half
            float2 cf2[7];
#pragma unroll
            for (int t = 0; t < 7; ++t) cf2[t] = make_float2(fc[t], fc[t]);
            float2 out[NP];
#pragma unroll
            for (int kk = 0; kk < NP; ++kk) {
              float2 acc = make_float2(0.f, 0.f);
#pragma unroll
              for (int t = 0; t < 7; ++t) {
                const int m = kk + t - 3;
                float2 v;
                if (m < 0)
                  v = make_float2(pin[m + NP].y, pin[m + NP].x);
                else if (m >= NP)
                  v = make_float2(pin[m - NP].y, pin[m - NP].x);
                else
                  v = pin[m];
                acc = ffma2(v, cf2[t], acc);
              }
              out[kk] = make_float2(acc.x < 0.f ? 0.f : acc.x, acc.y < 0.f ? 0.f : acc.y);  // posecell_network.py:314
            }
#pragma unroll
            for (int kk = 0; kk < NP; ++kk) {
              gst[kk * XY + p] = out[kk].x;
              if (out[kk].x > best) {  // k ascending within p, p ascending: strict '>' keeps the lowest flat index
                best = out[kk].x;
                bidx = (long long)p * T + kk;
              }
            }
#pragma unroll
            for (int kk = 0; kk < NP; ++kk) {
              gst[(kk + NP) * XY + p] = out[kk].y;
              if (out[kk].y > best) {
                best = out[kk].y;
                bidx = (long long)p * T + kk + NP;
              }
            }
          } else {
            float in[T];
#pragma unroll
            for (int kk = 0; kk < NP; ++kk) {
              in[kk] = pin[kk].x;
              if (kk + NP < T) in[kk + NP] = pin[kk].y;
            }
#pragma unroll
            for (int k = 0; k < T; ++k) {
              float c = 0.f;
#pragma unroll
              for (int t = 0; t < 7; ++t) c = fmaf(fc[t], in[(k + t + T - 3) % T], c);
              c = (c < 0.f) ? 0.f : c;
              gst[k * XY + p] = c;
              if (c > best) {
                best = c;
                bidx = (long long)p * T + k;
              }
            }
          }
        }
      }
      // block arg-max: value descending, reference flat index ascending (numpy.argmax)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float v2 = __shfl_xor_sync(0xffffffffu, best, o);
        const long long i2 = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (v2 > best || (v2 == best && i2 < bidx)) {
          best = v2;
          bidx = i2;
        }
      }
      {
        const int w = tid >> 5, l = tid & 31;
        if (l == 0) {
          red_f[w] = best;
          red_i[w] = bidx;
        }
        __syncthreads();
        if (w == 0) {
          float v = l < (NT + 31) / 32 ? red_f[l] : -INFINITY;
          long long ix = l < (NT + 31) / 32 ? red_i[l] : 0x7fffffffffffffffLL;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float v2 = __shfl_xor_sync(0xffffffffu, v, o);
            const long long i2 = __shfl_xor_sync(0xffffffffu, ix, o);
            if (v2 > v || (v2 == v && i2 < ix)) {
              v = v2;
              ix = i2;
            }
          }
          if (l == 0) {
            argmax[(size_t)step * B + b] = ix;
            total[(size_t)step * B + b] = tot;
          }
        }
      }
      __syncthreads();  // state in global and every SMEM slot are consistent before the next step / network
    }
  }
}

template <int X, int Y, int T, int NT>
int launch(prs_pc_plan* p, float* state, const double* odom, int n_steps, const float* gi, long long* argmax,
           float* total, int* err, cudaStream_t st) {
  using L = ResLayout<X, Y, T>;
  auto kern = k_pc_resident<X, Y, T, NT>;
  static bool configured = false;
  if (!configured) {
    PRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kBytes));
    configured = true;
  }
  int dev = 0, nsm = 148;
  PRS_CUDA(cudaGetDevice(&dev));
  PRS_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  const int grid = p->B < nsm ? p->B : nsm;
  kern<<<grid, NT, L::kBytes, st>>>(state, odom, n_steps, gi, argmax, total, err, p->cos_th, p->sin_th, p->vtrans_scale,
                                    p->vrot_scale, p->B, (const PcTables<float>*)p->tab_dev);
  PRS_CUDA(cudaGetLastError());
  return PRS_OK;
}

}  // namespace

int prs_pc_resident_supported(const prs_pc_plan* p) {
  if (p->dtype != PRS_F32) return 0;
  if (p->X == 21 && p->Y == 21 && p->Th == 36) return 1;
  return 0;
}

int prs_pc_resident_step(prs_pc_plan* p, void* state, const double* odom, int T, const void* gi, long long* argmax,
                         void* total, int* err, cudaStream_t st) {
  if (p->X == 21 && p->Y == 21 && p->Th == 36)
    return launch<21, 21, 36, 448>(p, (float*)state, odom, T, (const float*)gi, argmax, (float*)total, err, st);
  prs_set_error("resident path not available for this plan");
  return PRS_E_INVALID;
}
