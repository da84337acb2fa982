// Fused, SMEM-resident pose-cell update for reference-size networks (float32, sm_100a).
//
// One CTA owns one network at a time and performs the whole PoseCellNetwork.update()
// (ratslam/posecell_network.py:326-353) -- 7x7x7 DoG correlate, global inhibition, normalisation,
// per-heading shifted 7x7 correlate, 7-tap theta correlate, arg-max -- with the state held in shared
// memory between the stages; HBM sees one read and one write of the state per update (8 B / cell).
// The grid is persistent: min(B, #SM) CTAs stride over the B networks of an ensemble.
//
// Shared-memory plan (N = X*Y*Th cells):
//   buf2[N] float2      (E, I) pairs of the separable DoG, theta-major [th][x][y]        8N bytes
//     after the x pass the same bytes are reused as   bufA[N] float (inhibited activity A)
//                                                     bufB[N] float (path-integrated planes)
// Stages (one __syncthreads between each):
//   1 theta pass   thread = one (x,y) line of Th cells in registers, global -> buf2      11 op / cell
//   2 y pass       thread = one (th,x) line, in place, packed FFMA2 on (E,I)             7 FFMA2 / cell
//   3 x pass       thread = one (th,y) line, A = aE*E - aI*I, inhibit, block sum         7 FFMA2 / cell
//   4 2-D shift    thread = one x-row of TWO theta planes, packed FFMA2 over the plane pair;
//                  the integer x origin picks the source rows, the y origin rotates the store   24.5 FFMA2 / cell
//   5 theta pass   thread = one (x,y) line, clamp, arg-max, registers -> global          7 FFMA / cell
// All FMA-heavy stages use the sm_100 packed instruction (fma.rn.f32x2, SASS FFMA2): the FP32 pipe
// rate is the same as scalar FFMA (measured, bench_tools/microbench.cu) but it needs half the issue
// slots, which leaves room for the shared-memory loads.
#include "common.cuh"

namespace {

template <int X, int Y, int T>
struct ResLayout {
  static constexpr int XY = X * Y;
  static constexpr int N = XY * T;
  static constexpr size_t kTabOff = (size_t)8 * N;
  static constexpr size_t kIntOff = kTabOff + 1024;                 // ox[T], oy[T], fsel[T], misc[4]
  static constexpr size_t kRedOff = (kIntOff + (3 * T + 4) * 4 + 15) / 16 * 16;
  static constexpr size_t kPairOff = kRedOff + 32 * 8 + 32 * 4 + 16;  // long long[32], float[32], float[4]
  static constexpr size_t kBytes = kPairOff + 14 * 8;                 // float2 (ge,gi)[7], (gex,gix)[7]
};
static_assert(sizeof(PcTables<float>) <= 1024, "tables must fit their shared-memory slot");

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

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
  constexpr int XY = L::XY, N = L::N;
  static_assert(X >= 7 && Y >= 7 && T >= 3, "resident kernel needs X, Y >= 7");
  extern __shared__ __align__(16) unsigned char smem[];
  float2* buf2 = reinterpret_cast<float2*>(smem);
  float* bufA = reinterpret_cast<float*>(smem);
  float* bufB = bufA + N;
  const PcTables<float>* tab = reinterpret_cast<const PcTables<float>*>(smem + L::kTabOff);
  int* s_ox = reinterpret_cast<int*>(smem + L::kIntOff);
  int* s_oy = s_ox + T;
  int* s_fs = s_oy + T;
  int* s_misc = s_fs + T;
  long long* red_i = reinterpret_cast<long long*>(smem + L::kRedOff);
  float* red_f = reinterpret_cast<float*>(smem + L::kRedOff + 32 * 8);
  float* s_val = red_f + 32;
  // coefficient pairs for the packed FMAs, kept as float2 so that one LDS.64 fills an aligned register pair
  float2* s_cf_ty = reinterpret_cast<float2*>(smem + L::kPairOff);
  float2* s_cf_x = s_cf_ty + 7;
  const int tid = threadIdx.x;

  for (int i = tid; i < (int)(sizeof(PcTables<float>) / 4); i += NT)
    reinterpret_cast<float*>(smem + L::kTabOff)[i] = reinterpret_cast<const float*>(tab_g)[i];
  if (tid < 7) {
    s_cf_ty[tid] = make_float2(tab_g->ge[tid], tab_g->gi[tid]);
    s_cf_x[tid] = make_float2(tab_g->gex[tid], tab_g->gix[tid]);
  }
  __syncthreads();

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
        s_ox[tid] = (int)oxd;
        s_oy[tid] = (int)oyd;
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

      // ---- 1. theta pass of the separable DoG: global -> (E, I) pairs
      {
        const float e0 = tab->ge[3], e1 = tab->ge[2], e2 = tab->ge[1], e3 = tab->ge[0];
        const float i0 = tab->gi[3], i1 = tab->gi[2], i2 = tab->gi[1], i3 = tab->gi[0];
        for (int p = tid; p < XY; p += NT) {
          float in[T];
#pragma unroll
          for (int k = 0; k < T; ++k) in[k] = gst[k * XY + p];
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

      // ---- 2. y pass, in place on each (theta, x) line
      {
        float2 cf[7];
#pragma unroll
        for (int t = 0; t < 7; ++t) cf[t] = s_cf_ty[t];
        for (int ln = tid; ln < T * X; ln += NT) {
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
      __syncthreads();

      // ---- 3. x pass + global inhibition (posecell_network.py:339-340) + sum (:343)
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
            const int k = ln / Y, y = ln - k * Y;
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
      __syncthreads();  // every (E, I) pair has been consumed: the bytes become bufA / bufB
#pragma unroll
      for (int it = 0; it < IT3; ++it) {
        const int ln = tid + it * NT;
        if (ln < T * Y) {
          const int k = ln / Y, y = ln - k * Y;
#pragma unroll
          for (int x = 0; x < X; ++x) bufA[k * XY + x * Y + y] = keep[it][x];
        }
      }
      const float tot = block_sum_bcast<NT>(psum, red_f, s_val);  // contains the barrier that publishes bufA
      const float inv = (tot != 0.f) ? 1.f / tot : 1.f;           // posecell_network.py:344-345

      // ---- 4. per-plane shifted 7x7 correlate (convolution.py:320-340), two planes per thread
      {
        constexpr int NP = (T + 1) / 2;
        for (int item = tid; item < NP * X; item += NT) {
          const int kp = item / X, x = item - kp * X;
          const int k0 = 2 * kp;
          const bool two = (k0 + 1 < T);
          const int k1 = two ? k0 + 1 : k0;
          const int xb0 = modp(x + s_ox[k0] - 3, X), xb1 = modp(x + s_ox[k1] - 3, X);
          const float* F0 = tab->f2d[s_fs[k0]];
          const float* F1 = tab->f2d[s_fs[k1]];
          float2 acc[Y];
#pragma unroll
          for (int j = 0; j < Y; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll 1
          for (int a = 0; a < 7; ++a) {
            int xr0 = xb0 + a, xr1 = xb1 + a;
            xr0 -= (xr0 >= X) ? X : 0;
            xr1 -= (xr1 >= X) ? X : 0;
            const float* r0 = bufA + k0 * XY + xr0 * Y;
            const float* r1 = bufA + k1 * XY + xr1 * Y;
            float2 row[Y];
#pragma unroll
            for (int j = 0; j < Y; ++j) row[j] = make_float2(r0[j], r1[j]);
            float2 cf[7];
#pragma unroll
            for (int q = 0; q < 7; ++q) cf[q] = make_float2(F0[a * 7 + q], F1[a * 7 + q]);
#pragma unroll
            for (int j = 0; j < Y; ++j) {
#pragma unroll
              for (int q = 0; q < 7; ++q) acc[j] = ffma2(row[(j + q + Y - 3) % Y], cf[q], acc[j]);
            }
          }
          // acc[j] is the output for y = (j - oy) mod Y: the y origin becomes a rotation of the store
          const int yb0 = modp(-s_oy[k0], Y), yb1 = modp(-s_oy[k1], Y);
          float* o0 = bufB + k0 * XY + x * Y;
          float* o1 = bufB + k1 * XY + x * Y;
#pragma unroll
          for (int j = 0; j < Y; ++j) {
            int y0 = yb0 + j, y1 = yb1 + j;
            y0 -= (y0 >= Y) ? Y : 0;
            y1 -= (y1 >= Y) ? Y : 0;
            const float v0 = acc[j].x * inv, v1 = acc[j].y * inv;
            o0[y0] = (v0 < 0.f) ? 0.f : v0;  // posecell_network.py:300
            if (two) o1[y1] = (v1 < 0.f) ? 0.f : v1;
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
          float in[T];
#pragma unroll
          for (int k = 0; k < T; ++k) in[k] = bufB[k * XY + p];
#pragma unroll
          for (int k = 0; k < T; ++k) {
            float c = 0.f;
#pragma unroll
            for (int t = 0; t < 7; ++t) c = fmaf(fc[t], in[(k + t + T - 3) % T], c);
            c = (c < 0.f) ? 0.f : c;  // posecell_network.py:314
            gst[k * XY + p] = c;
            if (c > best) {  // k ascending, p ascending: strict '>' keeps the lowest flat index
              best = c;
              bidx = (long long)p * T + k;
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
