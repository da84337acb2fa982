// View-template library sweep (ratslam/view_templates.py:16-28,63-75) for sm_100a.
//
// One launch scores every stored 32x32 template against one query and leaves
// (min score << 32 | lowest index) in a single uint64 -- the value and the
// numpy.argmin tie-break the reference obtains with a Python list comprehension.
//
// uint8 (the ROS path, ros_simulate.py:100-101): the reference's ``abs(T - q)`` on
// uint8 arrays wraps modulo 256, so the score is  sum((T - q) mod 256).  Each lane
// owns one 4-byte column word of a template; the 16 (ref mode) or 32 (circular
// mode) query rows of that column word stay in registers for the whole sweep.
// A template row is loaded once (one 32-byte DRAM sector per 8 lanes) and used
// against every query row it is paired with; each pairing costs three integer ops
// per 4 bytes:   r = (a|H) - (q&~H)        (byte-wise subtract, no cross-byte borrow)
//                z = r ^ (~a&H) ^ (q&H)    (repair bit 7 of every byte)
//                acc = dp4a(z, 0x01010101, acc)
// The per-offset sums are reduced over the 8 lanes of a template with a halving
// butterfly, the minimum over offsets is taken, and the packed key is min-reduced
// warp -> block -> one atomicMin.
//
// float32 ("profiles", BASELINE config 5): one template per warp, lane == column,
// score = sum |T - q| accumulated in float32.
#include "common.cuh"

namespace {

constexpr int kVtThreads = 256;
constexpr unsigned kH = 0x80808080u;

template <int MODE>
struct VtShape;
template <>
struct VtShape<PRS_VT_MODE_REF> {  // stored rows 1..30 against query rows 8..23, o = t - s in [-7, 7]
  static constexpr int T0 = 1, T1 = 31, S0 = 8, NS = 16, NOFF = 15, NACC = 16;
  __host__ __device__ static constexpr int off(int t, int s) { return t - s + 7; }
  __host__ __device__ static constexpr bool valid(int t, int s) { return t - s >= -7 && t - s <= 7; }
};
template <>
struct VtShape<PRS_VT_MODE_CIRCULAR> {  // every stored row against every query row, o = (t - s) mod 32
  static constexpr int T0 = 0, T1 = 32, S0 = 0, NS = 32, NOFF = 32, NACC = 32;
  __host__ __device__ static constexpr int off(int t, int s) { return (t - s) & 31; }
  __host__ __device__ static constexpr bool valid(int, int) { return true; }
};

__device__ __forceinline__ unsigned ld_stream_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// Sum the per-offset accumulators over a group of LANES lanes (LANES a power of two, groups aligned),
// halving the number of live accumulators per lane at each step.  On return lane holds N_FINAL fully
// reduced accumulators acc[0..N_FINAL) whose offset indices are base + j.
template <int N, int M, typename V>
struct Butterfly {
  __device__ __forceinline__ static void run(V* acc, int lane, int& base) {
    if constexpr (M >= 1) {
      if constexpr (N > 1) {
        constexpr int h = N / 2;
        const bool up = (lane & M) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
          V send = up ? acc[i] : acc[i + h];
          V keep = up ? acc[i + h] : acc[i];
          acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, M);
        }
        base += up ? h : 0;
        Butterfly<h, M / 2, V>::run(acc, lane, base);
      } else {
        acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], M);
        Butterfly<1, M / 2, V>::run(acc, lane, base);
      }
    }
  }
};
template <int N, typename V>
struct Butterfly<N, 0, V> {
  __device__ __forceinline__ static void run(V*, int, int&) {}
};

constexpr int final_count(int n, int lanes) {
  while (lanes > 1 && n > 1) {
    n /= 2;
    lanes /= 2;
  }
  return n;
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long k) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long k2 = __shfl_xor_sync(0xffffffffu, k, o);
    k = k2 < k ? k2 : k;
  }
  return k;
}

__device__ __forceinline__ void block_min_to_global(unsigned long long key, unsigned long long* out) {
  __shared__ unsigned long long sm[kVtThreads / 32];
  key = warp_min_u64(key);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sm[w] = key;
  __syncthreads();
  if (w == 0) {
    key = l < kVtThreads / 32 ? sm[l] : ~0ull;
    key = warp_min_u64(key);
    if (l == 0 && key != ~0ull) atomicMin(out, key);
  }
}

// --------------------------------------------------------------------------- uint8
template <int MODE>
__global__ void __launch_bounds__(kVtThreads)
    k_vt_sweep_u8(const uint8_t* __restrict__ lib, long long n, const uint8_t* __restrict__ query, long long base_index,
                  unsigned long long* __restrict__ key_out, uint32_t* __restrict__ scores) {
  using S = VtShape<MODE>;
  const int lane = threadIdx.x & 31;
  const int w = lane & 7;   // column word of the 32-byte row
  const int tj = lane >> 3; // template within the group of four
  // query rows of this column word, pre-split for the byte-wise subtract
  unsigned qlo[S::NS], qhi[S::NS];
#pragma unroll
  for (int s = 0; s < S::NS; ++s) {
    unsigned q = reinterpret_cast<const unsigned*>(query)[(S::S0 + s) * 8 + w];
    qlo[s] = q & ~kH;
    qhi[s] = q & kH;
  }
  const long long n_groups = (n + 3) >> 2;
  const long long warp0 = ((long long)blockIdx.x * kVtThreads + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * kVtThreads) >> 5;
  unsigned long long best = ~0ull;
  for (long long g = warp0; g < n_groups; g += n_warps) {
    const long long ti = g * 4 + tj;
    const bool live = ti < n;
    const unsigned* tp = reinterpret_cast<const unsigned*>(lib) + (live ? ti : 0) * 256 + w;
    unsigned a[S::T1 - S::T0];
#pragma unroll
    for (int t = S::T0; t < S::T1; ++t) a[t - S::T0] = ld_stream_u32(tp + t * 8);
    unsigned acc[S::NACC];
#pragma unroll
    for (int i = 0; i < S::NACC; ++i) acc[i] = 0;
#pragma unroll
    for (int t = S::T0; t < S::T1; ++t) {
      const unsigned av = a[t - S::T0];
      const unsigned ahi = av | kH;
      const unsigned afix = ~av & kH;
#pragma unroll
      for (int s = 0; s < S::NS; ++s) {
        if (S::valid(t, S::S0 + s)) {
          unsigned z = (ahi - qlo[s]) ^ afix ^ qhi[s];
          acc[S::off(t, S::S0 + s)] = __dp4a(z, 0x01010101u, acc[S::off(t, S::S0 + s)]);
        }
      }
    }
    int obase = 0;
    Butterfly<S::NACC, 4, unsigned>::run(acc, lane, obase);
    constexpr int NF = final_count(S::NACC, 8);
    unsigned m = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < NF; ++j)
      if (obase + j < S::NOFF) m = min(m, acc[j]);
    m = min(m, __shfl_xor_sync(0xffffffffu, m, 4));
    m = min(m, __shfl_xor_sync(0xffffffffu, m, 2));
    m = min(m, __shfl_xor_sync(0xffffffffu, m, 1));
    if (live) {
      unsigned long long key = ((unsigned long long)m << 32) | (unsigned long long)(base_index + ti);
      best = key < best ? key : best;
      if (scores != nullptr && w == 0) scores[ti] = m;
    }
  }
  block_min_to_global(best, key_out);
}

// --------------------------------------------------------------------------- float32
template <int MODE>
__global__ void __launch_bounds__(kVtThreads)
    k_vt_sweep_f32(const float* __restrict__ lib, long long n, const float* __restrict__ query, long long base_index,
                   unsigned long long* __restrict__ key_out, float* __restrict__ scores) {
  using S = VtShape<MODE>;
  const int lane = threadIdx.x & 31;  // column
  float q[S::NS];
#pragma unroll
  for (int s = 0; s < S::NS; ++s) q[s] = query[(S::S0 + s) * 32 + lane];
  const long long warp0 = ((long long)blockIdx.x * kVtThreads + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * kVtThreads) >> 5;
  unsigned long long best = ~0ull;
  for (long long ti = warp0; ti < n; ti += n_warps) {
    const float* tp = lib + ti * 1024 + lane;
    float a[S::T1 - S::T0];
#pragma unroll
    for (int t = S::T0; t < S::T1; ++t) a[t - S::T0] = ld_stream_f32(tp + t * 32);
    float acc[S::NACC];
#pragma unroll
    for (int i = 0; i < S::NACC; ++i) acc[i] = 0.f;
#pragma unroll
    for (int t = S::T0; t < S::T1; ++t) {
#pragma unroll
      for (int s = 0; s < S::NS; ++s) {
        if (S::valid(t, S::S0 + s)) acc[S::off(t, S::S0 + s)] += fabsf(a[t - S::T0] - q[s]);
      }
    }
    int obase = 0;
    Butterfly<S::NACC, 16, float>::run(acc, lane, obase);
    float m = (obase < S::NOFF) ? acc[0] : INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
    unsigned long long key = ((unsigned long long)__float_as_uint(m) << 32) | (unsigned long long)(base_index + ti);
    best = key < best ? key : best;
    if (scores != nullptr && lane == 0) scores[ti] = m;
  }
  block_min_to_global(best, key_out);
}

// --------------------------------------------------------------------------- sub-sampling mask
__global__ void k_vt_extract_u8(const uint8_t* __restrict__ frame, int im_cols, int row_lo, int row_step, int col_lo,
                                int col_step, uint8_t* __restrict__ out, int n_rows, int n_cols) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * n_cols) return;
  int r = i / n_cols, c = i - r * n_cols;
  // n-th index strictly above lo with (idx - lo) % step != 0
  int rr = row_lo + (r / (row_step - 1)) * row_step + (r % (row_step - 1)) + 1;
  int cc = col_lo + (c / (col_step - 1)) * col_step + (c % (col_step - 1)) + 1;
  out[i] = frame[(size_t)rr * im_cols + cc];
}

int sweep_grid(long long units_per_warp_total) {
  // persistent-style grid: enough CTAs to fill 148 SMs a few times over, never more than the work
  long long warps = units_per_warp_total;
  long long blocks = (warps + (kVtThreads / 32) - 1) / (kVtThreads / 32);
  long long cap = 148LL * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

extern "C" int prs_vt_extract_u8(const uint8_t* frame, int im_rows, int im_cols, int row_lo, int row_hi, int row_step,
                                 int col_lo, int col_hi, int col_step, uint8_t* out, int n_rows, int n_cols,
                                 void* stream) {
  PRS_REQUIRE(frame && out, "prs_vt_extract_u8: null argument");
  PRS_REQUIRE(row_step >= 2 && col_step >= 2, "prs_vt_extract_u8: step must be >= 2 (step 1 selects nothing)");
  auto count = [](int lo, int hi, int step) { return (hi - lo - 1) - (hi - lo - 1) / step; };
  PRS_REQUIRE(row_lo >= 0 && col_lo >= 0 && row_hi <= im_rows && col_hi <= im_cols && row_hi > row_lo && col_hi > col_lo,
              "prs_vt_extract_u8: ranges outside the %dx%d frame", im_rows, im_cols);
  PRS_REQUIRE(count(row_lo, row_hi, row_step) == n_rows && count(col_lo, col_hi, col_step) == n_cols,
              "prs_vt_extract_u8: mask selects %dx%d pixels, not %dx%d", count(row_lo, row_hi, row_step),
              count(col_lo, col_hi, col_step), n_rows, n_cols);
  int total = n_rows * n_cols;
  k_vt_extract_u8<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(frame, im_cols, row_lo, row_step, col_lo,
                                                                         col_step, out, n_rows, n_cols);
  PRS_CUDA(cudaGetLastError());
  return PRS_OK;
}

extern "C" int prs_vt_sweep_u8(const uint8_t* lib, long long n, const uint8_t* query, int mode, long long base_index,
                               unsigned long long* key_out, uint32_t* scores, void* stream) {
  PRS_REQUIRE(query && key_out && n >= 0 && (lib || n == 0), "prs_vt_sweep_u8: bad argument");
  PRS_REQUIRE(mode == PRS_VT_MODE_REF || mode == PRS_VT_MODE_CIRCULAR, "prs_vt_sweep_u8: unknown mode %d", mode);
  PRS_REQUIRE(base_index >= 0 && base_index + n <= 0xffffffffLL, "prs_vt_sweep_u8: template index does not fit 32 bits");
  cudaStream_t st = (cudaStream_t)stream;
  PRS_CUDA(cudaMemsetAsync(key_out, 0xff, sizeof(unsigned long long), st));
  if (n == 0) return PRS_OK;
  int grid = sweep_grid((n + 3) / 4);
  if (mode == PRS_VT_MODE_REF)
    k_vt_sweep_u8<PRS_VT_MODE_REF><<<grid, kVtThreads, 0, st>>>(lib, n, query, base_index, key_out, scores);
  else
    k_vt_sweep_u8<PRS_VT_MODE_CIRCULAR><<<grid, kVtThreads, 0, st>>>(lib, n, query, base_index, key_out, scores);
  PRS_CUDA(cudaGetLastError());
  return PRS_OK;
}

extern "C" int prs_vt_sweep_f32(const float* lib, long long n, const float* query, int mode, long long base_index,
                                unsigned long long* key_out, float* scores, void* stream) {
  PRS_REQUIRE(query && key_out && n >= 0 && (lib || n == 0), "prs_vt_sweep_f32: bad argument");
  PRS_REQUIRE(mode == PRS_VT_MODE_REF || mode == PRS_VT_MODE_CIRCULAR, "prs_vt_sweep_f32: unknown mode %d", mode);
  PRS_REQUIRE(base_index >= 0 && base_index + n <= 0xffffffffLL, "prs_vt_sweep_f32: template index does not fit 32 bits");
  cudaStream_t st = (cudaStream_t)stream;
  PRS_CUDA(cudaMemsetAsync(key_out, 0xff, sizeof(unsigned long long), st));
  if (n == 0) return PRS_OK;
  int grid = sweep_grid(n);
  if (mode == PRS_VT_MODE_REF)
    k_vt_sweep_f32<PRS_VT_MODE_REF><<<grid, kVtThreads, 0, st>>>(lib, n, query, base_index, key_out, scores);
  else
    k_vt_sweep_f32<PRS_VT_MODE_CIRCULAR><<<grid, kVtThreads, 0, st>>>(lib, n, query, base_index, key_out, scores);
  PRS_CUDA(cudaGetLastError());
  return PRS_OK;
}

extern "C" int prs_vt_match_host_u8(const uint8_t* lib, long long n, const uint8_t* query_host, int mode,
                                    long long base_index, unsigned long long* key_host, void* scratch, void* stream) {
  PRS_REQUIRE(query_host && key_host && scratch, "prs_vt_match_host_u8: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* dq = (uint8_t*)scratch;
  unsigned long long* dk = (unsigned long long*)(dq + 1024);
  PRS_CUDA(cudaMemcpyAsync(dq, query_host, 1024, cudaMemcpyHostToDevice, st));
  int rc = prs_vt_sweep_u8(lib, n, dq, mode, base_index, dk, nullptr, st);
  if (rc != PRS_OK) return rc;
  PRS_CUDA(cudaMemcpyAsync(key_host, dk, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  PRS_CUDA(cudaStreamSynchronize(st));
  return PRS_OK;
}
