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
//
// Those two are the row-major kernels.  What the product runs on big libraries (further down in this file):
//   k_vt_sweep_packed_ref_ring   bit-sliced uint8 library, reference mode: 8 LOP3 + 1 POPC per 32 byte-compares,
//                                warp-private rings of TMA bulk copies, the warps of a CTA in lock step
//   k_vt_sweep_packed_circ       the same library, all 32 cyclic shifts (counters rotated in registers)
//   k_vt_sweep_f32_pair          float32 library: two columns per lane (add.f32x2), two templates per warp, rings
//   k_vt_sweep_packed_ref_small  small libraries inside replayed frames: one block per group of 32 templates
// and the frame entry points (prs_frame_*, prs_replay_run) that chain them with the pose-cell update.
#include <new>
#include <string.h>

#include <atomic>
#include <mutex>

#include "common.cuh"
#include "vt_pack.cuh"

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

// Tuning knobs of the sweeps (prs_vt_tune): [0] ring depth of the packed reference-mode sweep (0 = the
// register-prefetch kernel; 2, 4, 8 = slots per warp; 34 = 4 slots with the CTA's warps in lock step, 44 = the same
// with one 640-thread CTA per SM), [1] CTAs per SM its grid is sized for, [2] ring depth of the float32 reference-mode
// sweep (0 = the register kernel, 1..4 = one template per warp, 11..13 = the column-pair kernel with depth - 10 slots),
// [3] CTAs per SM of that ring sweep.
static std::atomic<int> g_vt_knob[4] = {{34}, {5}, {13}, {2}};  // measured best on B200 (bench_tools/vt_tune.py)
static int launch_f32_ring(int depth, int ctas_per_sm, const float* lib, long long n, const float* query,
                           long long base_index, unsigned long long* key_out, float* scores, cudaStream_t st);

// the boolean mask of view_templates.py:48-57 must select exactly n_rows x n_cols pixels inside the frame
static int check_mask(int im_rows, int im_cols, int row_lo, int row_hi, int row_step, int col_lo, int col_hi, int col_step,
                      int n_rows, int n_cols) {
  PRS_REQUIRE(row_step >= 2 && col_step >= 2, "prs_vt_extract_u8: step must be >= 2 (step 1 selects nothing)");
  auto count = [](int lo, int hi, int step) { return (hi - lo - 1) - (hi - lo - 1) / step; };
  PRS_REQUIRE(row_lo >= 0 && col_lo >= 0 && row_hi <= im_rows && col_hi <= im_cols && row_hi > row_lo && col_hi > col_lo,
              "prs_vt_extract_u8: ranges outside the %dx%d frame", im_rows, im_cols);
  PRS_REQUIRE(count(row_lo, row_hi, row_step) == n_rows && count(col_lo, col_hi, col_step) == n_cols,
              "prs_vt_extract_u8: mask selects %dx%d pixels, not %dx%d", count(row_lo, row_hi, row_step),
              count(col_lo, col_hi, col_step), n_rows, n_cols);
  return PRS_OK;
}

extern "C" int prs_vt_extract_u8(const uint8_t* frame, int im_rows, int im_cols, int row_lo, int row_hi, int row_step,
                                 int col_lo, int col_hi, int col_step, uint8_t* out, int n_rows, int n_cols,
                                 void* stream) {
  PRS_REQUIRE(frame && out, "prs_vt_extract_u8: null argument");
  if (int rc = check_mask(im_rows, im_cols, row_lo, row_hi, row_step, col_lo, col_hi, col_step, n_rows, n_cols)) return rc;
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
  if (mode == PRS_VT_MODE_REF && g_vt_knob[2] != 0)
    return launch_f32_ring(g_vt_knob[2], g_vt_knob[3], lib, n, query, base_index, key_out, scores, st);
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

// =============================================================================================
// Bit-sliced ("packed") uint8 library.
//
// The byte-wise SWAR sweep above is bound by the integer pipes (64 lanes/clk/SM on B200, measured):
// 3 instructions per 4 byte-compares = 0.50 of the HBM roof at best.  The reference score only needs
//     sum((a - b) mod 256) = sum(a) - sum(b) + 256 * #{a < b}            (checked in the CPU tests)
// so the library is stored as bit planes: for every template row, 8 words, word k holding bit k of the
// row's 32 pixels.  #{a < b} over a row pair is then 8 LOP3 (one per plane, LSB to MSB:
// lt = (~a & b) | (~(a ^ b) & lt)) and one POPC for 32 pixels -- 10 instructions per 32 compares instead
// of 24.  One lane owns one template, 32 templates are interleaved so that a warp's 128-bit loads are
// contiguous 512-byte runs, and the query planes are __constant__ operands of the LOP3s (no registers,
// no loads).  Row sums (uint16 x 32) ride along: 1088 bytes per template.
//
// Layout of one group of 32 templates (8704 uint32):
//   planes : uint4 [32 rows][2 halves][32 lanes]   .x..w = planes 4h..4h+3 of that row of template `lane`
//   rowsum : uint4 [4][32 lanes]                   16 words per lane, word w = R[2w] | R[2w+1] << 16
constexpr int kGroupU4 = kVtGroupU4;  // 2176 uint4 = 34816 bytes per 32 templates (vt_pack.cuh)

__constant__ uint32_t c_vtq[32 * 8 + 8];  // query planes [row][k], then sum(rows 8..23), sum(all rows)

// c_vtq is ONE buffer per device, and the sweeps need the planes as constant-bank operands of their LOP3s (read from
// anywhere else they cost registers the kernels do not have).  Sweeps issued on different streams -- or from different
// host threads -- of a device are therefore ordered through it: a sweep records an event behind its kernel, and the
// next writer of the buffer on another stream waits for that event on the device (no host blocking).  The host mutex
// is held from the wait to the record, so two threads cannot interleave their (write planes, sweep) pairs either.
struct VtqGuard {
  std::mutex mu;
  cudaEvent_t ev[64] = {};
  cudaStream_t last[64] = {};
  bool pending[64] = {};
};
VtqGuard g_vtq;

struct VtqScope {
  cudaStream_t st = nullptr;
  int dev = -1;
  bool locked = false;
  // `st` must not be capturing (a graph that contains a sweep is bracketed by its launcher, prs_frame_launch)
  int begin(cudaStream_t s) {
    st = s;
    PRS_CUDA(cudaGetDevice(&dev));
    PRS_REQUIRE(dev >= 0 && dev < 64, "view-template sweep: device index %d out of range", dev);
    g_vtq.mu.lock();
    locked = true;
    if (!g_vtq.ev[dev]) PRS_CUDA(cudaEventCreateWithFlags(&g_vtq.ev[dev], cudaEventDisableTiming));
    if (g_vtq.pending[dev] && g_vtq.last[dev] != st) PRS_CUDA(cudaStreamWaitEvent(st, g_vtq.ev[dev], 0));
    return PRS_OK;
  }
  int end() {
    if (!locked) return PRS_OK;
    cudaError_t e = cudaEventRecord(g_vtq.ev[dev], st);
    if (e == cudaSuccess) {
      g_vtq.pending[dev] = true;
      g_vtq.last[dev] = st;
    }
    locked = false;
    g_vtq.mu.unlock();
    PRS_CUDA(e);
    return PRS_OK;
  }
  ~VtqScope() {
    if (locked) g_vtq.mu.unlock();
  }
};

__global__ void k_vt_pack_u8(const uint8_t* __restrict__ src, long long n, uint4* __restrict__ packed,
                             long long first) {
  // one thread per (template, row)
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 32) return;
  const long long tl = i >> 5;
  const int t = (int)(i & 31);
  vt_pack_row(src + tl * 1024, packed, first + tl, t);
}

__global__ void k_vt_unpack_u8(const uint4* __restrict__ packed, long long ti, uint8_t* __restrict__ dst) {
  // 1024 threads: one per pixel of template ti
  const int t = threadIdx.x >> 5, c = threadIdx.x & 31;
  const uint4* grp = packed + (ti >> 5) * kGroupU4;
  const int lane = (int)(ti & 31);
  const uint4 lo = grp[(t * 2 + 0) * 32 + lane], hi = grp[(t * 2 + 1) * 32 + lane];
  const uint32_t pl[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
  uint32_t px = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) px |= ((pl[k] >> c) & 1u) << k;
  dst[t * 32 + c] = (uint8_t)px;
}

__global__ void k_vt_pack_query(const uint8_t* __restrict__ q, uint32_t* __restrict__ out,
                                unsigned long long* __restrict__ key_init = nullptr) {
  // 32 threads: thread t packs row t; then the two sums.  key_init: the sweep's packed key starts at "nothing found"
  // (one graph node less than a memset of its own on the sharded query chain)
  const int t = threadIdx.x;
  if (t == 0 && key_init != nullptr) *key_init = ~0ull;
  uint32_t pl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t sum = 0;
  for (int c = 0; c < 32; ++c) {
    const uint32_t px = q[t * 32 + c];
    sum += px;
#pragma unroll
    for (int k = 0; k < 8; ++k) pl[k] |= ((px >> k) & 1u) << c;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) out[t * 8 + k] = pl[k];
  uint32_t mid = (t >= 8 && t < 24) ? sum : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mid += __shfl_xor_sync(0xffffffffu, mid, o);
  }
  if (t == 0) {
    out[256] = mid;
    out[257] = sum;
  }
}

__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

// #{pixels of the stored row that are smaller than the query row's}: bit planes (lo, hi) of the stored
// row against query row s, whose planes are constant-bank operands.  LSB to MSB; the last differing bit wins.
// One borrow step of a < q on bit plane k, (~a & q) | (~(a ^ q) & lt), is ONE three-input LOP3 (truth table 0x8e for
// inputs a, q, lt).  Written as inline PTX: left to itself the compiler re-associates the first two steps of 90 of the
// 240 chains into three LOP3s (2 182 LOP3 per 32 templates instead of 1 942).
__device__ __forceinline__ uint32_t lt_step(uint32_t a, uint32_t q, uint32_t lt) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0x8e;" : "=r"(r) : "r"(a), "r"(q), "r"(lt));
  return r;
}
__device__ __forceinline__ uint32_t lt_row(const uint4& lo, const uint4& hi, int s) {
  const uint32_t* q = c_vtq + s * 8;
  uint32_t lt = lt_step(lo.x, q[0], 0u);
  lt = lt_step(lo.y, q[1], lt);
  lt = lt_step(lo.z, q[2], lt);
  lt = lt_step(lo.w, q[3], lt);
  lt = lt_step(hi.x, q[4], lt);
  lt = lt_step(hi.y, q[5], lt);
  lt = lt_step(hi.z, q[6], lt);
  lt = lt_step(hi.w, q[7], lt);
  return (uint32_t)__popc(lt);
}

// one stored row T against every query row it is paired with in reference mode (o = T - s in [-7, 7], s in [8, 23])
template <int T>
__device__ __forceinline__ void ref_row(const uint4& lo, const uint4& hi, uint32_t (&cnt)[15]) {
  constexpr int s_lo = (T - 7 > 8) ? T - 7 : 8;
  constexpr int s_hi = (T + 7 < 23) ? T + 7 : 23;
#pragma unroll
  for (int s = s_lo; s <= s_hi; ++s) cnt[T - s + 7] += lt_row(lo, hi, s);
}

template <int T>
struct RefRows {
  __device__ __forceinline__ static void run(const uint4* gp, uint4 lo, uint4 hi, uint32_t (&cnt)[15]) {
    if constexpr (T <= 30) {
      uint4 nlo = lo, nhi = hi;
      if constexpr (T < 30) {  // the next row is in flight while this one is compared
        nlo = ld_stream_u4(gp + ((T + 1) * 2 + 0) * 32);
        nhi = ld_stream_u4(gp + ((T + 1) * 2 + 1) * 32);
      }
      ref_row<T>(lo, hi, cnt);
      RefRows<T + 1>::run(gp, nlo, nhi, cnt);
    }
  }
};

constexpr int kPkThreads = 128;

// Reference mode (15 windowed row offsets, view_templates.py:16-28) over the packed library.
__global__ void __launch_bounds__(kPkThreads)  // 96 registers (15 interleaved compare chains); capping them spills
    k_vt_sweep_packed_ref(const uint4* __restrict__ packed, long long n, long long base_index,
                          unsigned long long* __restrict__ key_out, uint32_t* __restrict__ scores,
                          const int* __restrict__ n_dev) {
  if (n_dev != nullptr) n = *n_dev;  // graph replays: the library size lives on the device
  const int lane = threadIdx.x & 31;
  const long long n_groups = (n + 31) >> 5;
  const long long warp0 = ((long long)blockIdx.x * kPkThreads + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * kPkThreads) >> 5;
  unsigned long long best = ~0ull;
  for (long long g = warp0; g < n_groups; g += n_warps) {
    const uint4* gp = packed + g * kGroupU4 + lane;
    uint32_t cnt[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) cnt[i] = 0;
    RefRows<1>::run(gp, ld_stream_u4(gp + (1 * 2 + 0) * 32), ld_stream_u4(gp + (1 * 2 + 1) * 32), cnt);
    // window sums of the stored rows: A(o) = sum_{r=8+o}^{23+o} R[r]
    uint32_t R[32];
#pragma unroll
    for (int w4 = 0; w4 < 4; ++w4) {
      const uint4 v = ld_stream_u4(gp + 32 * 2 * 32 + w4 * 32);
      const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        R[(w4 * 4 + j) * 2] = ww[j] & 0xffffu;
        R[(w4 * 4 + j) * 2 + 1] = ww[j] >> 16;
      }
    }
    uint32_t A = 0;
#pragma unroll
    for (int r = 1; r <= 16; ++r) A += R[r];
    const uint32_t bq = c_vtq[256];
    uint32_t m = 0xffffffffu;
#pragma unroll
    for (int o = -7; o <= 7; ++o) {
      const uint32_t sc = A + 256u * cnt[o + 7] - bq;
      m = min(m, sc);
      if (o < 7) A = A - R[8 + o] + R[24 + o];
    }
    const long long ti = g * 32 + lane;
    if (ti < n) {
      const unsigned long long key = ((unsigned long long)m << 32) | (unsigned long long)(base_index + ti);
      best = key < best ? key : best;
      if (scores != nullptr) scores[ti] = m;
    }
  }
  __shared__ unsigned long long sm[kPkThreads / 32];
  best = warp_min_u64(best);
  if (lane == 0) sm[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long k = sm[0];
#pragma unroll
    for (int i = 1; i < kPkThreads / 32; ++i) k = sm[i] < k ? sm[i] : k;
    if (k != ~0ull) atomicMin(key_out, k);
  }
}

// ---- the same sweep fed by warp-private rings of TMA bulk copies -----------------------------------------
// The register-prefetch kernel above keeps ONE row (1 KB per warp) in flight: 20 warps x 1 KB x 148 SMs = 3 MB,
// less than half of what Little's law asks for at 6.5 TB/s x ~1 us.  Here every warp owns a ring of D slots of
// 2 KB (two stored rows of its 32 templates) with one mbarrier each; one elected lane refills a slot with one
// cp.async.bulk as soon as the warp has consumed it, so 2 D KB per warp are always in flight and the loads cost
// no registers and no LSU issue slots beyond two LDS.128 per row.  A group is 16 ring items: rows 1..30 in
// pairs (rows 0 and 31 are never read in reference mode) and the row-sum block; 16 / D fills per slot and
// group is even, so the mbarrier parity of an item is a compile-time constant.
__device__ __forceinline__ uint32_t vt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void vt_mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void vt_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_%=;\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// one ring fill: expect BYTES, then the bulk copy global -> shared that completes them on the barrier.  The
// offset is an immediate of the instruction: handed over as a pointer, every item's address becomes its own
// loop-carried 64-bit induction variable (measured: 168 registers instead of 96).
template <int OFF, int BYTES>
__device__ __forceinline__ void vt_ring_fill(uint32_t dst, const void* base, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "n"(BYTES) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1+%2], %3, [%4];" ::"r"(dst),
               "l"(base), "n"(OFF), "n"(BYTES), "r"(bar)
               : "memory");
}
// one elected lane of a converged warp (the form ptxas keeps on the uniform datapath)
__device__ __forceinline__ bool vt_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// A ring item is TWO stored rows (2 KB): rows 2j+1 and 2j+2 for j < 15, then the row-sum block (j = 15).
constexpr int kItemBytes = 2048;
constexpr int kItemsPerGroup = 16;
__host__ __device__ constexpr int pk_item_off(int j) { return (j < 15 ? 2 * j + 1 : 32) * 1024; }

template <int D, int S>
struct RingPrologue {
  __device__ __forceinline__ static void run(uint32_t ring, uint32_t bars, const char* gb) {
    if constexpr (S < D) {
      vt_ring_fill<pk_item_off(S), kItemBytes>(ring + S * kItemBytes, gb, bars + S * 8);
      RingPrologue<D, S + 1>::run(ring, bars, gb);
    }
  }
};

template <int D, int J, bool LOCK>
struct RingItems {
  // ring / bars: shared addresses of this warp's slots and barriers; gb / nb: this group and the warp's next one.
  // LOCK: the warps of the CTA meet at a barrier after every item, so that all of them execute the same stretch of
  // the (48 KB, straight-line) code at the same time; `active` is false for a warp that has run out of groups.
  __device__ __forceinline__ static void run(uint32_t ring, uint32_t bars, const char* gb, const char* nb, int lane,
                                             bool active, uint32_t (&cnt)[15], uint4 (&rs)[4]) {
    if constexpr (J < kItemsPerGroup) {
      constexpr int slot = J % D;
      if (!LOCK || active) {
        const uint32_t sl = ring + slot * kItemBytes + lane * 16;
        vt_mbar_wait(bars + slot * 8, (J / D) & 1);
        const uint4 lo0 = lds_u4(sl), hi0 = lds_u4(sl + 512), lo1 = lds_u4(sl + 1024), hi1 = lds_u4(sl + 1536);
        if constexpr (J < 15) {
          ref_row<2 * J + 1>(lo0, hi0, cnt);
          ref_row<2 * J + 2>(lo1, hi1, cnt);
        } else {
          rs[0] = lo0, rs[1] = hi0, rs[2] = lo1, rs[3] = hi1;
        }
        // the slot's contents are in registers (the compares above consumed them): refill it with item J + D
        __syncwarp();
        if (vt_elect_one()) {
          if constexpr (J + D < kItemsPerGroup)
            vt_ring_fill<pk_item_off(J + D), kItemBytes>(ring + slot * kItemBytes, gb, bars + slot * 8);
          else if (nb != nullptr)
            vt_ring_fill<pk_item_off((J + D) % kItemsPerGroup), kItemBytes>(ring + slot * kItemBytes, nb, bars + slot * 8);
        }
      }
      if constexpr (LOCK) __syncthreads();
      RingItems<D, J + 1, LOCK>::run(ring, bars, gb, nb, lane, active, cnt, rs);
    }
  }
};

template <int D, int NT, bool LOCK>
__global__ void __launch_bounds__(NT, NT == 128 ? 5 : 1)  // 96 registers: the query planes stay constant-bank operands
    k_vt_sweep_packed_ref_ring(const uint4* __restrict__ packed, long long n, long long base_index,
                               unsigned long long* __restrict__ key_out, uint32_t* __restrict__ scores,
                               const int* __restrict__ n_dev) {
  static_assert(kItemsPerGroup % D == 0 && (kItemsPerGroup / D) % 2 == 0, "the parity of an item must not depend on the group");
  extern __shared__ __align__(128) unsigned char vt_ring_smem[];
  if (n_dev != nullptr) n = *n_dev;
  const int lane = threadIdx.x & 31;
  const int wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform, and known to be
  const long long n_groups = (n + 31) >> 5;
  const long long warp0 = (long long)blockIdx.x * (NT / 32) + wid;
  const long long n_warps = (long long)gridDim.x * (NT / 32);
  const uint32_t ring = vt_smem_u32(vt_ring_smem) + wid * D * kItemBytes;
  const uint32_t bars = vt_smem_u32(vt_ring_smem) + (NT / 32) * D * kItemBytes + wid * D * 8;
  constexpr long long kGroupBytes = (long long)kGroupU4 * 16;
  if (vt_elect_one()) {
#pragma unroll
    for (int s = 0; s < D; ++s) vt_mbar_init(bars + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp0 < n_groups) {
      const char* gb = reinterpret_cast<const char*>(packed) + warp0 * kGroupBytes;
      RingPrologue<D, 0>::run(ring, bars, gb);
    }
  }
  __syncwarp();
  unsigned long long best = ~0ull;
  // LOCK: every warp of the CTA makes as many rounds as its first warp (the one with the lowest group index)
  for (long long g = warp0; (LOCK ? g - wid : g) < n_groups; g += n_warps) {
    const bool active = g < n_groups;
    const char* gb = reinterpret_cast<const char*>(packed) + g * kGroupBytes;
    const char* nb = (g + n_warps < n_groups) ? gb + n_warps * kGroupBytes : nullptr;
    uint32_t cnt[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) cnt[i] = 0;
    uint4 rs[4] = {};
    RingItems<D, 0, LOCK>::run(ring, bars, gb, nb, lane, active, cnt, rs);
    uint32_t R[32];
#pragma unroll
    for (int w4 = 0; w4 < 4; ++w4) {
      const uint32_t ww[4] = {rs[w4].x, rs[w4].y, rs[w4].z, rs[w4].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        R[(w4 * 4 + j) * 2] = ww[j] & 0xffffu;
        R[(w4 * 4 + j) * 2 + 1] = ww[j] >> 16;
      }
    }
    uint32_t A = 0;
#pragma unroll
    for (int r = 1; r <= 16; ++r) A += R[r];
    const uint32_t bq = c_vtq[256];
    uint32_t m = 0xffffffffu;
#pragma unroll
    for (int o = -7; o <= 7; ++o) {
      const uint32_t sc = A + 256u * cnt[o + 7] - bq;
      m = min(m, sc);
      if (o < 7) A = A - R[8 + o] + R[24 + o];
    }
    const long long ti = g * 32 + lane;
    if (active && ti < n) {
      const unsigned long long key = ((unsigned long long)m << 32) | (unsigned long long)(base_index + ti);
      best = key < best ? key : best;
      if (scores != nullptr) scores[ti] = m;
    }
  }
  __shared__ unsigned long long sm[NT / 32];
  best = warp_min_u64(best);
  if (lane == 0) sm[wid] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long k = sm[0];
#pragma unroll
    for (int i = 1; i < NT / 32; ++i) k = sm[i] < k ? sm[i] : k;
    if (k != ~0ull) atomicMin(key_out, k);
  }
}

// ---- float32 library, reference mode, fed the same way ----------------------------------------------------
// k_vt_sweep_f32 loads a template's 30 rows into registers and only then starts to compare: a warp has nothing
// in flight while it computes.  Here a warp owns a ring of D templates (rows 1..30 = 3840 contiguous bytes each,
// one bulk copy), so D templates per warp are always on their way while one is being compared out of shared
// memory.  Same arithmetic in the same order as k_vt_sweep_f32: the scores are bit-identical.
constexpr int kF32ItemBytes = 30 * 32 * 4;
constexpr int kF32RingThreads = 128;

template <int D>
__global__ void __launch_bounds__(kF32RingThreads)
    k_vt_sweep_f32_ring(const float* __restrict__ lib, long long n, const float* __restrict__ query, long long base_index,
                        unsigned long long* __restrict__ key_out, float* __restrict__ scores) {
  using S = VtShape<PRS_VT_MODE_REF>;
  extern __shared__ __align__(128) unsigned char vt_ring_smem[];
  const int lane = threadIdx.x & 31;
  const int wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const long long warp0 = (long long)blockIdx.x * (kF32RingThreads / 32) + wid;
  const long long n_warps = (long long)gridDim.x * (kF32RingThreads / 32);
  const uint32_t ring = vt_smem_u32(vt_ring_smem) + wid * D * kF32ItemBytes;
  const uint32_t bars = vt_smem_u32(vt_ring_smem) + (kF32RingThreads / 32) * D * kF32ItemBytes + wid * D * 8;
  const char* base = reinterpret_cast<const char*>(lib) + 128;  // row 1 of template 0
  if (vt_elect_one()) {
#pragma unroll
    for (int s = 0; s < D; ++s) vt_mbar_init(bars + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
    for (int s = 0; s < D; ++s)
      if (warp0 + s * n_warps < n)
        vt_ring_fill<0, kF32ItemBytes>(ring + s * kF32ItemBytes, base + (warp0 + s * n_warps) * 4096, bars + s * 8);
  }
  __syncwarp();
  float q[S::NS];
#pragma unroll
  for (int s = 0; s < S::NS; ++s) q[s] = query[(S::S0 + s) * 32 + lane];
  unsigned long long best = ~0ull;
  int slot = 0;
  uint32_t parity = 0;
  for (long long ti = warp0; ti < n; ti += n_warps) {
    const uint32_t sl = ring + slot * kF32ItemBytes;
    vt_mbar_wait(bars + slot * 8, parity);
    float a[S::T1 - S::T0];
#pragma unroll
    for (int t = 0; t < S::T1 - S::T0; ++t)
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a[t]) : "r"(sl + t * 128 + lane * 4));
    __syncwarp();
    if (vt_elect_one()) {
      const long long nt = ti + D * n_warps;
      if (nt < n) vt_ring_fill<0, kF32ItemBytes>(sl, base + nt * 4096, bars + slot * 8);
    }
    if (++slot == D) {
      slot = 0;
      parity ^= 1u;
    }
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
  __shared__ unsigned long long sm[kF32RingThreads / 32];
  best = warp_min_u64(best);
  if (lane == 0) sm[wid] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long k = sm[0];
#pragma unroll
    for (int i = 1; i < kF32RingThreads / 32; ++i) k = sm[i] < k ? sm[i] : k;
    if (k != ~0ull) atomicMin(key_out, k);
  }
}

// ---- float32, two columns per lane -------------------------------------------------------------------------
// The ring kernel above is bound by instruction issue (ncu: 85 % of the issue slots, 2 instructions per
// element-difference).  Here a lane owns two adjacent columns, so that the subtraction is one packed add.f32x2 for
// two differences (3 instructions per two) and a template needs only half a warp: a warp works on templates
// 2i and 2i+1 at once, a ring slot holds both (two bulk copies on one barrier).  The slot is refilled after it has
// been compared (the other D - 1 slots are in flight meanwhile).  The sums are formed in another order than in
// k_vt_sweep_f32: scores agree to rounding (exactly, for integer-valued profiles), not bit for bit.
template <int D>
__global__ void __launch_bounds__(kF32RingThreads)
    k_vt_sweep_f32_pair(const float* __restrict__ lib, long long n, const float* __restrict__ query, long long base_index,
                        unsigned long long* __restrict__ key_out, float* __restrict__ scores) {
  using S = VtShape<PRS_VT_MODE_REF>;
  extern __shared__ __align__(128) unsigned char vt_ring_smem[];
  const int lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15;
  const int wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const long long n_pairs = (n + 1) >> 1;
  const long long warp0 = (long long)blockIdx.x * (kF32RingThreads / 32) + wid;
  const long long n_warps = (long long)gridDim.x * (kF32RingThreads / 32);
  constexpr int kSlot = 2 * kF32ItemBytes;
  const uint32_t ring = vt_smem_u32(vt_ring_smem) + wid * D * kSlot;
  const uint32_t bars = vt_smem_u32(vt_ring_smem) + (kF32RingThreads / 32) * D * kSlot + wid * D * 8;
  const char* base = reinterpret_cast<const char*>(lib) + 128;  // row 1 of template 0
  auto fill = [&](uint32_t slot_addr, long long pi, uint32_t bar) {  // one elected lane
    const char* src = base + pi * 8192;
    if (2 * pi + 1 < n) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "n"(kSlot) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       slot_addr),
                   "l"(src), "n"(kF32ItemBytes), "r"(bar)
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1+4096], %2, [%3];" ::"r"(
                       slot_addr + kF32ItemBytes),
                   "l"(src), "n"(kF32ItemBytes), "r"(bar)
                   : "memory");
    } else {  // odd library size: the last pair has one template
      vt_ring_fill<0, kF32ItemBytes>(slot_addr, src, bar);
    }
  };
  if (vt_elect_one()) {
#pragma unroll
    for (int s = 0; s < D; ++s) vt_mbar_init(bars + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
    for (int s = 0; s < D; ++s)
      if (warp0 + s * n_warps < n_pairs) fill(ring + s * kSlot, warp0 + s * n_warps, bars + s * 8);
  }
  __syncwarp();
  float2 nq[S::NS];  // minus the query's two columns of this lane
#pragma unroll
  for (int s = 0; s < S::NS; ++s) {
    const float2 v = *reinterpret_cast<const float2*>(query + (S::S0 + s) * 32 + 2 * l16);
    nq[s] = make_float2(-v.x, -v.y);
  }
  unsigned long long best = ~0ull;
  int slot = 0;
  uint32_t parity = 0;
  for (long long pi = warp0; pi < n_pairs; pi += n_warps) {
    const uint32_t sl = ring + slot * kSlot;
    vt_mbar_wait(bars + slot * 8, parity);
    const uint32_t mine = sl + half * kF32ItemBytes + l16 * 8;
    float acc[S::NACC];
#pragma unroll
    for (int i = 0; i < S::NACC; ++i) acc[i] = 0.f;
#pragma unroll
    for (int t = S::T0; t < S::T1; ++t) {
      float2 a;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a.x), "=f"(a.y) : "r"(mine + (t - S::T0) * 128));
#pragma unroll
      for (int s = 0; s < S::NS; ++s) {
        if (S::valid(t, S::S0 + s)) {
          const float2 d = __fadd2_rn(a, nq[s]);
          acc[S::off(t, S::S0 + s)] = (acc[S::off(t, S::S0 + s)] + fabsf(d.x)) + fabsf(d.y);
        }
      }
    }
    // every lane's loads have been consumed by the sums above: refill the slot with the pair D rounds ahead
    __syncwarp();
    if (vt_elect_one()) {
      const long long np = pi + D * n_warps;
      if (np < n_pairs) fill(sl, np, bars + slot * 8);
    }
    if (++slot == D) {
      slot = 0;
      parity ^= 1u;
    }
    int obase = 0;
    Butterfly<S::NACC, 8, float>::run(acc, lane, obase);  // over the 16 lanes of the half-warp
    float m = (obase < S::NOFF) ? acc[0] : INFINITY;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const long long ti = 2 * pi + half;
    if (ti < n) {
      unsigned long long key = ((unsigned long long)__float_as_uint(m) << 32) | (unsigned long long)(base_index + ti);
      best = key < best ? key : best;
      if (scores != nullptr && l16 == 0) scores[ti] = m;
    }
  }
  __shared__ unsigned long long sm[kF32RingThreads / 32];
  best = warp_min_u64(best);
  if (lane == 0) sm[wid] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long k = sm[0];
#pragma unroll
    for (int i = 1; i < kF32RingThreads / 32; ++i) k = sm[i] < k ? sm[i] : k;
    if (k != ~0ull) atomicMin(key_out, k);
  }
}

// Circular mode (all 32 cyclic row shifts; extension).  The shift of a (stored row t, query row s) pair is
// (t - s) mod 32, which depends on the run-time row t -- but only through t mod 8 once the 32 counters are kept
// in registers and ROTATED by eight places after every eight stored rows: inside a block of eight rows the
// pair (i, s) always adds into register (i - s) mod 32, and the rotation re-labels the registers for the next
// block.  32 register MOVs per block replace 256 shared-memory read-modify-writes.
// eight stored rows (t0 .. t0+7, held in registers) against every query row; query row outermost so that only
// its eight plane words are live as (uniform-register) operands at a time
#ifndef PRS_VT_CIRC_SYNC_MASK
#define PRS_VT_CIRC_SYNC_MASK 3
#endif
// LOCK: the warps of the CTA meet at a barrier after every four query rows (32 row pairs), so that they fetch the
// same stretch of the unrolled code together; `active` is false for a warp that has run out of groups.
template <bool LOCK>
__device__ __forceinline__ void circ_block(const uint4 (&lo)[8], const uint4 (&hi)[8], uint32_t (&cnt)[32], bool active) {
#pragma unroll
  for (int s = 0; s < 32; ++s) {
    if (!LOCK || active) {
#pragma unroll
      for (int i = 0; i < 8; ++i) cnt[(i - s) & 31] += lt_row(lo[i], hi[i], s);
    }
    if (LOCK && (s & PRS_VT_CIRC_SYNC_MASK) == PRS_VT_CIRC_SYNC_MASK) __syncthreads();
  }
}

template <bool LOCK>
__global__ void __launch_bounds__(kPkThreads, 4)
    k_vt_sweep_packed_circ(const uint4* __restrict__ packed, long long n, long long base_index,
                           unsigned long long* __restrict__ key_out, uint32_t* __restrict__ scores,
                           const int* __restrict__ n_dev) {
  if (n_dev != nullptr) n = *n_dev;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long n_groups = (n + 31) >> 5;
  const long long warp0 = ((long long)blockIdx.x * kPkThreads + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * kPkThreads) >> 5;
  unsigned long long best = ~0ull;
  // LOCK: every warp of the CTA makes as many rounds as its first warp; one without a group only keeps the barriers
  for (long long g = warp0; (LOCK ? g - wid : g) < n_groups; g += n_warps) {
    const bool active = g < n_groups;
    const uint4* gp = packed + (active ? g : 0) * kGroupU4 + lane;
    uint32_t cnt[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) cnt[o] = 0;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      // during block c register r counts shift (r + 8c) mod 32
      uint4 lo[8], hi[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        lo[i] = ld_stream_u4(gp + ((8 * c + i) * 2 + 0) * 32);
        hi[i] = ld_stream_u4(gp + ((8 * c + i) * 2 + 1) * 32);
      }
      circ_block<LOCK>(lo, hi, cnt, active);
      uint32_t rot[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) rot[r] = cnt[(r + 8) & 31];
#pragma unroll
      for (int r = 0; r < 32; ++r) cnt[r] = rot[r];
    }
    // four rotations by eight: register r counts shift r again
    uint32_t A = 0;
#pragma unroll
    for (int w4 = 0; w4 < 4; ++w4) {
      const uint4 v = ld_stream_u4(gp + 32 * 2 * 32 + w4 * 32);
      A += (v.x & 0xffffu) + (v.x >> 16) + (v.y & 0xffffu) + (v.y >> 16) + (v.z & 0xffffu) + (v.z >> 16) +
           (v.w & 0xffffu) + (v.w >> 16);
    }
    const uint32_t bq = c_vtq[257];
    uint32_t m = 0xffffffffu;
#pragma unroll
    for (int o = 0; o < 32; ++o) m = min(m, A + 256u * cnt[o] - bq);
    const long long ti = g * 32 + lane;
    if (active && ti < n) {
      const unsigned long long key = ((unsigned long long)m << 32) | (unsigned long long)(base_index + ti);
      best = key < best ? key : best;
      if (scores != nullptr) scores[ti] = m;
    }
  }
  __shared__ unsigned long long sm[kPkThreads / 32];
  best = warp_min_u64(best);
  if (lane == 0) sm[wid] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long k = sm[0];
#pragma unroll
    for (int i = 1; i < kPkThreads / 32; ++i) k = sm[i] < k ? sm[i] : k;
    if (k != ~0ull) atomicMin(key_out, k);
  }
}

static int g_vt_circ_lock = 1;  // knob 4: warps of a CTA in lock step in the circular sweep

extern "C" int prs_vt_tune(int knob, int value) {
  PRS_REQUIRE(knob >= 0 && knob <= 4, "prs_vt_tune: unknown knob %d", knob);
  if (knob == 4) {
    g_vt_circ_lock = value != 0;
    return PRS_OK;
  }
  if (knob == 0) PRS_REQUIRE(value == 0 || value == 2 || value == 4 || value == 8 || value == 34 || value == 44,
                             "prs_vt_tune: ring depth must be 0, 2, 4 or 8, or 34 / 44 for the lock-step variants");
  if (knob == 2)
    PRS_REQUIRE((value >= 0 && value <= 4) || (value >= 11 && value <= 13),
                "prs_vt_tune: float32 ring depth must be in 0..4, or 11..13 for the column-pair kernel");
  if (knob == 1 || knob == 3) PRS_REQUIRE(value >= 1 && value <= 32, "prs_vt_tune: CTAs per SM must be in 1..32");
  g_vt_knob[knob] = value;
  return PRS_OK;
}

template <int D>
static int launch_f32_ring_d(int blocks, const float* lib, long long n, const float* query, long long base_index,
                             unsigned long long* key_out, float* scores, cudaStream_t st) {
  constexpr int smem = (kF32RingThreads / 32) * D * (kF32ItemBytes + 8);
  if (smem > 48 * 1024)
    PRS_CUDA(cudaFuncSetAttribute(k_vt_sweep_f32_ring<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_vt_sweep_f32_ring<D><<<blocks, kF32RingThreads, smem, st>>>(lib, n, query, base_index, key_out, scores);
  PRS_CUDA(cudaGetLastError());
  return PRS_OK;
}

template <int D>
static int launch_f32_pair_d(int blocks, const float* lib, long long n, const float* query, long long base_index,
                             unsigned long long* key_out, float* scores, cudaStream_t st) {
  constexpr int smem = (kF32RingThreads / 32) * D * (2 * kF32ItemBytes + 8);
  if (smem > 48 * 1024)
    PRS_CUDA(cudaFuncSetAttribute(k_vt_sweep_f32_pair<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_vt_sweep_f32_pair<D><<<blocks, kF32RingThreads, smem, st>>>(lib, n, query, base_index, key_out, scores);
  PRS_CUDA(cudaGetLastError());
  return PRS_OK;
}

static int launch_f32_ring(int depth, int ctas_per_sm, const float* lib, long long n, const float* query,
                           long long base_index, unsigned long long* key_out, float* scores, cudaStream_t st) {
  if (depth >= 11) {  // two columns per lane, two templates per warp; ring depth = depth - 10
    const long long pairs = (n + 1) / 2;
    long long blocks = (pairs + (kF32RingThreads / 32) - 1) / (kF32RingThreads / 32);
    if (blocks > 148LL * ctas_per_sm) blocks = 148LL * ctas_per_sm;
    switch (depth - 10) {
      case 1: return launch_f32_pair_d<1>((int)blocks, lib, n, query, base_index, key_out, scores, st);
      case 2: return launch_f32_pair_d<2>((int)blocks, lib, n, query, base_index, key_out, scores, st);
      default: return launch_f32_pair_d<3>((int)blocks, lib, n, query, base_index, key_out, scores, st);
    }
  }
  long long blocks = (n + (kF32RingThreads / 32) - 1) / (kF32RingThreads / 32);
  if (blocks > 148LL * ctas_per_sm) blocks = 148LL * ctas_per_sm;
  switch (depth) {
    case 1: return launch_f32_ring_d<1>((int)blocks, lib, n, query, base_index, key_out, scores, st);
    case 2: return launch_f32_ring_d<2>((int)blocks, lib, n, query, base_index, key_out, scores, st);
    case 3: return launch_f32_ring_d<3>((int)blocks, lib, n, query, base_index, key_out, scores, st);
    default: return launch_f32_ring_d<4>((int)blocks, lib, n, query, base_index, key_out, scores, st);
  }
}

template <int D, int NT, bool LOCK>
static int launch_ref_ring(long long groups, int ctas_per_sm, const uint4* packed, long long n, long long base_index,
                           unsigned long long* key_out, uint32_t* scores, const int* n_dev, cudaStream_t st) {
  constexpr int smem = (NT / 32) * D * (kItemBytes + 8);
  if (smem > 48 * 1024)  // per device; cheap enough to repeat
    PRS_CUDA(cudaFuncSetAttribute(k_vt_sweep_packed_ref_ring<D, NT, LOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long blocks = (groups + (NT / 32) - 1) / (NT / 32);
  if (blocks > 148LL * ctas_per_sm) blocks = 148LL * ctas_per_sm;
  if (blocks < 1) blocks = 1;
  k_vt_sweep_packed_ref_ring<D, NT, LOCK><<<(int)blocks, NT, smem, st>>>(packed, n, base_index, key_out, scores, n_dev);
  return PRS_OK;
}

// n: the library size the kernel uses when n_dev is NULL; n_grid: the size the grid is laid out for
static int launch_packed_sweep(const uint4* packed, long long n, long long n_grid, int mode, long long base_index,
                               unsigned long long* key_out, uint32_t* scores, const int* n_dev, cudaStream_t st) {
  const long long groups = (n_grid + 31) / 32;
  long long blocks = (groups + (kPkThreads / 32) - 1) / (kPkThreads / 32);
  if (blocks < 1) blocks = 1;
  if (mode == PRS_VT_MODE_REF) {
    const int depth = g_vt_knob[0];
    const long long cap = 148LL * (depth ? (int)g_vt_knob[1] : 16);
    if (blocks > cap) blocks = cap;
    int rc = PRS_OK;
    const int cps = g_vt_knob[1];
    switch (depth) {
      case 2: rc = launch_ref_ring<2, 128, false>(groups, cps, packed, n, base_index, key_out, scores, n_dev, st); break;
      case 4: rc = launch_ref_ring<4, 128, false>(groups, cps, packed, n, base_index, key_out, scores, n_dev, st); break;
      case 8: rc = launch_ref_ring<8, 128, false>(groups, cps, packed, n, base_index, key_out, scores, n_dev, st); break;
      // lock-step variants: the warps of a CTA meet at a barrier after every ring item
      case 34: rc = launch_ref_ring<4, 128, true>(groups, cps, packed, n, base_index, key_out, scores, n_dev, st); break;
      case 44: rc = launch_ref_ring<4, 640, true>(groups, 1, packed, n, base_index, key_out, scores, n_dev, st); break;
      default:
        k_vt_sweep_packed_ref<<<(int)blocks, kPkThreads, 0, st>>>(packed, n, base_index, key_out, scores, n_dev);
    }
    if (rc != PRS_OK) return rc;
  } else {
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    if (g_vt_circ_lock)
      k_vt_sweep_packed_circ<true><<<(int)blocks, kPkThreads, 0, st>>>(packed, n, base_index, key_out, scores, n_dev);
    else
      k_vt_sweep_packed_circ<false><<<(int)blocks, kPkThreads, 0, st>>>(packed, n, base_index, key_out, scores, n_dev);
  }
  PRS_CUDA(cudaGetLastError());
  return PRS_OK;
}

// VtqScope for callers outside this file (sharded.cu replays a graph that contains a packed sweep): the scope object
// lives from begin to end; one at a time per host thread is all the mutex allows anyway.
static thread_local VtqScope* t_ext_scope = nullptr;
int prs_vtq_begin(cudaStream_t st) {
  VtqScope* sc = new (std::nothrow) VtqScope();
  PRS_REQUIRE(sc, "out of host memory");
  int rc = sc->begin(st);
  if (rc != PRS_OK) {
    delete sc;
    return rc;
  }
  t_ext_scope = sc;
  return PRS_OK;
}
int prs_vtq_end(cudaStream_t) {
  VtqScope* sc = t_ext_scope;
  t_ext_scope = nullptr;
  if (!sc) return PRS_OK;
  int rc = sc->end();
  delete sc;
  return rc;
}

extern "C" size_t prs_vt_packed_bytes(long long n) {
  return (size_t)((n + 31) / 32) * kGroupU4 * sizeof(uint4);
}

extern "C" int prs_vt_pack_u8(const uint8_t* src, long long n, void* packed, long long first, void* stream) {
  PRS_REQUIRE(packed && n >= 0 && first >= 0 && (src || n == 0), "prs_vt_pack_u8: bad argument");
  if (n == 0) return PRS_OK;
  const long long threads = n * 32;
  k_vt_pack_u8<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, n, (uint4*)packed, first);
  PRS_CUDA(cudaGetLastError());
  return PRS_OK;
}

extern "C" int prs_vt_unpack_u8(const void* packed, long long index, uint8_t* dst, void* stream) {
  PRS_REQUIRE(packed && dst && index >= 0, "prs_vt_unpack_u8: bad argument");
  k_vt_unpack_u8<<<1, 1024, 0, (cudaStream_t)stream>>>((const uint4*)packed, index, dst);
  PRS_CUDA(cudaGetLastError());
  return PRS_OK;
}

extern "C" int prs_vt_sweep_packed_u8(const void* packed, long long n, const uint8_t* query, int mode,
                                      long long base_index, unsigned long long* key_out, uint32_t* scores,
                                      void* scratch, void* stream) {
  PRS_REQUIRE(query && key_out && scratch && n >= 0 && (packed || n == 0), "prs_vt_sweep_packed_u8: bad argument");
  PRS_REQUIRE(mode == PRS_VT_MODE_REF || mode == PRS_VT_MODE_CIRCULAR, "prs_vt_sweep_packed_u8: unknown mode %d", mode);
  PRS_REQUIRE(base_index >= 0 && base_index + n <= 0xffffffffLL,
              "prs_vt_sweep_packed_u8: template index does not fit 32 bits");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    PRS_CUDA(cudaMemsetAsync(key_out, 0xff, sizeof(unsigned long long), st));
    return PRS_OK;
  }
  // query -> bit planes -> constant bank (stream ordered; sweeps on other streams of the device are ordered behind
  // each other through VtqScope)
  k_vt_pack_query<<<1, 32, 0, st>>>(query, (uint32_t*)scratch, key_out);
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (st != nullptr) PRS_CUDA(cudaStreamIsCapturing(st, &cap));
  VtqScope scope;
  if (cap == cudaStreamCaptureStatusNone)  // a caller capturing its own graph orders its launches itself
    if (int rc = scope.begin(st)) return rc;
  PRS_CUDA(cudaMemcpyToSymbolAsync(c_vtq, scratch, (32 * 8 + 2) * sizeof(uint32_t), 0, cudaMemcpyDeviceToDevice, st));
  int rc = launch_packed_sweep((const uint4*)packed, n, n, mode, base_index, key_out, scores, nullptr, st);
  if (int rc2 = scope.end()) return rc != PRS_OK ? rc : rc2;
  return rc;
}

// The two halves of prs_vt_sweep_packed_u8 for a caller that replays the second one as a graph (sharded.cu): the query's
// bit planes (and the reset of the key) on the caller's stream, then constant upload + sweep from those planes.
int prs_vt_pack_query_launch(const uint8_t* query, void* scratch, unsigned long long* key, cudaStream_t st) {
  k_vt_pack_query<<<1, 32, 0, st>>>(query, (uint32_t*)scratch, key);
  PRS_CUDA(cudaGetLastError());
  return PRS_OK;
}
int prs_vt_sweep_packed_planes(const void* packed, long long n, const void* planes, int mode, long long base_index,
                               unsigned long long* key_out, cudaStream_t st) {
  if (n == 0) return PRS_OK;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (st != nullptr) PRS_CUDA(cudaStreamIsCapturing(st, &cap));
  VtqScope scope;
  if (cap == cudaStreamCaptureStatusNone)
    if (int rc = scope.begin(st)) return rc;
  PRS_CUDA(cudaMemcpyToSymbolAsync(c_vtq, planes, (32 * 8 + 2) * sizeof(uint32_t), 0, cudaMemcpyDeviceToDevice, st));
  int rc = launch_packed_sweep((const uint4*)packed, n, n, mode, base_index, key_out, nullptr, nullptr, st);
  if (int rc2 = scope.end()) return rc != PRS_OK ? rc : rc2;
  return rc;
}

// =============================================================================================
// One whole frame of the ROS loop (ratslam/ros_simulate.py:98-105,134-137) with a single host
// synchronisation: odometry H2D -> pose-cell update -> frame H2D -> sub-sample -> library sweep ->
// create-or-match decided ON THE DEVICE (view_templates.py:67-73; a created template is packed straight
// into slot n of the library) -> 32-byte result D2H.
struct FrameScratch {               // device scratch layout (bytes)
  static constexpr size_t kFrame = 0;                   // uint8 frame, up to 1 MiB
  static constexpr size_t kTpl = 1 << 20;               // uint8[1024] sub-sampled template
  static constexpr size_t kPlanes = kTpl + 1024;        // uint32[264] query planes
  static constexpr size_t kKey = kPlanes + 2048;        // uint64 key
  static constexpr size_t kResult = kKey + 64;          // prs_frame_result
  static constexpr size_t kOdom = kResult + 64;         // double[2]
  static constexpr size_t kBytes = kOdom + 64;
};

__global__ void k_vt_decide_append(const unsigned long long* __restrict__ key, const uint8_t* __restrict__ tpl,
                                   uint4* __restrict__ packed, int n, unsigned threshold,
                                   const long long* __restrict__ argmax, const int* __restrict__ pc_err,
                                   prs_frame_result* __restrict__ res, int* __restrict__ n_dev, int write_pc = 1) {
  const int t = threadIdx.x;  // 32 threads, one per template row
  if (n_dev != nullptr) n = *n_dev;
  __syncwarp();
  const unsigned long long k = *key;
  const unsigned score = (unsigned)(k >> 32);
  const bool create = (n == 0) || (k == ~0ull) || (score > threshold);  // strict '>' (view_templates.py:67)
  if (create) {
    vt_pack_row(tpl, packed, n, t);
  }
  if (t == 0) {
    if (write_pc) res->argmax = argmax ? argmax[0] : -1;  // else the pose-cell kernel writes both fields itself
    res->key = k;
    res->created = create ? 1 : 0;
    res->template_index = create ? n : (int)(k & 0xffffffffu);
    res->n_templates = n + (create ? 1 : 0);
    if (write_pc) res->pc_err = pc_err ? pc_err[0] : 0;
    if (n_dev != nullptr) *n_dev = n + (create ? 1 : 0);
  }
}

extern "C" size_t prs_frame_scratch_bytes(void) { return FrameScratch::kBytes; }

extern "C" int prs_frame_host(prs_pc_handle pc, void* pc_state, const void* gi, void* pc_work, const double* odom_host,
                              void* vt_packed, int n_templates, unsigned threshold, int mode,
                              const uint8_t* frame_host, int im_rows, int im_cols, int row_lo, int row_hi, int row_step,
                              int col_lo, int col_hi, int col_step, void* scratch, prs_frame_result* result_host,
                              void* stream) {
  PRS_REQUIRE(pc && pc_state && gi && pc_work && vt_packed && frame_host && scratch && result_host,
              "prs_frame_host: null argument");
  PRS_REQUIRE((size_t)im_rows * im_cols <= (1u << 20), "prs_frame_host: frame larger than 1 MiB");
  cudaStream_t st = (cudaStream_t)stream;
  char* sc = (char*)scratch;
  uint8_t* d_frame = (uint8_t*)(sc + FrameScratch::kFrame);
  uint8_t* d_tpl = (uint8_t*)(sc + FrameScratch::kTpl);
  uint32_t* d_planes = (uint32_t*)(sc + FrameScratch::kPlanes);
  unsigned long long* d_key = (unsigned long long*)(sc + FrameScratch::kKey);
  prs_frame_result* d_res = (prs_frame_result*)(sc + FrameScratch::kResult);
  double* d_odom = (double*)(sc + FrameScratch::kOdom);
  // pc_work: device int64 argmax[1], float/double total[1] (8 bytes), int err[1]
  long long* d_argmax = (long long*)pc_work;
  void* d_total = (char*)pc_work + 8;
  int* d_err = (int*)((char*)pc_work + 16);
  int rc;
  if (odom_host) {  // ros_simulate.py:134-137: one pose-cell update with this twist
    PRS_CUDA(cudaMemcpyAsync(d_odom, odom_host, 2 * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = prs_pc_step(pc, pc_state, d_odom, gi, d_argmax, d_total, d_err, st);
    if (rc != PRS_OK) return rc;
  }
  PRS_CUDA(cudaMemcpyAsync(d_frame, frame_host, (size_t)im_rows * im_cols, cudaMemcpyHostToDevice, st));
  rc = prs_vt_extract_u8(d_frame, im_rows, im_cols, row_lo, row_hi, row_step, col_lo, col_hi, col_step, d_tpl, 32, 32, st);
  if (rc != PRS_OK) return rc;
  rc = prs_vt_sweep_packed_u8(vt_packed, n_templates, d_tpl, mode, 0, d_key, nullptr, d_planes, st);
  if (rc != PRS_OK) return rc;
  k_vt_decide_append<<<1, 32, 0, st>>>(d_key, d_tpl, (uint4*)vt_packed, n_templates, threshold, d_argmax, d_err, d_res,
                                       nullptr);
  PRS_CUDA(cudaGetLastError());
  PRS_CUDA(cudaMemcpyAsync(result_host, d_res, sizeof(prs_frame_result), cudaMemcpyDeviceToHost, st));
  PRS_CUDA(cudaStreamSynchronize(st));
  return PRS_OK;
}

// =============================================================================================
// Any template shape (rows x cols) and any max_offset: one warp per template, lanes stride over the
// window's pixels.  Used when ViewTemplates is configured away from the reference's 32x32 / offset 8
// (view_templates.py:14,44); correctness path, not tuned.
template <typename T, typename ACC>
__global__ void __launch_bounds__(256)
    k_vt_sweep_any(const T* __restrict__ lib, long long n, const T* __restrict__ query, int rows, int cols, int max_offset,
                   long long base_index, unsigned long long* __restrict__ key_out, ACC* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * 256) >> 5;
  const int win = rows - 2 * max_offset;  // rows compared per offset
  unsigned long long best = ~0ull;
  for (long long ti = warp0; ti < n; ti += n_warps) {
    const T* tp = lib + ti * rows * cols;
    ACC m = 0;
    bool first = true;
    for (int o = -max_offset + 1; o < max_offset; ++o) {
      ACC s = 0;
      for (int i = lane; i < win * cols; i += 32) {
        const int r = i / cols, c = i - r * cols;
        const T a = tp[(max_offset + o + r) * cols + c], b = query[(max_offset + r) * cols + c];
        if constexpr (sizeof(T) == 1)
          s += (ACC)(uint8_t)(a - b);  // uint8 arithmetic wraps; abs() is the identity (view_templates.py:23)
        else
          s += fabsf(a - b);
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
      if (first || s < m) m = s;
      first = false;
    }
    unsigned hi;
    if constexpr (sizeof(T) == 1)
      hi = (unsigned)m;
    else
      hi = __float_as_uint(m);
    const unsigned long long key = ((unsigned long long)hi << 32) | (unsigned long long)(base_index + ti);
    best = key < best ? key : best;
    if (scores != nullptr && lane == 0) scores[ti] = m;
  }
  block_min_to_global(best, key_out);
}

template <typename T, typename ACC>
static int sweep_any(const T* lib, long long n, const T* query, int rows, int cols, int max_offset, long long base_index,
                     unsigned long long* key_out, ACC* scores, cudaStream_t st) {
  PRS_REQUIRE(query && key_out && n >= 0 && (lib || n == 0), "prs_vt_sweep_any: bad argument");
  PRS_REQUIRE(rows > 0 && cols > 0 && max_offset >= 1 && rows - 2 * max_offset >= 0,
              "prs_vt_sweep_any: %dx%d templates cannot be compared with max_offset %d", rows, cols, max_offset);
  PRS_REQUIRE(base_index >= 0 && base_index + n <= 0xffffffffLL, "prs_vt_sweep_any: template index does not fit 32 bits");
  PRS_CUDA(cudaMemsetAsync(key_out, 0xff, sizeof(unsigned long long), st));
  if (n == 0) return PRS_OK;
  k_vt_sweep_any<T, ACC><<<sweep_grid(n), 256, 0, st>>>(lib, n, query, rows, cols, max_offset, base_index, key_out, scores);
  PRS_CUDA(cudaGetLastError());
  return PRS_OK;
}

extern "C" int prs_vt_sweep_any_u8(const uint8_t* lib, long long n, const uint8_t* query, int rows, int cols,
                                   int max_offset, long long base_index, unsigned long long* key_out, uint32_t* scores,
                                   void* stream) {
  return sweep_any<uint8_t, uint32_t>(lib, n, query, rows, cols, max_offset, base_index, key_out, scores, (cudaStream_t)stream);
}

extern "C" int prs_vt_sweep_any_f32(const float* lib, long long n, const float* query, int rows, int cols, int max_offset,
                                    long long base_index, unsigned long long* key_out, float* scores, void* stream) {
  return sweep_any<float, float>(lib, n, query, rows, cols, max_offset, base_index, key_out, scores, (cudaStream_t)stream);
}


// ---------------------------------------------------------------------------------------------
// Short frame chain for replayed frames over small libraries.  With the plan's host buffers pinned (and therefore
// mapped into the device's address space) the template branch of a frame is three kernels and no copy node:
//   k_vt_frame_prepare   reads the 32x32 masked pixels straight from the pinned frame (about 2 KB cross the bus
//                        instead of the whole 64 KB frame), builds the template, its bit planes and row sums,
//                        resets the key
//   k_vt_sweep_packed_ref_small   the bit-sliced sweep with the query planes in shared memory instead of the
//                        constant bank (no cudaMemcpyToSymbol node) and one block per group of 32 templates;
//                        only used below kSmallLibrary templates
//   k_vt_decide_append   writes its 32-byte result straight into the pinned result buffer
// and the pose-cell kernel reads its two doubles of odometry from the pinned buffer as well.
constexpr int kSmallLibrary = 1 << 16;

__global__ void __launch_bounds__(1024)
    k_vt_frame_prepare(const uint8_t* __restrict__ frame, int im_cols, int row_lo, int row_step, int col_lo, int col_step,
                       uint8_t* __restrict__ tpl, uint32_t* __restrict__ planes, unsigned long long* __restrict__ key) {
  __shared__ uint32_t s_sum[32];
  const int t = threadIdx.x >> 5, c = threadIdx.x & 31;  // template row, column
  const int rr = row_lo + (t / (row_step - 1)) * row_step + (t % (row_step - 1)) + 1;
  const int cc = col_lo + (c / (col_step - 1)) * col_step + (c % (col_step - 1)) + 1;
  const uint32_t px = frame[(size_t)rr * im_cols + cc];
  tpl[t * 32 + c] = (uint8_t)px;
  uint32_t mine = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t b = __ballot_sync(0xffffffffu, (px >> k) & 1u);
    if (c == k) mine = b;
  }
  if (c < 8) planes[t * 8 + c] = mine;
  uint32_t sum = px;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (c == 0) s_sum[t] = sum;
  __syncthreads();
  if (t == 0) {
    uint32_t all = s_sum[c], mid = (c >= 8 && c < 24) ? all : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      all += __shfl_xor_sync(0xffffffffu, all, o);
      mid += __shfl_xor_sync(0xffffffffu, mid, o);
    }
    if (c == 0) {
      planes[256] = mid;
      planes[257] = all;
      *key = ~0ull;
    }
  }
}

__device__ __forceinline__ uint32_t lt_row_q(const uint4& lo, const uint4& hi, const uint32_t* q) {
  uint32_t lt = ~lo.x & q[0];
  lt = (~lo.y & q[1]) | (~(lo.y ^ q[1]) & lt);
  lt = (~lo.z & q[2]) | (~(lo.z ^ q[2]) & lt);
  lt = (~lo.w & q[3]) | (~(lo.w ^ q[3]) & lt);
  lt = (~hi.x & q[4]) | (~(hi.x ^ q[4]) & lt);
  lt = (~hi.y & q[5]) | (~(hi.y ^ q[5]) & lt);
  lt = (~hi.z & q[6]) | (~(hi.z ^ q[6]) & lt);
  lt = (~hi.w & q[7]) | (~(hi.w ^ q[7]) & lt);
  return (uint32_t)__popc(lt);
}

// Small libraries are latency-bound, not bandwidth-bound: with one warp per group of 32 templates the 30 stored
// rows are 30 dependent load rounds (measured: 25 us for a 30-template library).  Here a block of 16 warps shares
// one group: warp w < 15 compares rows 2w+1 and 2w+2 (one load round, 480 compares), warp 15 forms the window sums,
// and the per-offset counts meet in shared memory (integer adds: exact, order-free).
constexpr int kSmallThreads = 512;

__global__ void __launch_bounds__(kSmallThreads)
    k_vt_sweep_packed_ref_small(const uint4* __restrict__ packed, const uint32_t* __restrict__ qplanes,
                                unsigned long long* __restrict__ key_out, const int* __restrict__ n_dev) {
  __shared__ uint32_t q[32 * 8 + 8];
  __shared__ uint32_t s_cnt[15][32];
  __shared__ uint32_t s_A[15][32];
  const long long n = *n_dev;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long n_groups = (n + 31) >> 5;
  if ((long long)blockIdx.x >= n_groups) return;  // the grid is laid out for the capacity, not the live count
  for (int i = threadIdx.x; i < 32 * 8 + 2; i += kSmallThreads) q[i] = qplanes[i];
  unsigned long long best = ~0ull;
  for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
    if (threadIdx.x < 15 * 32) s_cnt[wid][lane] = 0;
    __syncthreads();  // also publishes q on the first pass
    const uint4* gp = packed + g * kGroupU4 + lane;
    if (wid < 15) {
      const int t0 = 2 * wid + 1;
      const uint4 lo0 = ld_stream_u4(gp + (t0 * 2 + 0) * 32), hi0 = ld_stream_u4(gp + (t0 * 2 + 1) * 32);
      const uint4 lo1 = ld_stream_u4(gp + (t0 * 2 + 2) * 32), hi1 = ld_stream_u4(gp + (t0 * 2 + 3) * 32);
#pragma unroll
      for (int o = -7; o <= 7; ++o) {
        uint32_t c = 0;
        const int s0 = t0 - o, s1 = t0 + 1 - o;  // warp-uniform
        if (s0 >= 8 && s0 <= 23) c += lt_row_q(lo0, hi0, q + s0 * 8);
        if (s1 >= 8 && s1 <= 23) c += lt_row_q(lo1, hi1, q + s1 * 8);
        if (c) atomicAdd(&s_cnt[o + 7][lane], c);
      }
    } else {
      uint32_t R[32];
#pragma unroll
      for (int w4 = 0; w4 < 4; ++w4) {
        const uint4 v = ld_stream_u4(gp + 32 * 2 * 32 + w4 * 32);
        const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          R[(w4 * 4 + j) * 2] = ww[j] & 0xffffu;
          R[(w4 * 4 + j) * 2 + 1] = ww[j] >> 16;
        }
      }
      uint32_t A = 0;
#pragma unroll
      for (int r = 1; r <= 16; ++r) A += R[r];
#pragma unroll
      for (int o = -7; o <= 7; ++o) {
        s_A[o + 7][lane] = A;
        if (o < 7) A = A - R[8 + o] + R[24 + o];
      }
    }
    __syncthreads();
    if (wid == 0) {
      const uint32_t bq = q[256];
      uint32_t m = 0xffffffffu;
#pragma unroll
      for (int o = 0; o < 15; ++o) m = min(m, s_A[o][lane] + 256u * s_cnt[o][lane] - bq);
      const long long ti = g * 32 + lane;
      if (ti < n) {
        const unsigned long long key = ((unsigned long long)m << 32) | (unsigned long long)ti;
        best = key < best ? key : best;
      }
    }
    __syncthreads();
  }
  if (wid == 0) {
    best = warp_min_u64(best);
    if (lane == 0 && best != ~0ull) atomicMin(key_out, best);
  }
}

// ---------------------------------------------------------------------------------------------
// The same frame as a CUDA graph: every pointer is fixed at creation (pinned host buffers included), the
// library size lives in device memory, so a frame is ONE graph launch + one synchronisation instead of
// eleven stream operations.  Two executables: with and without the pose-cell update (ros_simulate.py:128
// drops near-zero twists).
struct prs_frame_plan {
  prs_pc_handle pc;
  void *pc_state, *pc_work, *vt_packed, *scratch;
  const void* gi;
  double* odom_host;
  uint8_t* frame_host;
  prs_frame_result* result_host;
  unsigned threshold;
  int mode, capacity;
  int im_rows, im_cols, row_lo, row_hi, row_step, col_lo, col_hi, col_step;
  cudaGraphExec_t exec[2];
  bool ready[2];
  int* d_n;
  int warm[2];
  cudaStream_t side;       // the pose-cell update runs here, concurrently with the template branch
  cudaEvent_t ev_fork, ev_join;
  cudaEvent_t ev_done;     // recorded behind every frame launched by prs_replay_run
  // device-side aliases of the pinned host buffers (zero-copy), null when a buffer is not pinned
  const double* odom_map;
  const uint8_t* frame_map;
  prs_frame_result* result_map;
};

static int frame_enqueue(prs_frame_plan* f, bool moved, cudaStream_t st) {
  char* sc = (char*)f->scratch;
  uint8_t* d_frame = (uint8_t*)(sc + FrameScratch::kFrame);
  uint8_t* d_tpl = (uint8_t*)(sc + FrameScratch::kTpl);
  uint32_t* d_planes = (uint32_t*)(sc + FrameScratch::kPlanes);
  unsigned long long* d_key = (unsigned long long*)(sc + FrameScratch::kKey);
  prs_frame_result* d_res = (prs_frame_result*)(sc + FrameScratch::kResult);
  double* d_odom = (double*)(sc + FrameScratch::kOdom);
  long long* d_argmax = (long long*)f->pc_work;
  void* d_total = (char*)f->pc_work + 8;
  int* d_err = (int*)((char*)f->pc_work + 16);
  const bool zero_copy = f->odom_map && f->frame_map && f->result_map;
  // the short chain (see k_vt_frame_prepare): reference mode, 32x32 templates out of a step >= 2 mask, small library
  const bool short_chain = zero_copy && f->mode == PRS_VT_MODE_REF && f->capacity <= kSmallLibrary;
  int rc;
  int pc_mirrored = 0;
  if (moved) {  // fork: the pose-cell update does not depend on the frame until the decision
    PRS_CUDA(cudaEventRecord(f->ev_fork, st));
    PRS_CUDA(cudaStreamWaitEvent(f->side, f->ev_fork, 0));
    const double* od = f->odom_map;
    if (!zero_copy) {
      PRS_CUDA(cudaMemcpyAsync(d_odom, f->odom_host, 2 * sizeof(double), cudaMemcpyHostToDevice, f->side));
      od = d_odom;
    }
    // with mapped buffers the cluster kernel writes the arg-max and the error bits into the host record itself:
    // nothing on the template branch waits for the pose-cell update then
    if (zero_copy)
      rc = prs_pc_step_mirror(f->pc, f->pc_state, od, f->gi, d_argmax, d_total, d_err, &f->result_map->argmax,
                              &f->result_map->pc_err, &pc_mirrored, f->side);
    else
      rc = prs_pc_step(f->pc, f->pc_state, od, f->gi, d_argmax, d_total, d_err, f->side);
    if (rc != PRS_OK) return rc;
    PRS_CUDA(cudaEventRecord(f->ev_join, f->side));
  }
  if (short_chain) {
    rc = check_mask(f->im_rows, f->im_cols, f->row_lo, f->row_hi, f->row_step, f->col_lo, f->col_hi, f->col_step, 32, 32);
    if (rc != PRS_OK) return rc;
    k_vt_frame_prepare<<<1, 1024, 0, st>>>(f->frame_map, f->im_cols, f->row_lo, f->row_step, f->col_lo, f->col_step,
                                           d_tpl, d_planes, d_key);
    long long blocks = ((long long)f->capacity + 31) / 32;
    if (blocks > 148LL * 4) blocks = 148LL * 4;
    k_vt_sweep_packed_ref_small<<<(int)blocks, kSmallThreads, 0, st>>>((const uint4*)f->vt_packed, d_planes, d_key, f->d_n);
  } else {
    PRS_CUDA(cudaMemcpyAsync(d_frame, f->frame_host, (size_t)f->im_rows * f->im_cols, cudaMemcpyHostToDevice, st));
    rc = prs_vt_extract_u8(d_frame, f->im_rows, f->im_cols, f->row_lo, f->row_hi, f->row_step, f->col_lo, f->col_hi,
                           f->col_step, d_tpl, 32, 32, st);
    if (rc != PRS_OK) return rc;
    PRS_CUDA(cudaMemsetAsync(d_key, 0xff, sizeof(unsigned long long), st));
    k_vt_pack_query<<<1, 32, 0, st>>>(d_tpl, d_planes);
    PRS_CUDA(cudaMemcpyToSymbolAsync(c_vtq, d_planes, (32 * 8 + 2) * sizeof(uint32_t), 0, cudaMemcpyDeviceToDevice, st));
    // grid sized for the capacity; the kernels read the live count from device memory
    rc = launch_packed_sweep((const uint4*)f->vt_packed, 0, f->capacity, f->mode, 0, d_key, nullptr, f->d_n, st);
    if (rc != PRS_OK) return rc;
  }
  if (moved && !pc_mirrored) PRS_CUDA(cudaStreamWaitEvent(st, f->ev_join, 0));  // the decision reports the new arg-max
  k_vt_decide_append<<<1, 32, 0, st>>>(d_key, d_tpl, (uint4*)f->vt_packed, 0, f->threshold, d_argmax, d_err,
                                       zero_copy ? f->result_map : d_res, f->d_n, pc_mirrored ? 0 : 1);
  PRS_CUDA(cudaGetLastError());
  if (moved && pc_mirrored) PRS_CUDA(cudaStreamWaitEvent(st, f->ev_join, 0));  // join: the frame ends with both branches
  if (!zero_copy)
    PRS_CUDA(cudaMemcpyAsync(f->result_host, d_res, sizeof(prs_frame_result), cudaMemcpyDeviceToHost, st));
  return PRS_OK;
}

extern "C" int prs_frame_create(prs_pc_handle pc, void* pc_state, const void* gi, void* pc_work, void* vt_packed,
                                int n_templates, int capacity, unsigned threshold, int mode, int im_rows, int im_cols,
                                int row_lo, int row_hi, int row_step, int col_lo, int col_hi, int col_step,
                                void* scratch, double* odom_host, uint8_t* frame_host, prs_frame_result* result_host,
                                prs_frame_plan** out) {
  PRS_REQUIRE(pc && pc_state && gi && pc_work && vt_packed && scratch && odom_host && frame_host && result_host && out,
              "prs_frame_create: null argument");
  PRS_REQUIRE((size_t)im_rows * im_cols <= (1u << 20), "prs_frame_create: frame larger than 1 MiB");
  PRS_REQUIRE(capacity >= 1 && n_templates >= 0 && n_templates < capacity, "prs_frame_create: library has no free slot");
  prs_frame_plan* f = new (std::nothrow) prs_frame_plan();
  PRS_REQUIRE(f, "prs_frame_create: out of host memory");
  *f = prs_frame_plan{pc, pc_state, pc_work, vt_packed, scratch, gi, odom_host, frame_host, result_host, threshold, mode,
                      capacity, im_rows, im_cols, row_lo, row_hi, row_step, col_lo, col_hi, col_step, {nullptr, nullptr},
                      {false, false}, nullptr, {0, 0}, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  {  // pinned buffers are mapped: kernels can read the inputs and write the result in place
    auto mapped = [](const void* p) -> void* {
      cudaPointerAttributes a;
      if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
      }
      return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
    };
    f->odom_map = (const double*)mapped(odom_host);
    f->frame_map = (const uint8_t*)mapped(frame_host);
    f->result_map = (prs_frame_result*)mapped(result_host);
  }
  f->d_n = (int*)((char*)scratch + FrameScratch::kOdom + 32);
  cudaError_t e = cudaStreamCreateWithFlags(&f->side, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->ev_join, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->ev_done, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMemcpy(f->d_n, &n_templates, sizeof(int), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    prs_set_error("prs_frame_create: %s", cudaGetErrorString(e));
    delete f;
    return PRS_E_CUDA;
  }
  *out = f;
  return PRS_OK;
}

extern "C" int prs_frame_destroy(prs_frame_plan* f) {
  if (f) {
    for (int i = 0; i < 2; ++i)
      if (f->ready[i]) cudaGraphExecDestroy(f->exec[i]);
    if (f->side) cudaStreamDestroy(f->side);
    if (f->ev_fork) cudaEventDestroy(f->ev_fork);
    if (f->ev_join) cudaEventDestroy(f->ev_join);
    if (f->ev_done) cudaEventDestroy(f->ev_done);
    delete f;
  }
  return PRS_OK;
}

// The library size the plan's kernels read lives in device memory (shared by the plans built on one scratch
// buffer): a caller that appended templates outside the plan (ViewTemplates.match / create) puts the host's count back.
extern "C" int prs_frame_set_count(prs_frame_plan* f, int n_templates, void* stream) {
  PRS_REQUIRE(f && n_templates >= 0 && n_templates < f->capacity, "prs_frame_set_count: bad argument");
  PRS_CUDA(cudaMemcpyAsync(f->d_n, &n_templates, sizeof(int), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  PRS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return PRS_OK;
}

// One frame.  `moved` != 0: odom_host (given at creation) holds (vtrans, vrot) and the pose cells are updated first.
// The first call of each kind runs eagerly (it also warms every lazily initialised kernel attribute), the second
// captures the graph, later ones replay it.
extern "C" int prs_frame_launch(prs_frame_plan* f, int moved, void* stream) {
  PRS_REQUIRE(f, "prs_frame_launch: null plan");
  cudaStream_t st = (cudaStream_t)stream;
  const int v = moved ? 1 : 0;
  int rc = PRS_OK;
  // the long chain sweeps through the per-device constant buffer: order the whole frame behind sweeps on other streams
  const bool uses_vtq = !(f->odom_map && f->frame_map && f->result_map && f->mode == PRS_VT_MODE_REF && f->capacity <= kSmallLibrary);
  VtqScope scope;
  if (uses_vtq)
    if (int rc0 = scope.begin(st)) return rc0;
  if (f->ready[v]) {
    PRS_CUDA(cudaGraphLaunch(f->exec[v], st));
  } else if (f->warm[v] < 1 || st == nullptr) {  // the legacy default stream cannot be captured
    ++f->warm[v];
    rc = frame_enqueue(f, moved != 0, st);
    if (rc != PRS_OK) return rc;
  } else {
    cudaGraph_t g = nullptr;
    PRS_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    rc = frame_enqueue(f, moved != 0, st);
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (rc != PRS_OK || e != cudaSuccess) {
      if (g) cudaGraphDestroy(g);
      if (rc == PRS_OK) prs_set_error("prs_frame_launch: capture failed: %s", cudaGetErrorString(e));
      return rc != PRS_OK ? rc : PRS_E_CUDA;
    }
    PRS_CUDA(cudaGraphInstantiate(&f->exec[v], g, 0));
    cudaGraphDestroy(g);
    f->ready[v] = true;
    PRS_CUDA(cudaGraphLaunch(f->exec[v], st));
  }
  return scope.end();
}

extern "C" int prs_frame_run(prs_frame_plan* f, int moved, void* stream) {
  int rc = prs_frame_launch(f, moved, stream);
  if (rc != PRS_OK) return rc;
  PRS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return PRS_OK;
}

// The whole replay loop (ros_simulate.py:152-166 driven from arrays) on the host side of the C ABI: frame t is
// staged into the pinned buffers of plan t mod n_plans and launched while the earlier frames are still running;
// a plan is reused once its event has fired and its 32-byte result has been copied out.  All plans must share the
// device-side state (pose cells, library, template count, scratch) and differ only in their pinned host buffers;
// the library must have room for T more templates.
//   frames  : host uint8 [T][im_rows][im_cols] (pageable is fine: each frame is copied into the plan's pinned buffer)
//   odom    : host double [T][2] = (vtrans, vrot) as passed to PoseCellNetwork.update
//   moved   : host uint8 [T], 0 = no pose-cell update for this frame (ros_simulate.py:128)
//   results : host prs_frame_result [T]
extern "C" int prs_replay_run(prs_frame_plan* const* plans, int n_plans, const uint8_t* frames, const double* odom,
                              const uint8_t* moved, int T, prs_frame_result* results, void* stream) {
  PRS_REQUIRE(plans && n_plans >= 1 && n_plans <= 16 && T >= 0 && results, "prs_replay_run: bad argument");
  PRS_REQUIRE(T == 0 || (frames && odom && moved), "prs_replay_run: null input");
  PRS_REQUIRE(stream != nullptr, "prs_replay_run: needs a non-default stream (the frames replay as CUDA graphs)");
  for (int i = 0; i < n_plans; ++i) PRS_REQUIRE(plans[i], "prs_replay_run: null plan");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t fbytes = (size_t)plans[0]->im_rows * plans[0]->im_cols;
  int rc = PRS_OK;
  int launched = 0;
  for (int t = 0; t < T; ++t) {
    prs_frame_plan* f = plans[t % n_plans];
    if (t >= n_plans) {
      PRS_CUDA(cudaEventSynchronize(f->ev_done));
      results[t - n_plans] = *f->result_host;
    }
    memcpy(f->frame_host, frames + (size_t)t * fbytes, fbytes);
    if (moved[t]) {
      f->odom_host[0] = odom[2 * t];
      f->odom_host[1] = odom[2 * t + 1];
    }
    rc = prs_frame_launch(f, moved[t] ? 1 : 0, st);
    if (rc != PRS_OK) break;
    PRS_CUDA(cudaEventRecord(f->ev_done, st));
    launched = t + 1;
  }
  // drain: the last min(n_plans, launched) frames
  for (int t = launched > n_plans ? launched - n_plans : 0; t < launched; ++t) {
    prs_frame_plan* f = plans[t % n_plans];
    if (cudaEventSynchronize(f->ev_done) != cudaSuccess) {
      prs_set_error("prs_replay_run: %s", cudaGetErrorString(cudaGetLastError()));
      return PRS_E_CUDA;
    }
    results[t] = *f->result_host;
  }
  return rc;
}
