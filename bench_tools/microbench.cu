// Pipe-rate microbenchmarks that decide the kernel designs (run on the B200 via gpurun):
//   FFMA (3-register / constant-bank coefficient), FFMA2 (fma.rn.f32x2), FADD, DFMA,
//   the view-template inner loop mix (IADD + LOP3 + IDP4A), and shared-memory load widths.
// Prints ops per clock per SM, derived from in-kernel clock64() deltas, and wall-clock Tops/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CHECK(x)                                                                  \
  do {                                                                            \
    cudaError_t e = (x);                                                          \
    if (e != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      return 1;                                                                   \
    }                                                                             \
  } while (0)

constexpr int ITERS = 4096;
__constant__ float c_coef[16];

template <int MODE>
__global__ void __launch_bounds__(256) k_fp(float* out, long long* cycles, float seed) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
  float x = seed * 1.0001f, y = seed * 0.9999f;
  long long t0 = clock64();
  if (MODE == 0) {  // FFMA, three register operands
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
    }
  } else if (MODE == 1) {  // FFMA, coefficient from the constant bank
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], c_coef[i], y);
    }
  } else if (MODE == 2) {  // FFMA2: packed pair of fp32 FMAs
    unsigned long long p[8], xx, yy;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x), "f"(x));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(yy) : "f"(y), "f"(y));
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(xx), "l"(yy));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(p[i]));
  } else if (MODE == 3) {  // FADD
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = a[i] + x;
    }
  } else if (MODE == 4) {  // the float SAD pair: d = t - q ; acc += |d|
    for (int it = 0; it < ITERS / 2; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] += fabsf(x - a[(i + 1) & 15] * 0.5f);
    }
  } else if (MODE == 5) {  // FFMA where two operands are shared by all (register reuse cache friendly)
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(x, y, a[i]);
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// The stage-4 inner loop of the resident pose-cell kernel without any memory traffic:
// acc[j] += row[(j+q-3) mod 21] * cf[q], 21 outputs x 7 taps, as packed FFMA2 (PACKED) or scalar FFMA.
template <bool PACKED, int ORDER>
__global__ void __launch_bounds__(448, 1) k_pattern(float* out, long long* cycles, float seed) {
  constexpr int Y = 21;
  float2 acc[Y], row[Y], cf[7];
#pragma unroll
  for (int j = 0; j < Y; ++j) {
    acc[j] = make_float2(0.f, 0.f);
    row[j] = make_float2(seed + j + threadIdx.x, seed - j);
  }
#pragma unroll
  for (int q = 0; q < 7; ++q) cf[q] = make_float2(1.0f + 1e-6f * q * seed, 1.0f - 1e-6f * q * seed);
  long long t0 = clock64();
  for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
    for (int u = 0; u < (ORDER == 0 ? Y : (ORDER == 1 ? 7 : Y)); ++u) {
#pragma unroll
      for (int v = 0; v < (ORDER == 0 ? 7 : (ORDER == 1 ? Y : 7)); ++v) {
        // ORDER 0: outputs outer, taps inner.  1: taps outer, outputs inner (cf reused).  2: inputs outer (row reused).
        const int j = ORDER == 0 ? u : (ORDER == 1 ? v : (u - v + 3 + Y) % Y);
        const int q = ORDER == 0 ? v : (ORDER == 1 ? u : v);
        if (PACKED) {
          acc[j] = __ffma2_rn(row[(j + q + Y - 3) % Y], cf[q], acc[j]);
        } else {
          acc[j].x = fmaf(row[(j + q + Y - 3) % Y].x, cf[q].x, acc[j].x);
          acc[j].y = fmaf(row[(j + q + Y - 3) % Y].y, cf[q].y, acc[j].y);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < Y; ++j) row[j].x += 1e-7f * acc[(j + 1) % Y].y;  // keep the rows live and changing
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < Y; ++j) s += acc[j].x + acc[j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(256) k_dfma(double* out, long long* cycles, double seed) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x;
  double x = seed * 1.0001, y = seed * 0.9999;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// view-template inner loop: z = (ahi - qlo) ^ afix ^ qhi ; acc = dp4a(z, 0x01010101, acc)
template <int MODE>
__global__ void __launch_bounds__(256) k_int(unsigned* out, long long* cycles, unsigned seed) {
  unsigned acc[8], q1[8], q2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] = 0;
    q1[i] = (seed * (i + 3) + threadIdx.x) & 0x7f7f7f7fu;
    q2[i] = (seed * (i + 7)) & 0x80808080u;
  }
  unsigned a = seed ^ threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    unsigned ahi = a | 0x80808080u, afix = ~a & 0x80808080u;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      unsigned z = (ahi - q1[i]) ^ afix ^ q2[i];
      if (MODE == 0)
        acc[i] = __dp4a(z, 0x01010101u, acc[i]);
      else
        acc[i] += z;  // without the dot product: the ALU-only ceiling
    }
    a = a * 1664525u + 1013904223u;
  }
  long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int WIDTH>  // 1: LDS.32, 2: LDS.64, 4: LDS.128
__global__ void __launch_bounds__(256) k_lds(float* out, long long* cycles) {
  __shared__ __align__(16) float sm[256 * 4 * 4];
  for (int i = threadIdx.x; i < 256 * 16; i += 256) sm[i] = i;
  __syncthreads();
  float s = 0;
  long long t0 = clock64();
  for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int base = ((threadIdx.x + j * 256 + it) & 1023) * WIDTH;
      if (WIDTH == 1) {
        s += sm[base];
      } else if (WIDTH == 2) {
        float2 v = *reinterpret_cast<float2*>(&sm[base]);
        s += v.x + v.y;
      } else {
        float4 v = *reinterpret_cast<float4*>(&sm[base]);
        s += v.x + v.y + v.z + v.w;
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <typename F>
int run(const char* name, double ops_per_thread, int blocks_per_sm, F launch) {
  const int nsm = 148;
  const int grid = nsm * blocks_per_sm;
  long long* cyc;
  CHECK(cudaMalloc(&cyc, grid * sizeof(long long)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch(grid, cyc);  // warm-up
  CHECK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  launch(grid, cyc);
  cudaEventRecord(e1);
  CHECK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  static long long h[148 * 16];
  CHECK(cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double mean = 0;
  for (int i = 0; i < grid; ++i) mean += (double)h[i];
  mean /= grid;
  double ops_sm = ops_per_thread * 256.0 * blocks_per_sm;  // per SM (blocks co-resident)
  printf("%-34s blocks/SM=%d  %8.1f lane-ops/clk/SM   %8.2f Tops/s   (%.3f ms, %.0f clk)\n", name, blocks_per_sm,
         ops_sm / mean, ops_per_thread * 256.0 * grid / (ms * 1e-3) / 1e12, ms, mean);
  cudaFree(cyc);
  return 0;
}

int main() {
  float* outf;
  double* outd;
  unsigned* outu;
  CHECK(cudaMalloc(&outf, 148 * 16 * 256 * sizeof(float)));
  CHECK(cudaMalloc(&outd, 148 * 16 * 256 * sizeof(double)));
  CHECK(cudaMalloc(&outu, 148 * 16 * 256 * sizeof(unsigned)));
  float hc[16];
  for (int i = 0; i < 16; ++i) hc[i] = 1.0f + 1e-6f * i;
  CHECK(cudaMemcpyToSymbol(c_coef, hc, sizeof(hc)));
  cudaDeviceProp prop;
  CHECK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s, %d SMs, %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
  {
    // one 448-thread CTA per SM, like the resident kernel: FMA/clk/SM from the in-kernel cycle count
    for (int var = 0; var < 6; ++var) {
      const int packed = var & 1, order = var >> 1;
      long long* cyc;
      CHECK(cudaMalloc(&cyc, 148 * sizeof(long long)));
      for (int rep = 0; rep < 2; ++rep) {
        switch (var) {
          case 0: k_pattern<false, 0><<<148, 448>>>(outf, cyc, 1.0f); break;
          case 1: k_pattern<true, 0><<<148, 448>>>(outf, cyc, 1.0f); break;
          case 2: k_pattern<false, 1><<<148, 448>>>(outf, cyc, 1.0f); break;
          case 3: k_pattern<true, 1><<<148, 448>>>(outf, cyc, 1.0f); break;
          case 4: k_pattern<false, 2><<<148, 448>>>(outf, cyc, 1.0f); break;
          default: k_pattern<true, 2><<<148, 448>>>(outf, cyc, 1.0f); break;
        }
        CHECK(cudaDeviceSynchronize());
      }
      long long h[148];
      CHECK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
      double mean = 0;
      for (int i = 0; i < 148; ++i) mean += (double)h[i];
      mean /= 148;
      double fma = 2.0 * 147.0 * (ITERS / 8) * 448.0;
      printf("stage-4 pattern %-5s order %d, 448 thr/SM: %7.1f FMA/clk/SM (peak 128)\n", packed ? "FFMA2" : "FFMA", order,
             fma / mean);
      cudaFree(cyc);
    }
  }
  for (int bps : {2, 4, 8}) {
    run("FFMA reg,reg,reg", 16.0 * ITERS, bps, [&](int g, long long* c) { k_fp<0><<<g, 256>>>(outf, c, 1.0f); });
    run("FFMA reg,const,reg", 16.0 * ITERS, bps, [&](int g, long long* c) { k_fp<1><<<g, 256>>>(outf, c, 1.0f); });
    run("FFMA2 (2 fma per lane-op; x2)", 16.0 * ITERS, bps, [&](int g, long long* c) { k_fp<2><<<g, 256>>>(outf, c, 1.0f); });
    run("FADD", 16.0 * ITERS, bps, [&](int g, long long* c) { k_fp<3><<<g, 256>>>(outf, c, 1.0f); });
    run("SAD pair (sub + abs-add) as 2 ops", 2 * 16.0 * ITERS / 2, bps, [&](int g, long long* c) { k_fp<4><<<g, 256>>>(outf, c, 1.0f); });
    run("FFMA shared a,b (reuse)", 16.0 * ITERS, bps, [&](int g, long long* c) { k_fp<5><<<g, 256>>>(outf, c, 1.0f); });
    run("DFMA", 8.0 * ITERS, bps, [&](int g, long long* c) { k_dfma<<<g, 256>>>(outd, c, 1.0); });
    run("VT mix IADD+LOP3+IDP4A (3 ops)", 3 * 8.0 * ITERS, bps, [&](int g, long long* c) { k_int<0><<<g, 256>>>(outu, c, 12345u); });
    run("VT mix IADD+LOP3+IADD (3 ops)", 3 * 8.0 * ITERS, bps, [&](int g, long long* c) { k_int<1><<<g, 256>>>(outu, c, 12345u); });
    run("LDS.32  (words)", 1.0 * ITERS, bps, [&](int g, long long* c) { k_lds<1><<<g, 256>>>(outf, c); });
    run("LDS.64  (words)", 2.0 * ITERS, bps, [&](int g, long long* c) { k_lds<2><<<g, 256>>>(outf, c); });
    run("LDS.128 (words)", 4.0 * ITERS, bps, [&](int g, long long* c) { k_lds<4><<<g, 256>>>(outf, c); });
  }
  return 0;
}
