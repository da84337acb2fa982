// Stand-alone probe of a 3-D tiled tensor map with a 40 x 22 x 1 float box (the large-grid pose-cell kernel's fetch).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu   (no -lcuda: the encoder comes from the runtime)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void probe(const __grid_constant__ CUtensorMap tmap, const CUtensorMap* tmap_g, int use_global, int c0, int c1,
                      int c2, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* dst = reinterpret_cast<float*>(smem);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + 4096);
  const unsigned bar_a = (unsigned)__cvta_generic_to_shared(bar), dst_a = (unsigned)__cvta_generic_to_shared(dst);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(40 * 22 * 4) : "memory");
    const void* d = use_global ? (const void*)tmap_g : (const void*)&tmap;
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst_a),
        "l"(d), "r"(c0), "r"(c1), "r"(c2), "r"(bar_a)
        : "memory");
  }
  __syncthreads();
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@!p bra WAIT_%=;\n}\n" ::"r"(bar_a)
      : "memory");
  for (int i = threadIdx.x; i < 40 * 22; i += blockDim.x) out[i] = dst[i];
}

int main() {
  const int YP = 272, XP = 272, NP = 8;
  float* d;
  cudaMalloc(&d, (size_t)NP * XP * YP * 4);
  float* h = (float*)malloc((size_t)NP * XP * YP * 4);
  for (size_t i = 0; i < (size_t)NP * XP * YP; ++i) h[i] = (float)(i % 1000003);
  cudaMemcpy(d, h, (size_t)NP * XP * YP * 4, cudaMemcpyHostToDevice);
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  printf("entry point: %s q=%d fn=%p\n", cudaGetErrorString(e), (int)q, fn);
  alignas(64) CUtensorMap tm;
  const cuuint64_t dims[3] = {YP, XP, NP};
  const cuuint64_t strides[2] = {(cuuint64_t)YP * 4, (cuuint64_t)YP * XP * 4};
  const cuuint32_t box[3] = {40, 22, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  for (int l2 = 0; l2 < 2; ++l2) {
    CUresult r = ((encode_fn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, l2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode (l2 promotion %d): %d\n", l2, (int)r);
    CUtensorMap* tg;
    cudaMalloc(&tg, 128);
    cudaMemcpy(tg, &tm, 128, cudaMemcpyHostToDevice);
    float* out;
    cudaMalloc(&out, 40 * 22 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
    for (int ug = 0; ug < 2; ++ug) {
      const int cs[3][3] = {{0, 0, 0}, {5, 7, 3}, {-3, 260, 7}};
      for (int t = 0; t < 3; ++t) {
        probe<<<1, 128, 8192>>>(tm, tg, ug, cs[t][0], cs[t][1], cs[t][2], out);
        e = cudaDeviceSynchronize();
        float ho[40 * 22];
        cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r2 = 0; r2 < 22; ++r2)
          for (int c = 0; c < 40; ++c) {
            const int gy = cs[t][0] + c, gx = cs[t][1] + r2;
            const float want = (gy < 0 || gy >= YP || gx < 0 || gx >= XP) ? 0.f : h[((size_t)cs[t][2] * XP + gx) * YP + gy];
            if (ho[r2 * 40 + c] != want) ++bad;
          }
        printf("  desc in %s, coords (%d,%d,%d): %s, %d mismatches\n", ug ? "global" : "param", cs[t][0], cs[t][1], cs[t][2],
               cudaGetErrorString(e), bad);
        if (e != cudaSuccess) return 1;
      }
    }
  }
  return 0;
}
