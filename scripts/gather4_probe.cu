// Stand-alone probe of TMA tile::gather4 as the row-gather engine of the sparse conv kernels:
//   (1) semantics: 4 row indices per instruction, rows land 128 B (or 64 B) apart in the swizzled
//       layout the UMMA descriptors expect; indices < 0 or >= n_rows must zero-fill and still count
//       their bytes on the mbarrier;
//   (2) throughput: stages of 128 gathered rows x (64 + 32) channels per CTA, 2 CTAs/SM.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I openpcseg_b200/csrc \
//        scripts/gather4_probe.cu -o scripts/_bin/gather4_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace b2s::tc;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

__device__ __forceinline__ void arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tm, int col, int r0, int r1, int r2,
                                            int r3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
      : "memory");
}

// one stage, results copied out raw (still swizzled) for the host check
__global__ void __launch_bounds__(32) check_kernel(const __grid_constant__ CUtensorMap tm64,
                                                   const __grid_constant__ CUtensorMap tm32, const int* idx,
                                                   uint4* out_wide, uint4* out_tail) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  __shared__ __align__(8) uint64_t bar;
  const int lane = threadIdx.x;
  if (lane == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (lane == 0) arrive_expect_tx(smem_u32(&bar), 128 * 128 + 128 * 64);
  __syncwarp();
  const int4 r = reinterpret_cast<const int4*>(idx)[lane];
  tma_gather4(base + lane * 512, &tm64, 0, r.x, r.y, r.z, r.w, smem_u32(&bar));
  tma_gather4(base + 16384 + lane * 256, &tm32, 64, r.x, r.y, r.z, r.w, smem_u32(&bar));
  mbar_wait(smem_u32(&bar), 0);
  for (int i = lane; i < 1024; i += 32) out_wide[i] = reinterpret_cast<const uint4*>(gen)[i];
  for (int i = lane; i < 512; i += 32) out_tail[i] = reinterpret_cast<const uint4*>(gen + 16384)[i];
}

// throughput: every CTA runs n_stage stages (ring of S), one loader warp, one consumer thread
__global__ void __launch_bounds__(288) rate_kernel(const __grid_constant__ CUtensorMap tm64,
                                                  const __grid_constant__ CUtensorMap tm32, const int* idx,
                                                  int n_idx_tiles, int n_stage, int S, int with_tail, int W) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t s_full[8];
  __shared__ __align__(8) uint64_t s_empty[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&s_empty[s]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t stage_bytes = 128 * 128 + (with_tail ? 128 * 64 : 0);
  if (warp < W) {
    // 32 gather4 operations per chunk and stage, spread over W warps (32 / W lanes active in each)
    const int per = 32 / W;
    const int op = warp * per + lane;                 // which 4-row group this lane fetches
    int s = 0, wraps = 0;
    for (int i = 0; i < n_stage; ++i) {
      const int t = (blockIdx.x * 131 + i * 7) % n_idx_tiles;
      int4 r = make_int4(-1, -1, -1, -1);
      if (lane < per) r = reinterpret_cast<const int4*>(idx + (size_t)t * 128)[op];
      if (wraps > 0) mbar_wait(smem_u32(&s_empty[s]), (wraps - 1) & 1);
      const uint32_t bar = smem_u32(&s_full[s]);
      if (warp == 0 && lane == 0) arrive_expect_tx(bar, stage_bytes);
      __syncwarp();
      const uint32_t a = base + s * 24576;
      if (lane < per) {
        tma_gather4(a + op * 512, &tm64, 0, r.x, r.y, r.z, r.w, bar);
        if (with_tail) tma_gather4(a + 16384 + op * 256, &tm32, 64, r.x, r.y, r.z, r.w, bar);
      }
      if (++s == S) { s = 0; ++wraps; }
    }
  } else if (warp == 8 && lane == 0) {
    int s = 0, wraps = 0;
    for (int i = 0; i < n_stage; ++i) {
      mbar_wait(smem_u32(&s_full[s]), wraps & 1);
      mbar_arrive(smem_u32(&s_empty[s]));
      if (++s == S) { s = 0; ++wraps; }
    }
  }
}

int main() {
  const int n_rows = 400000, c = 96;
  std::vector<__half> hx((size_t)n_rows * c);
  for (int r = 0; r < n_rows; ++r)
    for (int j = 0; j < c; ++j) hx[(size_t)r * c + j] = __float2half((float)((r * 7 + j * 3) % 2048));
  __half* dx;
  cudaMalloc(&dx, hx.size() * 2);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(ptr);
  CUtensorMap tm64, tm32;
  cuuint64_t dims[2] = {(cuuint64_t)c, (cuuint64_t)n_rows};
  cuuint64_t strides[1] = {(cuuint64_t)c * 2};
  cuuint32_t estr[2] = {1, 1};
  cuuint32_t box64[2] = {64, 1}, box32[2] = {32, 1};
  CUresult r1 = enc(&tm64, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dx, dims, strides, box64, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&tm32, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dx, dims, strides, box32, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d %d\n", (int)r1, (int)r2);

  // ---- semantics
  std::vector<int> hidx(128);
  srand(1);
  for (int i = 0; i < 128; ++i) {
    int v = rand() % n_rows;
    if (i % 5 == 1) v = -1;
    if (i % 17 == 3) v = n_rows + 5;
    hidx[i] = v;
  }
  int* didx;
  cudaMalloc(&didx, 128 * 4);
  cudaMemcpy(didx, hidx.data(), 128 * 4, cudaMemcpyHostToDevice);
  uint4 *dw, *dt;
  cudaMalloc(&dw, 16384);
  cudaMalloc(&dt, 8192);
  cudaMemset(dw, 0xEE, 16384);
  cudaMemset(dt, 0xEE, 8192);
  cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  check_kernel<<<1, 32, 32768>>>(tm64, tm32, didx, dw, dt);
  cudaError_t e = cudaDeviceSynchronize();
  printf("check kernel: %s\n", cudaGetErrorString(e));
  std::vector<__half> hw(8192), ht(4096);
  cudaMemcpy(hw.data(), dw, 16384, cudaMemcpyDeviceToHost);
  cudaMemcpy(ht.data(), dt, 8192, cudaMemcpyDeviceToHost);
  int bad_w = 0, bad_t = 0;
  for (int row = 0; row < 128; ++row) {
    const int src = hidx[row];
    const bool ok = src >= 0 && src < n_rows;
    for (int ch = 0; ch < 64; ++ch) {
      const int chunk = ch / 8, off = row * 128 + ((chunk ^ (row & 7)) << 4) + (ch % 8) * 2;
      const float got = __half2float(hw[off / 2]);
      const float want = ok ? __half2float(hx[(size_t)src * c + ch]) : 0.f;
      if (got != want) ++bad_w;
    }
    for (int ch = 0; ch < 32; ++ch) {
      const int chunk = ch / 8, off = row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4) + (ch % 8) * 2;
      const float got = __half2float(ht[off / 2]);
      const float want = ok ? __half2float(hx[(size_t)src * c + 64 + ch]) : 0.f;
      if (got != want) ++bad_t;
    }
  }
  printf("semantics: wide mismatches %d / 8192, tail mismatches %d / 4096\n", bad_w, bad_t);

  // ---- throughput
  const int n_tiles = 2048;
  std::vector<int> big((size_t)n_tiles * 128);
  for (size_t i = 0; i < big.size(); ++i) big[i] = (rand() % 3 == 0) ? -1 : rand() % n_rows;
  int* dbig;
  cudaMalloc(&dbig, big.size() * 4);
  cudaMemcpy(dbig, big.data(), big.size() * 4, cudaMemcpyHostToDevice);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  for (int cps = 1; cps <= 2; ++cps)
    for (int W = 1; W <= 8; W *= 2)
      for (int tail = 0; tail <= 1; ++tail) {
        const int S = 3;
        cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(cps == 2 ? 110 * 1024 : 200 * 1024));
        const size_t launch_smem = cps == 2 ? 110 * 1024 : 200 * 1024;
        const int grid = prop.multiProcessorCount * cps, n_stage = 400;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        rate_kernel<<<grid, 288, launch_smem>>>(tm64, tm32, dbig, n_tiles, n_stage, S, tail, W);
        cudaEventRecord(e0);
        rate_kernel<<<grid, 288, launch_smem>>>(tm64, tm32, dbig, n_tiles, n_stage, S, tail, W);
        cudaEventRecord(e1);
        e = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double rows = (double)grid * n_stage * 128 * (2.0 / 3.0);
        const double bytes = rows * (tail ? 192 : 128);
        printf("ctas/sm %d loader warps %d tail %d: %7.1f us, %6.1f ns/stage/CTA, %7.1f GB/s gathered (%s)\n", cps,
               W, tail, ms * 1e3, ms * 1e6 / n_stage, bytes / ms / 1e6, cudaGetErrorString(e));
      }
  return 0;
}
