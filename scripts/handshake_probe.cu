// Stand-alone probe: what does one pipeline stage of the conv kernels' barrier skeleton cost when
// no data moves at all?  (The v3 gather-GEMM with copies, TMA, MMAs and stores switched off still
// takes 2/3 of its full time - profiles/r1_conv_ablation.txt.)  Same role layout as conv_tc3.cu:
// 4 producer warps, one MMA thread, one TMA thread, ring of S stages, full/empty mbarriers.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I openpcseg_b200/csrc \
//        scripts/handshake_probe.cu -o gpurun_out/handshake_probe && gpurun_out/handshake_probe
//
// PMODE (how producers signal "stage full"):
//   0  cp.async.mbarrier.arrive.noinc from all 128 threads           (what conv_tc3.cu does)
//   1  __syncwarp + one plain arrive per warp
//   2  fence.proxy.async + __syncwarp + one plain arrive per warp
//   3  plain mbarrier.arrive from all 128 threads
//   4  cp.async.commit_group + wait_group 0 + __syncwarp + one arrive per warp
//   5  like 0 but only one producer warp exists (32 arrivals)
// EMODE (how the MMA thread frees a stage): 0 tcgen05.commit, 1 plain mbarrier.arrive
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace b2s::tc;

__device__ __forceinline__ void arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

template <int PMODE, int EMODE>
__global__ void __launch_bounds__(192) probe_kernel(int n_stage, int S, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // only there to pin CTAs/SM
  __shared__ __align__(8) uint64_t s_full[8];
  __shared__ __align__(8) uint64_t s_empty[8];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int prod_warps = PMODE == 5 ? 1 : 4;
  const int per_thread = (PMODE == 0 || PMODE == 3 || PMODE == 5);
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&s_full[s]), (per_thread ? prod_warps * 32 : prod_warps) + 1);
      mbar_init(smem_u32(&s_empty[s]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(smem_u32(&s_tmem), 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const long long t0 = clock64();
  if (warp < prod_warps) {
    int s = 0, wraps = 0;
    for (int i = 0; i < n_stage; ++i) {
      if (wraps > 0) mbar_wait(smem_u32(&s_empty[s]), (wraps - 1) & 1);
      const uint32_t bar = smem_u32(&s_full[s]);
      if (PMODE == 0 || PMODE == 5) {
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
      } else if (PMODE == 3) {
        mbar_arrive(bar);
      } else {
        if (PMODE == 2) fence_proxy_async();
        if (PMODE == 4) {
          cp_async_commit();
          cp_async_wait<0>();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar);
      }
      if (++s == S) { s = 0; ++wraps; }
    }
  } else if (warp == 4) {
    if (lane == 0) {
      int s = 0, wraps = 0;
      for (int i = 0; i < n_stage; ++i) {
        mbar_wait(smem_u32(&s_full[s]), wraps & 1);
        tc_fence_after();
        if (EMODE == 0) umma_commit(smem_u32(&s_empty[s]));
        else mbar_arrive(smem_u32(&s_empty[s]));
        if (++s == S) { s = 0; ++wraps; }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      int s = 0, wraps = 0;
      for (int i = 0; i < n_stage; ++i) {
        if (wraps > 0) mbar_wait(smem_u32(&s_empty[s]), (wraps - 1) & 1);
        arrive_expect_tx(smem_u32(&s_full[s]), 0u);
        if (++s == S) { s = 0; ++wraps; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(s_tmem, 64);
  }
}

template <int PMODE, int EMODE>
void run(int ctas_per_sm, int S, int n_stage, long long* d_cycles, int sms) {
  const size_t smem = ctas_per_sm == 2 ? 100 * 1024 : 200 * 1024;
  cudaFuncSetAttribute(probe_kernel<PMODE, EMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = sms * ctas_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe_kernel<PMODE, EMODE><<<grid, 192, smem>>>(n_stage, S, d_cycles);   // warm-up
  cudaEventRecord(e0);
  probe_kernel<PMODE, EMODE><<<grid, 192, smem>>>(n_stage, S, d_cycles);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[1024];
  cudaMemcpy(h, d_cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < grid; ++i) avg += (double)h[i];
  avg /= grid;
  printf("pmode %d emode %d ctas/sm %d S %d: %8.1f us total, %7.1f ns/stage, %7.1f cycles/stage (%s)\n", PMODE,
         EMODE, ctas_per_sm, S, ms * 1e3, ms * 1e6 / n_stage, avg / n_stage, cudaGetErrorString(err));
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  long long* d_cycles;
  cudaMalloc(&d_cycles, sizeof(long long) * 1024);
  const int n = 2000;
  for (int cps = 1; cps <= 2; ++cps)
    for (int S = 2; S <= 6; S += (S == 2 ? 1 : 3)) {
      run<0, 0>(cps, S, n, d_cycles, sms);
      run<0, 1>(cps, S, n, d_cycles, sms);
      run<1, 0>(cps, S, n, d_cycles, sms);
      run<1, 1>(cps, S, n, d_cycles, sms);
      run<2, 0>(cps, S, n, d_cycles, sms);
      run<3, 0>(cps, S, n, d_cycles, sms);
      run<4, 0>(cps, S, n, d_cycles, sms);
      run<5, 0>(cps, S, n, d_cycles, sms);
    }
  return 0;
}
