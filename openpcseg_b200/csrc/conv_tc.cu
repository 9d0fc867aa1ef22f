// Sparse convolution, tensor-core kernel family (fp16 in, fp32 accumulate) for sm_100a: dispatch of the
// forward / input-gradient gather-GEMM (conv_tc4.cu step-table kernel; conv_tc2.cu for maps without a step
// table and for inputs of 4 GiB and more) and the weight-gradient kernel (below).
//
// Output-stationary implicit GEMM: one CTA tile owns 128 (or 256) output rows and accumulates ALL kernel
// offsets into one TMEM accumulator; one write per output row, fp32 accumulation across offsets, no atomics.
// Algorithmic FLOPs = 2 * M * C_in * C_out (M = map pairs); algorithmic bytes = 2*C_red*M (gathered rows) +
// 2*C_res*N_rows + 4*K*N_rows + 2*K*C_in*C_out.
#include <stdlib.h>

#include <cuda.h>
#include <string.h>

#include "tc_common.cuh"

namespace b2s {

namespace tc {
// W [K][c_in][c_out] -> W^T [K][c_out][c_in] (the K-major B operand of the forward pass)
__global__ void __launch_bounds__(256) transpose_weight_kernel(const __half* __restrict__ w,
                                                                __half* __restrict__ wt, int c_in,
                                                                int c_out) {
  __shared__ __half tile[32][34];
  const int k = blockIdx.z;
  const __half* src = w + (int64_t)k * c_in * c_out;
  __half* dst = wt + (int64_t)k * c_in * c_out;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;     // x: c_out, y: c_in
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    int ci = y0 + j, co = x0 + tx;
    tile[j][tx] = (ci < c_in && co < c_out) ? src[(int64_t)ci * c_out + co] : __float2half(0.f);
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    int co = x0 + j, ci = y0 + tx;
    if (co < c_out && ci < c_in) dst[(int64_t)co * c_in + ci] = tile[tx][j];
  }
}

// One launch for the fp16 operand copies of MANY fp32 master weights (b2s_weights_refresh): block = one 32 x 32
// (c_in x c_out) tile of one offset of one weight; the fp32 tile is read once (coalesced over c_out), written as the
// fp16 parameter layout [K][c_in][c_out] (the input gradient's operand) and, through shared memory, as the K-major
// layout [K][c_out][c_in] (the forward operand).
struct WeightDesc {
  const float* src;
  __half* cast;
  __half* kmajor;
  int k, c_in, c_out, unit_start;
};
static_assert(sizeof(WeightDesc) == 40, "b2s_weight_desc layout");

__global__ void __launch_bounds__(256) weights_refresh_kernel(const WeightDesc* __restrict__ desc, int n) {
  __shared__ __half tile[32][34];
  const int u = blockIdx.x;
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (desc[mid].unit_start <= u) lo = mid; else hi = mid - 1;
  }
  const WeightDesc d = desc[lo];
  const int tx_n = (d.c_out + 31) >> 5, ty_n = (d.c_in + 31) >> 5;
  const int t = u - d.unit_start;
  const int k = t / (tx_n * ty_n), r = t - k * (tx_n * ty_n);
  if (k >= d.k) return;
  const int y0 = (r / tx_n) << 5, x0 = (r % tx_n) << 5;      // y: c_in, x: c_out
  const int64_t base = (int64_t)k * d.c_in * d.c_out;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int ci = y0 + j, co = x0 + tx;
    __half h = __float2half(0.f);
    if (ci < d.c_in && co < d.c_out) {
      h = __float2half_rn(d.src[base + (int64_t)ci * d.c_out + co]);
      if (d.cast) d.cast[base + (int64_t)ci * d.c_out + co] = h;
    }
    tile[j][tx] = h;
  }
  if (!d.kmajor) return;
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int co = x0 + j, ci = y0 + tx;
    if (co < d.c_out && ci < d.c_in) d.kmajor[base + (int64_t)co * d.c_in + ci] = tile[tx][j];
  }
}

static int tmem_cols_for(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

}  // namespace tc

int launch_gather_gemm_tc2(const void* in, const void* wt, int k, int c_red, int c_res, int flip_k,
                           const int32_t* nbr, const uint32_t* tile_mask, const int32_t* row_perm,
                           int64_t n_rows, const void* bias, void* out, cudaStream_t st);

int launch_gather_gemm_tc4(const void* in, int64_t n_src, const void* wt, int k, int c_red, int c_res, int flip_k,
                           const int32_t* step_rows, const int32_t* step_start, const uint32_t* tile_mask,
                           int tile_rows, const int32_t* row_perm, int64_t n_rows, const void* bias, void* out,
                           double* bn_sums, cudaStream_t st);

bool tc_gather_gemm_supported(int c_red, int c_res) {
  if (c_red % 32 != 0 || c_red < 32) return false;
  if (c_res % 16 != 0 || c_res < 16 || c_res > 512) return false;
  if (c_res > 256 && (c_res / 2) % 16 != 0) return false;
  return true;
}

size_t tc_gather_gemm_workspace(int k, int c_in, int c_out) {
  return align_up((size_t)k * c_in * c_out * sizeof(__half), 256);   // W^T for the forward pass
}

// fp16 W [K][c_in][c_out] -> the K-major B operand of the forward pass [K][c_out][c_in]
void launch_weight_to_kmajor(const void* w, int k, int c_in, int c_out, void* out, cudaStream_t st) {
  dim3 g((unsigned)ceil_div(c_out, 32), (unsigned)ceil_div(c_in, 32), (unsigned)k);
  tc::transpose_weight_kernel<<<g, 256, 0, st>>>(reinterpret_cast<const __half*>(w), reinterpret_cast<__half*>(out),
                                                 c_in, c_out);
}

void launch_weights_refresh(const void* desc, int n, int64_t total_units, cudaStream_t st) {
  tc::weights_refresh_kernel<<<(unsigned)total_units, 256, 0, st>>>(reinterpret_cast<const tc::WeightDesc*>(desc), n);
}

// weight_kmajor != 0: `weight` already is the K-major B operand of this pass ([K][c_res][c_red]); otherwise it is
// the parameter layout [K][c_in][c_out], which the input gradient consumes as is and the forward pass transposes
// into the workspace.  steps != nullptr selects the step-table kernel.
int launch_gather_gemm_tc(const void* in, int64_t n_src, const void* weight, int weight_kmajor, int k, int c_in,
                          int c_out, int transpose_w, int flip_k, const int32_t* nbr, const uint32_t* tile_mask,
                          const int32_t* step_rows, const int32_t* step_start, int tile_rows,
                          const int32_t* row_perm, int64_t n_rows, const void* bias, void* out, double* bn_sums,
                          void* ws, size_t ws_bytes, cudaStream_t st) {
  using namespace tc;
  const int c_red = transpose_w ? c_out : c_in, c_res = transpose_w ? c_in : c_out;
  const __half* wt = reinterpret_cast<const __half*>(weight);
  if (!transpose_w && !weight_kmajor) {
    // forward: B_k[n = c_out][c = c_in] = W[k][c][n]  -> needs W^T
    B2S_REQUIRE(ws && ws_bytes >= tc_gather_gemm_workspace(k, c_in, c_out), B2S_ERR_WORKSPACE,
                "b2s_conv_gather_gemm: workspace needs %zu bytes",
                tc_gather_gemm_workspace(k, c_in, c_out));
    launch_weight_to_kmajor(weight, k, c_in, c_out, ws, st);
    wt = reinterpret_cast<const __half*>(ws);
  }  // input gradient: B_k[n = c_in][c = c_out] = W[k][n][c] is the stored layout already
  if (step_rows) {
    // the step-table kernel addresses source rows by 32-bit byte offsets (b2s_conv_steps_supported)
    B2S_REQUIRE(step_start && tile_mask && n_src * (int64_t)c_red * 2 < (int64_t)0xFFFFFF00LL, B2S_ERR_INVALID,
                "b2s_conv_gather_gemm: step table without masks, or a source tensor of 4 GiB and more");
    return launch_gather_gemm_tc4(in, n_src, wt, k, c_red, c_res, flip_k, step_rows, step_start, tile_mask,
                                  tile_rows, row_perm, n_rows, bias, out, bn_sums, st);
  }
  B2S_REQUIRE(!bn_sums, B2S_ERR_UNSUPPORTED, "b2s_conv_gather_gemm: bn_sums needs the step-table kernel");
  return launch_gather_gemm_tc2(in, wt, k, c_red, c_res, flip_k, nbr, tile_mask, row_perm, n_rows, bias, out, st);
}

// =====================================================================================
// Weight gradient on tensor cores:  dW[k] (C_in x C_out) = sum over the pairs (i, o) of
// offset k of  X[i]^T * dY[o].   The reduction (GEMM-K) dimension is the pair list, so both
// operands are "MN-major": a gathered row (64 channels = 128 bytes) is one K-row of a
// SWIZZLE_128B MN-major panel.  The gather code is the one of the forward kernel; only the
// descriptors differ (a_major = b_major = MN, LBO = panel stride, SBO = 8 K-rows).
//
// Work unit = (offset k, run of <= unit_pairs pairs, 128-channel slab of C_in).  Persistent
// CTAs stride over the units; pair counts come from the device-resident nbsizes (no host
// sync).  Per unit the accumulator [128 lanes = c_in, C_out columns] lives in TMEM and is
// added into the fp32 dW with vector reds.
namespace tcw {
using namespace tc;

constexpr int kRows = 64;                    // pairs per pipeline stage (4 MMAs of K = 16)
constexpr int kPanelBytes = kRows * 128;     // one 64-channel panel of one stage

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t panel_stride_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((panel_stride_bytes >> 4) & 0x3FFF) << 16;   // LBO: next 64-channel panel
  d |= (uint64_t)(1024 >> 4) << 32;                            // SBO: next 8 K-rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                                      // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t make_idesc_mn(int n) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(kTileM >> 4) << 24);
}
__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

struct WParams {
  const __half* x;        // [n_in, c_in]
  const __half* gy;       // [n_out, c_out]
  const int32_t* pairs;   // [M, 2] (in, out) or nullptr (identity)
  const int32_t* nbsizes; // [n_seg] segment sizes (n_seg == K: the per-offset counts) or nullptr
  float* gw;              // [K, c_in, c_out]
  int64_t n_identity;
  int kvol, n_seg, c_in, c_out, swap_pairs;
  int m_tiles, unit_pairs, stages, tmem_cols;
  int dbg;                // ablation: bit0 no gathers, bit1 no MMAs, bit2 no epilogue reds
};

__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tm, int c0, int r0, int r1, int r2,
                                            int r3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(tm), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
      : "memory");
}

// kTma = false: rows copied with cp.async.  kTma = true (default for C_out <= 128, B2S_WG_GATHER4 forces):
// the producer warps hand the pair indices to the TMA unit, four rows per tile::gather4
// (tmx / tmy describe x and gy as [rows, channels] tensors with a one-row, 64-channel box; channels past
// the tensor width are zero-filled), and the stage barrier counts bytes instead of thread arrivals.
template <bool kTma>
__global__ void __launch_bounds__(kThreads) wgrad_tc_kernel(const WParams p, const __grid_constant__ CUtensorMap tmx,
                                                            const __grid_constant__ CUtensorMap tmy) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_panels = (p.c_out + 63) / 64;
  const int a_bytes = 2 * kPanelBytes;
  const int stage_stride = a_bytes + b_panels * kPanelBytes;       // multiples of 8 KiB
  __shared__ __align__(8) uint64_t s_full[8];
  __shared__ __align__(8) uint64_t s_empty[8];
  __shared__ __align__(8) uint64_t s_acc;
  __shared__ uint32_t s_tmem;
  __shared__ int32_t s_start[1025];        // exclusive prefix of the segment sizes (nbsizes when n_seg == K)
  __shared__ int32_t s_units[1025];        // exclusive prefix of units per segment

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.stages;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&s_full[s]), kTma ? 1 : kProducerThreads);
      mbar_init(smem_u32(&s_empty[s]), 1);
    }
    mbar_init(smem_u32(&s_acc), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    // exclusive prefixes over the n_seg segments: every lane sums a contiguous run, one warp scan joins them
    const int per = (p.n_seg + 31) / 32;
    const int s0 = lane * per, s1 = s0 + per < p.n_seg ? s0 + per : p.n_seg;
    int my_pairs = 0, my_units = 0;
    for (int sgm = s0; sgm < s1; ++sgm) {
      const int cnt = p.pairs ? __ldg(p.nbsizes + sgm) : (int)p.n_identity;
      my_pairs += cnt;
      my_units += (cnt + p.unit_pairs - 1) / p.unit_pairs;
    }
    int inc_p = my_pairs, inc_u = my_units;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int tp = __shfl_up_sync(0xffffffffu, inc_p, d), tu = __shfl_up_sync(0xffffffffu, inc_u, d);
      if (lane >= d) {
        inc_p += tp;
        inc_u += tu;
      }
    }
    int acc = inc_p - my_pairs, u = inc_u - my_units;
    for (int sgm = s0; sgm < s1; ++sgm) {
      s_start[sgm] = acc;
      s_units[sgm] = u;
      const int cnt = p.pairs ? __ldg(p.nbsizes + sgm) : (int)p.n_identity;
      acc += cnt;
      u += (cnt + p.unit_pairs - 1) / p.unit_pairs;
    }
    if (lane == 31) {
      s_start[p.n_seg] = inc_p;
      s_units[p.n_seg] = inc_u;
    }
  }
  if (warp == 4) tmem_alloc(smem_u32(&s_tmem), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = s_tmem;
  const int total_units = s_units[p.n_seg] * p.m_tiles;
  // C_out > 256: two MMAs per K step, N = 256 and N = C_out - 256 (the B operand of the second one starts
  // at the fifth 64-channel panel, so any C_out % 16 == 0 up to 512 works, e.g. 448 of RPVNet cr1.75)
  const int n_first = p.c_out > 256 ? 256 : p.c_out;

  // ring position shared by the roles (each role advances its own copy identically)
  int rs = 0, rwraps = 0;
  int unit_no = 0;       // units processed by this CTA (phase of s_acc)
  for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++unit_no) {
    const int mt = unit % p.m_tiles;
    const int ku = unit / p.m_tiles;
    int sg = 0;                                                     // segment of this unit: last s_units[sg] <= ku
    for (int hi_s = p.n_seg; hi_s - sg > 1;) {
      const int mid = (sg + hi_s) >> 1;
      if (s_units[mid] <= ku) sg = mid; else hi_s = mid;
    }
    const int k = sg % p.kvol;                                      // the offset the segment belongs to
    const int64_t cnt_k = s_start[sg + 1] - s_start[sg];
    const int64_t lo = (int64_t)(ku - s_units[sg]) * p.unit_pairs;
    const int64_t hi = lo + p.unit_pairs < cnt_k ? lo + p.unit_pairs : cnt_k;
    const int n_stage = (int)((hi - lo + kRows - 1) / kRows);
    const int ch_base = mt * 128;                                   // first c_in channel of the slab

    if (warp < 4) {
      // ---------------------------------------------------------------- producers
      const int sub = lane >> 3, chunk = lane & 7;
      const int2* pr = reinterpret_cast<const int2*>(p.pairs);
      const int64_t pair0 = s_start[sg];
      // Rows are addressed by 32-bit byte offsets from x / gy (the dispatcher keeps tensors of 4 GiB
      // and more on the SIMT kernel): one multiply per pair instead of 64-bit address arithmetic per
      // 16-byte copy, which made these warps latency-bound on their own instruction stream.
      constexpr uint32_t kNoRow = 0xFFFFFFFFu;
      const uint32_t x_row_bytes = (uint32_t)p.c_in * 2u, y_row_bytes = (uint32_t)p.c_out * 2u;
      const char* x_lane = reinterpret_cast<const char*>(p.x) + ch_base * 2 + chunk * 16;
      const char* y_lane = reinterpret_cast<const char*>(p.gy) + chunk * 16;
      // Pair indices are fetched two stages per load (lanes 0-15: stage st2, lanes 16-31: stage
      // st2+1; this warp stages rows [warp*16, warp*16+16) of every 64-pair stage) and two such
      // loads ahead of the gathers that consume them, so the index latency is off the issue path.
      auto load_idx = [&](int st2, uint32_t& ii, uint32_t& oo) {
        ii = kNoRow;
        oo = kNoRow;
        const int st = st2 + (lane >> 4);
        const int64_t q = lo + (int64_t)st * kRows + warp * 16 + (lane & 15);
        if (st < n_stage && q < hi) {
          if (pr) {
            const int2 v = __ldg(pr + pair0 + q);
            ii = (uint32_t)(p.swap_pairs ? v.y : v.x) * (kTma ? 1u : x_row_bytes);
            oo = (uint32_t)(p.swap_pairs ? v.x : v.y) * (kTma ? 1u : y_row_bytes);
          } else {
            ii = (uint32_t)q * (kTma ? 1u : x_row_bytes);
            oo = (uint32_t)q * (kTma ? 1u : y_row_bytes);
          }
        }
      };
      uint32_t i0, o0, i1, o1, i2, o2;
      load_idx(0, i0, o0);
      load_idx(2, i1, o1);
      for (int st2 = 0; st2 < n_stage; st2 += 2) {
        load_idx(st2 + 4, i2, o2);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (st2 + h >= n_stage) break;
          if (rwraps > 0) mbar_wait(smem_u32(&s_empty[rs]), (rwraps - 1) & 1);
          const uint32_t a_base = smem_base + rs * stage_stride;
          const uint32_t b_base = a_base + a_bytes;
          if constexpr (kTma) {
            // slots: the 1-2 live X panels, then the dY panels; lane = (slot % 8) * 4 + j fetches rows
            // [warp*16 + 4j, +4) of its slot(s); kNoRow is -1 as a coordinate: out of range -> zeros
            const int a_used = (p.c_in - ch_base > 64) ? 2 : 1;
            const int n_slots = a_used + b_panels;
            const uint32_t bar = smem_u32(&s_full[rs]);
            if (warp == 0 && lane == 0)
              mbar_arrive_expect_tx(bar, (p.dbg & 1) ? 0u : (uint32_t)(n_slots * kPanelBytes));
            const int j = lane & 3, src = h * 16 + j * 4;
            const int xi0 = (int)__shfl_sync(0xffffffffu, i0, src), xi1 = (int)__shfl_sync(0xffffffffu, i0, src + 1);
            const int xi2 = (int)__shfl_sync(0xffffffffu, i0, src + 2), xi3 = (int)__shfl_sync(0xffffffffu, i0, src + 3);
            const int yo0 = (int)__shfl_sync(0xffffffffu, o0, src), yo1 = (int)__shfl_sync(0xffffffffu, o0, src + 1);
            const int yo2 = (int)__shfl_sync(0xffffffffu, o0, src + 2), yo3 = (int)__shfl_sync(0xffffffffu, o0, src + 3);
            const uint32_t row_off = (uint32_t)(warp * 16 + j * 4) * 128u;
            for (int slot = lane >> 2; slot < n_slots && !(p.dbg & 1); slot += 8) {
              if (slot < a_used)
                tma_gather4(a_base + slot * kPanelBytes + row_off, &tmx, ch_base + slot * 64, xi0, xi1, xi2, xi3, bar);
              else
                tma_gather4(b_base + (slot - a_used) * kPanelBytes + row_off, &tmy, (slot - a_used) * 64, yo0, yo1,
                            yo2, yo3, bar);
            }
          } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (p.dbg & 1) break;
            const int rl = g * 4 + sub;                             // row within the warp's 16
            const int row = warp * 16 + rl;
            const uint32_t io = __shfl_sync(0xffffffffu, i0, h * 16 + rl);
            const uint32_t oo = __shfl_sync(0xffffffffu, o0, h * 16 + rl);
            const bool vi = io != kNoRow, vo = oo != kNoRow;
            const char* xa = x_lane + (vi ? io : 0u);
            const char* ya = y_lane + (vo ? oo : 0u);
            const uint32_t off = swz<128>(row, chunk);
            // A: two 64-channel panels of X
#pragma unroll
            for (int pn = 0; pn < 2; ++pn) {
              if (ch_base + pn * 64 + chunk * 8 < p.c_in)
                cp_async16(a_base + pn * kPanelBytes + off, xa + pn * 128, vi ? 16u : 0u);
            }
            // B: all panels of dY
            for (int pn = 0; pn < b_panels; ++pn) {
              if (pn * 64 + chunk * 8 < p.c_out)
                cp_async16(b_base + pn * kPanelBytes + off, ya + pn * 128, vo ? 16u : 0u);
            }
          }
          // the copy engine signals the stage when this thread's gathers have landed
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(
                           smem_u32(&s_full[rs]))
                       : "memory");
          }
          if (++rs == S) { rs = 0; ++rwraps; }
        }
        i0 = i1; o0 = o1;
        i1 = i2; o1 = o2;
      }

      // ----------------------------------------------------------------- epilogue
      mbar_wait(smem_u32(&s_acc), unit_no & 1);
      tc_fence_after();
      const int ci = ch_base + warp * 32 + lane;
      const uint32_t t_lane = tmem_acc + ((uint32_t)(warp * 32) << 16);
      float* dst = p.gw + ((int64_t)k * p.c_in + ci) * p.c_out;
      for (int c0 = 0; c0 < p.c_out; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(t_lane + (uint32_t)c0, v);
        tmem_ld_wait();
        if (ci < p.c_in && !(p.dbg & 4)) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            red_add_v4(dst + c0 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                       __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
      }
      tc_fence_before();
    } else {
      // --------------------------------------------------------------- MMA issuer
      const uint32_t idesc = make_idesc_mn(n_first);
      const uint32_t idesc2 = make_idesc_mn(p.c_out - n_first);
      for (int st = 0; st < n_stage; ++st) {
        mbar_wait(smem_u32(&s_full[rs]), rwraps & 1);
        fence_proxy_async();           // rows were written through the generic proxy (cp.async)
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_base = smem_base + rs * stage_stride;
          const uint32_t b_base = a_base + a_bytes;
#pragma unroll
          for (int kk = 0; kk < kRows / 16; ++kk) {
            if (p.dbg & 2) break;
            const uint64_t ad = make_desc_mn(a_base + kk * 2048, kPanelBytes);
            const uint64_t bd = make_desc_mn(b_base + kk * 2048, kPanelBytes);
            umma_f16(tmem_acc, ad, bd, idesc, (st | kk) ? 1u : 0u);
            if (n_first != p.c_out) {
              const uint64_t bd2 = make_desc_mn(b_base + (n_first / 64) * kPanelBytes + kk * 2048,
                                                kPanelBytes);
              umma_f16(tmem_acc + (uint32_t)n_first, ad, bd2, idesc2, (st | kk) ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(&s_empty[rs]));
          if (st == n_stage - 1) umma_commit(smem_u32(&s_acc));
        }
        __syncwarp();
        if (++rs == S) { rs = 0; ++rwraps; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, (uint32_t)p.tmem_cols);
  }
}

}  // namespace tcw

bool tc_wgrad_supported(int c_in, int c_out) {
  if (c_in % 8 != 0 || c_in < 8) return false;
  if (c_out % 16 != 0 || c_out < 16 || c_out > 512) return false;
  return true;                                                 // C_out > 256 splits as 256 + rest
}

bool tc_make_row_map(CUtensorMap* tm, const void* base, int64_t rows, int cols, int box_cols);   // conv_tc4.cu

namespace tcw {
template <bool kTma>
static cudaError_t launch_wgrad_variant(const WParams& p, const CUtensorMap& tmx, const CUtensorMap& tmy, int grid,
                                        size_t smem, cudaStream_t st) {
  static size_t opted_in = 0;
  if (smem > opted_in) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<kTma>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    opted_in = smem;
  }
  wgrad_tc_kernel<kTma><<<grid, kThreads, smem, st>>>(p, tmx, tmy);
  return cudaSuccess;
}
}  // namespace tcw

int launch_wgrad_tc(const void* in, int64_t n_in, const void* gout, int64_t n_out, const int32_t* nbmaps,
                    const int32_t* nbsizes, int n_seg, int64_t n_identity, int64_t n_pairs_bound, int k,
                    int c_in, int c_out, int swap_pairs, float* gw, cudaStream_t st) {
  using namespace tcw;
  B2S_REQUIRE(k <= 128 && n_seg >= 1 && n_seg <= 1024 && n_seg % k == 0, B2S_ERR_UNSUPPORTED,
              "b2s_conv_wgrad: kernel volume %d > 128 or %d segments", k, n_seg);
  WParams p;
  p.x = reinterpret_cast<const __half*>(in);
  p.gy = reinterpret_cast<const __half*>(gout);
  p.pairs = nbmaps;
  p.nbsizes = nbsizes;
  p.gw = gw;
  p.n_identity = n_identity;
  p.kvol = k;
  p.n_seg = n_seg;
  p.c_in = c_in;
  p.c_out = c_out;
  p.swap_pairs = swap_pairs;
  p.m_tiles = (c_in + 127) / 128;
  p.dbg = 0;
  {
    const char* ed = getenv("B2S_WG_DBG");
    if (ed) p.dbg = atoi(ed);
  }
  p.tmem_cols = tc::tmem_cols_for(c_out);
  const int stage = (2 + (c_out + 63) / 64) * kPanelBytes;
  // CTAs per SM: as many as fit with a 2-stage ring each (every CTA is one independent gather -> MMA -> commit
  // chain, and the chains, not the ring depth, are what hide the hand-shake and gather latencies: 256-channel
  // layers went 150 -> 109 us from 1 x 4 stages to 2 x 2 stages, profiles/r2_wgrad_ctas.txt).  Budget: 228 KiB
  // per SM, per CTA 1 KiB reserved + ~8.5 KiB static (segment prefixes, barriers) + 1 KiB alignment slack.
  static const int max_ctas = [] {
    const char* e = getenv("B2S_WG_CTAS");
    return e ? atoi(e) : 3;
  }();
  int ctas = 1, stages = 0;
  for (int c = max_ctas < 1 ? 1 : (max_ctas > 4 ? 4 : max_ctas); c >= 1; --c) {
    const int st = (int)(((228 * 1024) / c - 11 * 1024) / stage);
    if ((st >= 2 && c * p.tmem_cols <= 512) || c == 1) {
      ctas = c;
      stages = st;
      break;
    }
  }
  if (stages > 6) stages = 6;
  B2S_REQUIRE(stages >= 2, B2S_ERR_UNSUPPORTED, "b2s_conv_wgrad: tile does not fit (C_out=%d)", c_out);
  p.stages = stages;
  // unit size: ~6 units per CTA slot so that the static round-robin stays balanced even though the
  // per-offset pair counts differ by 10x (the centre offset has N pairs, corner offsets ~N/20);
  // at least 16 stages per unit so the fp32 reds of the epilogue stay a few % of the gathered bytes.
  // (the true pair count lives on the device; on LiDAR surfaces ~1/4 of the K*N slots exist)
  const int sms = persistent_sms();
  const int64_t est = k > 1 ? n_pairs_bound / 4 + 1 : n_pairs_bound;
  const int64_t slots_est = (int64_t)sms * ctas;
  int64_t per = est / (slots_est * 5) + 1;
  int64_t unit = ((per + kRows - 1) / kRows) * kRows;
  if (unit < 16 * kRows) unit = 16 * kRows;
  if (unit > 256 * kRows) unit = 256 * kRows;
  p.unit_pairs = (int)unit;
  const size_t smem = (size_t)stages * stage + 1024;
  int64_t max_units = (n_pairs_bound / unit + n_seg) * p.m_tiles;
  const int64_t slots = (int64_t)sms * ctas;
  int grid = (int)(max_units < slots ? (max_units < 1 ? 1 : max_units) : slots);
  // Both operands fetched by the TMA unit (tile::gather4) where it measured faster than cp.async
  // (profiles/r2_gather4_validation.txt, batch 4: C_out <= 128 layers 3-23 % faster, 256-channel layers
  // 3-6 % slower: the unit sustains one 4-row gather per ~13 cycles per SM); B2S_WG_GATHER4=0/1 forces.
  static const int g4_env = [] {
    const char* e = getenv("B2S_WG_GATHER4");
    return e ? (e[0] == '1' ? 1 : 0) : -1;
  }();
  const bool gather4 = g4_env >= 0 ? g4_env == 1 : (c_out <= 128 && c_in % 8 == 0);
  CUtensorMap tmx, tmy;
  memset(&tmx, 0, sizeof(tmx));
  memset(&tmy, 0, sizeof(tmy));
  if (gather4) {
    B2S_REQUIRE(c_in % 8 == 0 && c_out % 8 == 0 && tc_make_row_map(&tmx, in, n_in, c_in, 64) &&
                    tc_make_row_map(&tmy, gout, n_out, c_out, 64),
                B2S_ERR_CUDA, "b2s_conv_wgrad: cuTensorMapEncodeTiled failed");
  }
  cudaError_t e = gather4 ? launch_wgrad_variant<true>(p, tmx, tmy, grid, smem, st)
                          : launch_wgrad_variant<false>(p, tmx, tmy, grid, smem, st);
  B2S_REQUIRE(e == cudaSuccess, B2S_ERR_CUDA, "b2s_conv_wgrad: cannot opt in to %zu B smem: %s", smem,
              cudaGetErrorString(e));
  return B2S_OK;
}

}  // namespace b2s
