// Sparse convolution gather-GEMM, tensor-core family, revision 2 (fp16, sm_100a).
//
// Same output-stationary formulation as conv_tc.cu (one CTA = 128 output rows, all kernel
// offsets accumulated in one TMEM accumulator), with the producer side rebuilt after the
// first ncu capture (profiles/r1a_*): the v1 producers were issue/latency bound (tensor
// pipe 12-27 % active, integer divisions and weight-tile copies in the gather loop).
//
//   warps 0-3  A producers: cp.async 16 B gathers of the neighbour rows (zero-fill for
//              missing neighbours) into SWIZZLE_128B (64-channel chunks) or SWIZZLE_64B
//              (32-channel tail chunk) tiles; later the epilogue warps
//   warp  4    MMA issuer (tcgen05.mma kind::f16, cta_group::1, M=128, N=C_res)
//   warp  5    B producer: ONE TMA tensor-tile load per stage brings the weight slice
//              W_k^T[:, chunk] with the hardware swizzle (cp.async.bulk.tensor.2d +
//              mbarrier complete_tx) - no LSU instructions spent on weights
//
// Ring bookkeeping is incremental (no div/mod), the neighbour index of the next active
// offset is prefetched while the current one is gathered, and a 96- or 32-channel reduction
// uses one 64-wide chunk + one 32-wide tail chunk instead of three 32-wide ones.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace b2s {
namespace tc2 {
using namespace tc;

constexpr int kThreads2 = 192;
constexpr int kABytes = kTileM * 128;       // A region of a stage (SW128 worst case)

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(tm), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

struct Params {
  const __half* in;     // [n_src, c_red]
  const int32_t* nbr;   // [K][n_rows] or nullptr (identity, K == 1)
  const uint32_t* tile_mask;  // [tiles][ceil(K/32)] active-offset bits of nbr's tiles, or nullptr
  const int32_t* row_perm;    // out row of launch row j, or nullptr
  const __half* bias;   // [c_res] or nullptr
  __half* out;          // [n_rows, c_res]
  int64_t n_rows;
  int kvol, c_red, c_res, flip_k;
  int n64, tail32;      // c_red = 64 * n64 + 32 * tail32
  int dbg;              // bit0: zero-fill all A rows, bit1: no A copies, bit2: no B TMA, bit3: no MMA
  int consumer_fence;   // debugging aid: generic->async proxy fence in the MMA warp per stage
  int sa, sb;           // ring depths: A (gathered rows, deep: hides the gather latency), B (weights)
  int b_stride;         // bytes of one B slot (multiple of 1024)
  int tmem_cols;
};

struct Ring {           // ring position without div/mod
  int s = 0;            // stage index
  int wraps = 0;        // completed passes over the ring
  int S;
  __device__ explicit Ring(int stages) : S(stages) {}
  __device__ __forceinline__ void advance() {
    if (++s == S) {
      s = 0;
      ++wraps;
    }
  }
};

template <int ROWB>
__device__ __forceinline__ void gather_chunk(const Params& p, int32_t my_src, uint32_t a_base,
                                             int col0, int warp, int lane) {
  constexpr int CH = ROWB / 16, RPI = 32 / CH;
  const int sub = lane / CH, chunk = lane % CH;
  const __half* base = p.in + col0 + chunk * 8;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int rl = i * RPI + sub;
    int32_t src = __shfl_sync(0xffffffffu, my_src, rl);
    if (p.dbg & 1) src = -1;
    if (p.dbg & 2) continue;
    const __half* g = src >= 0 ? base + (int64_t)src * p.c_red : p.in;
    cp_async16(a_base + swz<ROWB>(warp * 32 + rl, chunk), g, src >= 0 ? 16u : 0u);
  }
}

__device__ __forceinline__ int32_t load_src(const Params& p, int k, int64_t r) {
  if (r >= p.n_rows) return -1;
  if (!p.nbr) return (int32_t)r;
  return __ldg(p.nbr + (int64_t)(p.flip_k ? p.kvol - 1 - k : k) * p.n_rows + r);
}

struct KMask {   // active-offset bits of one tile, kept in registers (K <= 128)
  uint32_t w0, w1, w2, w3;
};
__device__ __forceinline__ int first_above(uint32_t bits, int lo) {   // lowest set bit >= lo, or -1
  if (lo >= 32) return -1;
  if (lo > 0) bits &= 0xFFFFFFFFu << lo;
  return bits ? __ffs(bits) - 1 : -1;
}
__device__ __forceinline__ int next_active(const KMask& m, int k, int kvol) {
  // smallest active offset > k, or kvol
  int b;
  if ((b = first_above(m.w0, k + 1)) >= 0) return b;
  if (kvol > 32 && (b = first_above(m.w1, k + 1 - 32)) >= 0) return 32 + b;
  if (kvol > 64 && (b = first_above(m.w2, k + 1 - 64)) >= 0) return 64 + b;
  if (kvol > 96 && (b = first_above(m.w3, k + 1 - 96)) >= 0) return 96 + b;
  return kvol;
}

__global__ void __launch_bounds__(kThreads2) gather_gemm_tc2_kernel(
    const Params p, const __grid_constant__ CUtensorMap tm64, const __grid_constant__ CUtensorMap tm32) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t s_full[8];     // A slots: 128 gather-thread arrivals
  __shared__ __align__(8) uint64_t s_empty[8];
  __shared__ __align__(8) uint64_t s_fullb[16];   // B slots: TMA expect_tx
  __shared__ __align__(8) uint64_t s_emptyb[16];
  __shared__ __align__(8) uint64_t s_acc;
  __shared__ uint32_t s_tmem;
  __shared__ uint32_t s_active[4];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * kTileM;
  const int S = p.sa;
  const uint32_t smem_b = smem_base + (uint32_t)p.sa * kABytes;

  if (tid < 4) s_active[tid] = 0;
  if (tid == 0) {
    for (int s = 0; s < p.sa; ++s) {
      mbar_init(smem_u32(&s_full[s]), kProducerThreads);
      mbar_init(smem_u32(&s_empty[s]), 1);
    }
    for (int s = 0; s < p.sb; ++s) {
      mbar_init(smem_u32(&s_fullb[s]), 1);                      // TMA expect_tx arrival
      mbar_init(smem_u32(&s_emptyb[s]), 1);
    }
    mbar_init(smem_u32(&s_acc), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(smem_u32(&s_tmem), (uint32_t)p.tmem_cols);
  __syncthreads();
  if (p.tile_mask) {
    // precomputed by b2s_kmap_build: one word per 32 offsets (bit-reversed use when flip_k)
    const int words = (p.kvol + 31) >> 5;
    const uint32_t* tm = p.tile_mask + (int64_t)blockIdx.x * words;
    if (!p.flip_k) {
      if (tid < words) s_active[tid] = __ldg(tm + tid);
    } else if (tid < p.kvol) {
      const int kk = p.kvol - 1 - tid;
      if ((__ldg(tm + (kk >> 5)) >> (kk & 31)) & 1u) atomicOr(&s_active[tid >> 5], 1u << (tid & 31));
    }
  } else if (warp < 4) {
    const int64_t r = row0 + tid;
    for (int k = 0; k < p.kvol; ++k) {
      const bool have = load_src(p, k, r) >= 0;
      const unsigned m = __ballot_sync(0xffffffffu, have);
      if (lane == 0 && m) atomicOr(&s_active[k >> 5], 1u << (k & 31));
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = s_tmem;
  const KMask amask = {s_active[0], s_active[1], s_active[2], s_active[3]};
  const int k_first = next_active(amask, -1, p.kvol);
  const bool any = k_first < p.kvol;

  if (warp < 4) {
    // ---------------------------------------------------------------- A producers
    // Completion of a stage is signalled by the copy engine itself
    // (cp.async.mbarrier.arrive.noinc: the arrive fires when this thread's prior cp.asyncs have
    // landed), so the gather warps never wait on their own loads: the only blocking point is
    // a full ring.  (Measured alternatives, profiles/: per-stage cp.async.wait_group + one
    // arrive per warp is ~25 % slower.)
    Ring ring(S);
    const int64_t my_row = row0 + warp * 32 + lane;
    int32_t nxt = any ? load_src(p, k_first, my_row) : -1;
    for (int k = k_first; k < p.kvol;) {
      const int32_t src = nxt;
      const int kn = next_active(amask, k, p.kvol);
      if (kn < p.kvol) nxt = load_src(p, kn, my_row);          // prefetch the next offset's map
      const int n_chunks = p.n64 + p.tail32;
      for (int c = 0; c < n_chunks; ++c) {
        if (ring.wraps > 0) mbar_wait(smem_u32(&s_empty[ring.s]), (ring.wraps - 1) & 1);
        const uint32_t a_base = smem_base + ring.s * kABytes;
        if (c < p.n64) gather_chunk<128>(p, src, a_base, c * 64, warp, lane);
        else gather_chunk<64>(p, src, a_base, p.n64 * 64, warp, lane);
        cp_async_mbar_arrive_noinc(smem_u32(&s_full[ring.s]));
        ring.advance();
      }
      k = kn;
    }

    // ------------------------------------------------------------------- epilogue
    if (any) {
      mbar_wait(smem_u32(&s_acc), 0);
      tc_fence_after();
    }
    const int64_t r = my_row;
    const int64_t r_out = (p.row_perm && r < p.n_rows) ? (int64_t)__ldg(p.row_perm + r) : r;
    const uint32_t t_lane = tmem_acc + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < p.c_res; c0 += 16) {
      uint32_t v[16];
      if (any) {
        tmem_ld16(t_lane + (uint32_t)c0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
      if (r < p.n_rows) {
        __align__(16) __half h[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float f = __uint_as_float(v[j]);
          if (p.bias) f += __half2float(__ldg(p.bias + c0 + j));
          h[j] = __float2half_rn(f);
        }
        uint4* dst = reinterpret_cast<uint4*>(p.out + r_out * p.c_res + c0);
        dst[0] = reinterpret_cast<const uint4*>(h)[0];
        dst[1] = reinterpret_cast<const uint4*>(h)[1];
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // ----------------------------------------------------------------- MMA issuer
    const int n_half = p.c_res > 256 ? p.c_res / 2 : p.c_res;
    const uint32_t idesc = make_idesc(n_half);
    Ring ring(S), rb(p.sb);
    uint32_t acc_flag = 0;
    for (int k = k_first; k < p.kvol; k = next_active(amask, k, p.kvol)) {
      const int kn = next_active(amask, k, p.kvol);
      const int n_chunks = p.n64 + p.tail32;
      for (int c = 0; c < n_chunks; ++c) {
        if (lane == 0) {
          mbar_wait(smem_u32(&s_fullb[rb.s]), rb.wraps & 1);
          mbar_wait(smem_u32(&s_full[ring.s]), ring.wraps & 1);
          if (p.consumer_fence) fence_proxy_async();   // off by default, see launch_gather_gemm_tc2
          tc_fence_after();
          const uint32_t a_base = smem_base + ring.s * kABytes;
          const uint32_t b_base = smem_b + rb.s * p.b_stride;
          if (p.dbg & 8) {
          } else if (c < p.n64) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t ad = make_desc<128>(a_base + kk * 32);
              umma_f16(tmem_acc, ad, make_desc<128>(b_base + kk * 32), idesc, acc_flag);
              if (n_half != p.c_res)
                umma_f16(tmem_acc + (uint32_t)n_half, ad,
                         make_desc<128>(b_base + n_half * 128 + kk * 32), idesc, acc_flag);
              acc_flag = 1;
            }
          } else {
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint64_t ad = make_desc<64>(a_base + kk * 32);
              umma_f16(tmem_acc, ad, make_desc<64>(b_base + kk * 32), idesc, acc_flag);
              if (n_half != p.c_res)
                umma_f16(tmem_acc + (uint32_t)n_half, ad,
                         make_desc<64>(b_base + n_half * 64 + kk * 32), idesc, acc_flag);
              acc_flag = 1;
            }
          }
          if (p.dbg & 16) {
            mbar_arrive(smem_u32(&s_empty[ring.s]));
            mbar_arrive(smem_u32(&s_emptyb[rb.s]));
            if (kn >= p.kvol && c == n_chunks - 1) mbar_arrive(smem_u32(&s_acc));
          } else {
            umma_commit(smem_u32(&s_empty[ring.s]));
            umma_commit(smem_u32(&s_emptyb[rb.s]));
            if (kn >= p.kvol && c == n_chunks - 1) umma_commit(smem_u32(&s_acc));
          }
        }
        __syncwarp();
        ring.advance();
        rb.advance();
      }
    }
    tc_fence_before();
  } else {
    // ------------------------------------------------------- B producer (TMA, warp 5)
    if (lane == 0) {
      const int n_half = p.c_res > 256 ? p.c_res / 2 : p.c_res;
      Ring ring(p.sb);
      for (int k = k_first; k < p.kvol; k = next_active(amask, k, p.kvol)) {
        const int n_chunks = p.n64 + p.tail32;
        for (int c = 0; c < n_chunks; ++c) {
          if (ring.wraps > 0) mbar_wait(smem_u32(&s_emptyb[ring.s]), (ring.wraps - 1) & 1);
          const uint32_t bar = smem_u32(&s_fullb[ring.s]);
          const uint32_t b_base = smem_b + ring.s * p.b_stride;
          const bool wide = c < p.n64;
          const int rowb = wide ? 128 : 64;
          if (p.dbg & 4) {
            mbar_arrive(bar);
          } else {
            mbar_arrive_expect_tx(bar, (uint32_t)(p.c_res * rowb));
            const CUtensorMap* tm = wide ? &tm64 : &tm32;
            const int col = wide ? c * 64 : p.n64 * 64;
            tma_load_2d(b_base, tm, col, k * p.c_res, bar);
            if (n_half != p.c_res) tma_load_2d(b_base + n_half * rowb, tm, col, k * p.c_res + n_half, bar);
          }
          ring.advance();
        }
      }
    }
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, (uint32_t)p.tmem_cols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D map over W^T viewed as [K * c_res rows, c_red cols] fp16; box = box_cols x box_rows
static bool make_weight_map(CUtensorMap* tm, const void* w, int k, int c_res, int c_red, int box_cols,
                            int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)c_red, (cuuint64_t)k * c_res};
  cuuint64_t strides[1] = {(cuuint64_t)c_red * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace tc2

// wt: [K][c_res][c_red] fp16 (K-major B operand), already transposed by the caller if needed
int launch_gather_gemm_tc2(const void* in, const void* wt, int k, int c_red, int c_res, int flip_k,
                           const int32_t* nbr, const uint32_t* tile_mask, const int32_t* row_perm,
                           int64_t n_rows, const void* bias, void* out, cudaStream_t st) {
  using namespace tc2;
  Params p;
  p.in = reinterpret_cast<const __half*>(in);
  p.nbr = nbr;
  p.tile_mask = nbr ? tile_mask : nullptr;
  p.row_perm = row_perm;
  p.bias = reinterpret_cast<const __half*>(bias);
  p.out = reinterpret_cast<__half*>(out);
  p.n_rows = n_rows;
  p.kvol = k;
  p.c_red = c_red;
  p.c_res = c_res;
  p.flip_k = flip_k;
  p.n64 = c_red / 64;
  p.tail32 = (c_red % 64) ? 1 : 0;
  p.tmem_cols = 32;
  while (p.tmem_cols < c_res) p.tmem_cols <<= 1;
  const int n_half = c_res > 256 ? c_res / 2 : c_res;
  CUtensorMap tm64, tm32;
  memset(&tm64, 0, sizeof(tm64));
  memset(&tm32, 0, sizeof(tm32));
  if (p.n64) B2S_REQUIRE(make_weight_map(&tm64, wt, k, c_res, c_red, 64, n_half), B2S_ERR_CUDA,
                         "b2s_conv_gather_gemm: cuTensorMapEncodeTiled failed (64-wide)");
  if (p.tail32) B2S_REQUIRE(make_weight_map(&tm32, wt, k, c_res, c_red, 32, n_half), B2S_ERR_CUDA,
                            "b2s_conv_gather_gemm: cuTensorMapEncodeTiled failed (32-wide)");
  const int rowb_max = p.n64 ? 128 : 64;
  p.b_stride = (((c_res + 7) / 8) * 8 * rowb_max + 1023) & ~1023;
  // Two CTAs per SM (epilogue of one overlaps the main loop of the other) when a 5-deep
  // A ring + 2 weight slots fit twice; otherwise one CTA per SM with the deepest rings.
  const int budget2 = 112 * 1024, budget1 = 224 * 1024;
  if (p.tmem_cols <= 256 && 4 * kABytes + 2 * p.b_stride + 1024 <= budget2) {
    p.sa = 4;      // measured (profiles/microbench): 2 CTAs x 4-deep beats 1 CTA x 8-deep by ~1.7x
    p.sb = 2;
    while (p.sa < 6 && (p.sa + 1) * kABytes + p.sb * p.b_stride + 1024 <= budget2) ++p.sa;
  } else {
    p.sb = 3;
    p.sa = (budget1 - 1024 - p.sb * p.b_stride) / kABytes;
    if (p.sa < 4) {
      p.sb = 2;
      p.sa = (budget1 - 1024 - p.sb * p.b_stride) / kABytes;
    }
    if (p.sa > 8) p.sa = 8;
  }
  // The gathered rows are written by cp.async (LDGSTS) and completion reaches the MMA warp through
  // cp.async.mbarrier.arrive + mbarrier wait - the same hand-off CUTLASS's sm100 cp.async/UMMA
  // mainloop uses, without a proxy fence.  A per-stage fence.proxy.async in the MMA warp costs
  // ~1 us per stage (measured) and is therefore only available for A/B debugging.
  p.consumer_fence = 0;
  p.dbg = 0;
  {
    const char* ed = getenv("B2S_TC_DBG");
    if (ed) p.dbg = atoi(ed);
  }
  {
    const char* ef = getenv("B2S_TC_FENCE");
    if (ef && ef[0] == '1') p.consumer_fence = 1;
  }
  {  // tuning overrides (microbenchmarks only)
    const char* ea = getenv("B2S_TC_SA");
    const char* eb = getenv("B2S_TC_SB");
    if (ea && atoi(ea) >= 2 && atoi(ea) <= 8) p.sa = atoi(ea);
    if (eb && atoi(eb) >= 1 && atoi(eb) <= 16) p.sb = atoi(eb);
    while ((size_t)p.sa * kABytes + (size_t)p.sb * p.b_stride + 1024 > (size_t)budget1 && p.sa > 2) --p.sa;
  }
  B2S_REQUIRE(p.sa >= 2, B2S_ERR_UNSUPPORTED, "b2s_conv_gather_gemm: tile does not fit (C=%d)", c_res);
  const size_t smem = (size_t)p.sa * kABytes + (size_t)p.sb * p.b_stride + 1024;
  static size_t smem_opt_in = 0;   // one process drives one GPU: grow the opt-in limit lazily
  if (smem > smem_opt_in) {
    cudaError_t e = cudaFuncSetAttribute(gather_gemm_tc2_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    B2S_REQUIRE(e == cudaSuccess, B2S_ERR_CUDA, "b2s_conv_gather_gemm: cannot opt in to %zu B smem: %s",
                smem, cudaGetErrorString(e));
    smem_opt_in = smem;
  }
  const unsigned grid = (unsigned)ceil_div(n_rows, kTileM);
  gather_gemm_tc2_kernel<<<grid, kThreads2, smem, st>>>(p, tm64, tm32);
  return B2S_OK;
}

}  // namespace b2s
