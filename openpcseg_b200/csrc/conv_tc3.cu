// Sparse convolution gather-GEMM, tensor-core family, revision 3: PERSISTENT CTAs (fp16, sm_100a).
//
// Why (ablation of revision 2, profiles/r1_conv_ablation.txt): with copies, TMA and MMAs all
// switched off, the v2 kernel still needed ~2/3 of its time - ~10 us of fixed cost per 128-row
// tile (CTA launch, barrier init, TMEM alloc, mask scan, epilogue, TMEM free, exit) and ~440
// cycles of mbarrier hand-shakes per pipeline stage in the single MMA-issuing thread (two waits
// + two commits).  Revision 3 removes both:
//
//   * one CTA per SM slot loops over tiles (static stride); barriers and TMEM are set up once;
//   * the accumulator is double-buffered in TMEM (2 x C_res columns) and drained by four
//     dedicated epilogue warps, so the epilogue of tile i overlaps the main loop of tile i+1;
//   * one "full" barrier per stage (128 gather arrivals + the TMA transaction bytes) and one
//     "empty" barrier (one tcgen05.commit): the MMA thread does 1 wait + 1 commit per stage.
//
// Revision 3b - T = 2 row tiles (256 rows) per CTA share each TMA weight tile (two MMAs, separate TMEM
// accumulators).  It was motivated by a first reading of the v3 profile (5.8-6.2 TB/s of L2->SM reads,
// 2/3 of them weights); the later skeleton ablation (all copies, MMAs and stores off: 70 % of the time
// remains) showed the limiter is the instruction stream of the gather warps per (tile, offset) step, not
// L2 bandwidth, which is why T = 2 only pays where a step is a single 32-channel stage (C_res <= 32).
//
// Revision 3e/3f - tiles are composed by neighbourhood pattern (row_perm from b2s_tile_order_key: 25 % of
// the (tile, offset) steps stay active at stride 1), the walk over active steps is flattened with the
// map entries prefetched three steps ahead, source rows are addressed by 32-bit byte offsets, waiting
// roles use try_wait suspend hints.  An experimental variant hands the row gathers to the TMA unit
// (tile::gather4, kTmaGather).
//
// Warp roles (6 + 4T warps): 0..4T-1 gather producers (cp.async 16 B, zero-fill, SW128 / SW64 tiles),
// 4 MMA issuer (+ TMEM alloc), 5 TMA weight-tile producer, 6-9 epilogue (TMEM -> fp16 rows).
// Requires the per-tile active-offset masks that b2s_kmap_build emits.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace b2s {
namespace tc3 {
using namespace tc;

constexpr int kABytes = kTileM * 128;

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(tm), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// Four rows of a 2-D tensor (row indices r0..r3, column c0) land as four consecutive smem rows in the
// tensor map's swizzle mode; rows outside [0, n_rows) - the map's -1 entries - are zero-filled and
// still count their bytes on the barrier (checked by scripts/gather4_probe.cu).
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tm, int c0, int r0, int r1, int r2,
                                            int r3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(tm), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(256);
  }
}

struct Params {
  const __half* in;           // [n_src, c_red]
  const int32_t* nbr;         // [K][n_rows]
  const uint32_t* tile_mask;  // [tiles][ceil(K/32)]
  const int32_t* row_perm;    // out row of launch row j, or nullptr
  const __half* bias;         // [c_res] or nullptr
  __half* out;                // [n_rows, c_res]
  int64_t n_rows;
  int n_tiles;                // CTA tiles (128*T rows each)
  int n_tiles128;             // 128-row tiles (granularity of tile_mask)
  int kvol, c_red, c_res, flip_k;
  int n64, tail32;            // c_red = 64 * n64 + 32 * tail32
  int stages, stage_stride;   // ring of (A tile, weight tile) pairs
  int acc_stride, acc_bufs;   // TMEM columns per accumulator, 1 or 2 accumulators
  int tmem_cols;
  int dbg;                    // ablation (B2S_TC3_DBG): 1 no gathers, 2 no weight TMA, 4 no MMAs, 8 no stores, 16 no map loads
};

struct Ring {
  int s = 0, wraps = 0, S;
  __device__ explicit Ring(int stages) : S(stages) {}
  __device__ __forceinline__ void advance() {
    if (++s == S) {
      s = 0;
      ++wraps;
    }
  }
};

// active-offset bits of one tile (K <= 128), in the iteration order of this launch
struct KMask {
  uint32_t w0, w1, w2, w3;
  __device__ __forceinline__ bool any() const { return (w0 | w1 | w2 | w3) != 0; }
};
__device__ __forceinline__ int first_above(uint32_t bits, int lo) {
  if (lo >= 32) return -1;
  if (lo > 0) bits &= 0xFFFFFFFFu << lo;
  return bits ? __ffs(bits) - 1 : -1;
}
__device__ __forceinline__ int next_active(const KMask& m, int k, int kvol) {
  int b;
  if ((b = first_above(m.w0, k + 1)) >= 0) return b;
  if (kvol > 32 && (b = first_above(m.w1, k + 1 - 32)) >= 0) return 32 + b;
  if (kvol > 64 && (b = first_above(m.w2, k + 1 - 64)) >= 0) return 64 + b;
  if (kvol > 96 && (b = first_above(m.w3, k + 1 - 96)) >= 0) return 96 + b;
  return kvol;
}
__device__ __forceinline__ KMask load_mask(const Params& p, int tile) {
  const int words = (p.kvol + 31) >> 5;
  const uint32_t* tm = p.tile_mask + (int64_t)tile * words;
  KMask m = {__ldg(tm), 0u, 0u, 0u};
  if (words > 1) m.w1 = __ldg(tm + 1);
  if (words > 2) m.w2 = __ldg(tm + 2);
  if (words > 3) m.w3 = __ldg(tm + 3);
  if (p.flip_k) {                      // bit k of the flipped mask = bit K-1-k of the stored one
    if (p.kvol <= 32) {
      m.w0 = __brev(m.w0) >> (32 - p.kvol);
    } else {
      KMask r = {0u, 0u, 0u, 0u};
      for (int k = 0; k < p.kvol; ++k) {
        const int kk = p.kvol - 1 - k;
        const uint32_t src = kk < 32 ? m.w0 : kk < 64 ? m.w1 : kk < 96 ? m.w2 : m.w3;
        if ((src >> (kk & 31)) & 1u) {
          const uint32_t bit = 1u << (k & 31);
          if (k < 32) r.w0 |= bit; else if (k < 64) r.w1 |= bit; else if (k < 96) r.w2 |= bit; else r.w3 |= bit;
        }
      }
      m = r;
    }
  }
  return m;
}

// Gather bookkeeping of one producer thread, set up ONCE per kernel offset and reused by every
// channel chunk of that offset: the global row pointers of the rows this lane copies (8 row
// slots for 128-byte tiles, 4 for the 64-byte tail tile), a bit per slot telling whether the
// neighbour exists (missing ones are zero-filled by cp.async with src-size 0), and the constant
// swizzled shared-memory offsets.  The per-stage work is then one add + one LDGSTS per slot - the
// first persistent version recomputed shuffles and 64-bit address math per stage and was bound by
// instruction issue in these four warps (profiles/r1_conv_ablation.txt).
struct GatherSlots {
  const char* base128;        // in + this lane's 16-byte column of a 128-byte row chunk
  const char* base64;         // in + first tail channel + this lane's column of a 64-byte row chunk
  uint32_t o128[8];           // byte offset of the source row of each slot, kNoRow = no neighbour
  uint32_t o64[4];
  uint32_t off128[8];         // swizzled shared-memory offsets (constant)
  uint32_t off64[4];
};
constexpr uint32_t kNoRow = 0xFFFFFFFFu;

__device__ __forceinline__ void slots_init(GatherSlots& g, const Params& p, int warp, int lane) {
  g.base128 = reinterpret_cast<const char*>(p.in) + (lane & 7) * 16;
  g.base64 = reinterpret_cast<const char*>(p.in) + p.n64 * 128 + (lane & 3) * 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = warp * 32 + i * 4 + (lane >> 3);          // 0 .. 128T-1
    g.off128[i] = (uint32_t)(row >> 7) * kABytes + swz<128>(row & 127, lane & 7);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = warp * 32 + i * 8 + (lane >> 2);
    g.off64[i] = (uint32_t)(row >> 7) * kABytes + swz<64>(row & 127, lane & 3);
  }
}

// my_src = map entry of row (warp * 32 + lane).  Rows are addressed by 32-bit byte offsets from
// `in` (the launcher routes inputs of 4 GiB and more to the v2 kernel): one multiply per lane, then
// one shuffle per slot - the 64-bit pointer per slot of the first version was 9 instructions per
// slot and made these warps issue-bound (profiles/r1_conv_ablation.txt).
__device__ __forceinline__ void slots_set_rows(GatherSlots& g, const Params& p, int32_t my_src, int lane) {
  const uint32_t my_off = my_src >= 0 ? (uint32_t)my_src * (uint32_t)(p.c_red * 2) : kNoRow;
  if (p.n64) {
#pragma unroll
    for (int i = 0; i < 8; ++i) g.o128[i] = __shfl_sync(0xffffffffu, my_off, i * 4 + (lane >> 3));
  }
  if (p.tail32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) g.o64[i] = __shfl_sync(0xffffffffu, my_off, i * 8 + (lane >> 2));
  }
}

__device__ __forceinline__ void gather_wide(const GatherSlots& g, uint32_t a_base, int col) {
  const char* b = g.base128 + col * 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool ok = g.o128[i] != kNoRow;
    cp_async16(a_base + g.off128[i], b + (ok ? g.o128[i] : 0u), ok ? 16u : 0u);
  }
}
__device__ __forceinline__ void gather_tail(const GatherSlots& g, uint32_t a_base) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool ok = g.o64[i] != kNoRow;
    cp_async16(a_base + g.off64[i], g.base64 + (ok ? g.o64[i] : 0u), ok ? 16u : 0u);
  }
}

__device__ __forceinline__ int32_t load_src(const Params& p, int k, int64_t r) {
  if (r >= p.n_rows) return -1;
  return __ldg(p.nbr + (int64_t)(p.flip_k ? p.kvol - 1 - k : k) * p.n_rows + r);
}

// union of the active-offset masks of the T row tiles of a CTA tile
template <int T>
__device__ __forceinline__ KMask load_mask_t(const Params& p, int ctile) {
  KMask m = load_mask(p, ctile * T);
  if (T == 2 && ctile * T + 1 < p.n_tiles128) {
    const KMask m2 = load_mask(p, ctile * T + 1);
    m.w0 |= m2.w0;
    m.w1 |= m2.w1;
    m.w2 |= m2.w2;
    m.w3 |= m2.w3;
  }
  return m;
}

// kTmaGather = false: the gather warps copy rows with cp.async (16 bytes per thread and instruction).
// kTmaGather = true (experimental, B2S_TC_GATHER4=1): the same warps only hand row indices to the TMA
// unit, four rows per tile::gather4 instruction (ta64 / ta32 describe `in` as a [n_src, c_red] tensor
// with 64- / 32-channel boxes of one row); no per-row address arithmetic is left on the SM.
template <int T, bool kTmaGather>
__global__ void __launch_bounds__(128 * T + 192) gather_gemm_tc3_kernel(
    const Params p, const __grid_constant__ CUtensorMap tm64, const __grid_constant__ CUtensorMap tm32,
    const __grid_constant__ CUtensorMap ta64, const __grid_constant__ CUtensorMap ta32) {
  constexpr int kProdWarps = 4 * T;
  constexpr int kRows = kTileM * T;
  constexpr int kAStage = kABytes * T;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t s_full[8];
  __shared__ __align__(8) uint64_t s_empty[8];
  __shared__ __align__(8) uint64_t s_acc_full[2];
  __shared__ __align__(8) uint64_t s_acc_empty[2];
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.stages;
  const int n_chunks = p.n64 + p.tail32;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&s_full[s]), kTmaGather ? 1 : 128 * T + 1);   // (gather threads +) TMA expect_tx
      mbar_init(smem_u32(&s_empty[s]), 1);                     // one tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&s_acc_full[a]), 1);                  // tcgen05.commit after the last stage
      mbar_init(smem_u32(&s_acc_empty[a]), 4);                 // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kProdWarps) tmem_alloc(smem_u32(&s_tmem), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  if (warp < kProdWarps) {
    // ================================================================ A producers
    Ring ring(S);
    GatherSlots g;
    slots_init(g, p, warp, lane);
    // The walk over the active (tile, offset) steps is flattened and the map entries are loaded
    // three steps ahead of the gathers that use them, also across tile boundaries: a step is only
    // 1-3 pipeline stages, less than the latency of the dependent mask -> index loads.
    struct Step { int tile, k; KMask m; };
    auto step_next = [&](Step& it) {
      while (it.tile < p.n_tiles) {
        it.k = next_active(it.m, it.k, p.kvol);
        if (it.k < p.kvol) return;
        it.tile += gridDim.x;
        if (it.tile < p.n_tiles) it.m = load_mask_t<T>(p, it.tile);
        it.k = -1;
      }
    };
    auto step_src = [&](const Step& it) -> int32_t {
      if (it.tile >= p.n_tiles) return -1;
      if (p.dbg & 16) return (int32_t)(((int64_t)it.tile * kRows + warp * 32 + lane) % p.n_rows);
      return load_src(p, it.k, (int64_t)it.tile * kRows + warp * 32 + lane);
    };
    Step it;
    it.tile = blockIdx.x;
    it.k = -1;
    it.m = KMask{0u, 0u, 0u, 0u};
    if (it.tile < p.n_tiles) it.m = load_mask_t<T>(p, it.tile);
    step_next(it);
    bool live0 = it.tile < p.n_tiles;
    int32_t s0 = step_src(it);
    step_next(it);
    bool live1 = it.tile < p.n_tiles;
    int32_t s1 = step_src(it);
    step_next(it);
    bool live2 = it.tile < p.n_tiles;
    int32_t s2 = step_src(it);
    while (live0) {
      if constexpr (kTmaGather) {
        // lanes 0-7 fetch rows [warp*32 + 4*lane, +4) of the CTA tile; their map entries sit in lanes
        // 4*lane .. 4*lane+3 of s0
        const int q = (lane & 7) * 4;
        const int r0 = __shfl_sync(0xffffffffu, s0, q), r1 = __shfl_sync(0xffffffffu, s0, q + 1);
        const int r2 = __shfl_sync(0xffffffffu, s0, q + 2), r3 = __shfl_sync(0xffffffffu, s0, q + 3);
        const int row = warp * 32 + q;
        step_next(it);
        const bool live3 = it.tile < p.n_tiles;
        const int32_t s3 = step_src(it);
        for (int c = 0; c < n_chunks; ++c) {
          if (ring.wraps > 0) mbar_wait(smem_u32(&s_empty[ring.s]), (ring.wraps - 1) & 1);
          const bool wide = c < p.n64;
          const uint32_t dst = smem_base + ring.s * p.stage_stride + (uint32_t)(row >> 7) * kABytes +
                               (uint32_t)(row & 127) * (wide ? 128u : 64u);
          if (lane < 8 && !(p.dbg & 1))
            tma_gather4(dst, wide ? &ta64 : &ta32, wide ? c * 64 : p.n64 * 64, r0, r1, r2, r3,
                        smem_u32(&s_full[ring.s]));
          ring.advance();
        }
        s0 = s1; s1 = s2; s2 = s3;
        live0 = live1; live1 = live2; live2 = live3;
        continue;
      }
      slots_set_rows(g, p, s0, lane);
      step_next(it);
      const bool live3 = it.tile < p.n_tiles;
      const int32_t s3 = step_src(it);
      for (int c = 0; c < p.n64; ++c) {
        if (ring.wraps > 0) mbar_wait(smem_u32(&s_empty[ring.s]), (ring.wraps - 1) & 1);
        if (!(p.dbg & 1)) gather_wide(g, smem_base + ring.s * p.stage_stride, c * 64);
        cp_async_mbar_arrive_noinc(smem_u32(&s_full[ring.s]));
        ring.advance();
      }
      if (p.tail32) {
        if (ring.wraps > 0) mbar_wait(smem_u32(&s_empty[ring.s]), (ring.wraps - 1) & 1);
        if (!(p.dbg & 1)) gather_tail(g, smem_base + ring.s * p.stage_stride);
        cp_async_mbar_arrive_noinc(smem_u32(&s_full[ring.s]));
        ring.advance();
      }
      s0 = s1; s1 = s2; s2 = s3;
      live0 = live1; live1 = live2; live2 = live3;
    }
    cp_async_wait<0>();
  } else if (warp == kProdWarps) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      const int n_half = p.c_res > 256 ? p.c_res / 2 : p.c_res;
      const uint32_t idesc = make_idesc(n_half);
      Ring ring(S);
      int used = 0;                                     // non-empty tiles so far (accumulator turn)
      KMask nmask = blockIdx.x < p.n_tiles ? load_mask_t<T>(p, blockIdx.x) : KMask{0u, 0u, 0u, 0u};
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const KMask amask = nmask;                       // next tile's mask loads during this tile
        if (tile + (int)gridDim.x < p.n_tiles) nmask = load_mask_t<T>(p, tile + gridDim.x);
        if (!amask.any()) continue;
        const int ab = p.acc_bufs == 2 ? (used & 1) : 0;
        const int turn = p.acc_bufs == 2 ? (used >> 1) : used;   // uses of this accumulator before
        if (turn > 0) mbar_wait(smem_u32(&s_acc_empty[ab]), (turn - 1) & 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (uint32_t)(ab * T * p.acc_stride);
        uint32_t acc_flag = 0;
        for (int k = next_active(amask, -1, p.kvol); k < p.kvol; k = next_active(amask, k, p.kvol)) {
          for (int c = 0; c < n_chunks; ++c) {
            mbar_wait(smem_u32(&s_full[ring.s]), ring.wraps & 1);
            tc_fence_after();
            const uint32_t a_base = smem_base + ring.s * p.stage_stride;
            const uint32_t b_base = a_base + kAStage;
            if (p.dbg & 4) {
            } else if (c < p.n64) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t bd = make_desc<128>(b_base + kk * 32);
#pragma unroll
                for (int t = 0; t < T; ++t) {                  // one weight tile, T row tiles
                  const uint64_t ad = make_desc<128>(a_base + t * kABytes + kk * 32);
                  umma_f16(tmem_acc + (uint32_t)(t * p.acc_stride), ad, bd, idesc, acc_flag);
                  if (n_half != p.c_res)
                    umma_f16(tmem_acc + (uint32_t)(t * p.acc_stride + n_half), ad,
                             make_desc<128>(b_base + n_half * 128 + kk * 32), idesc, acc_flag);
                }
                acc_flag = 1;
              }
            } else {
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {
                const uint64_t bd = make_desc<64>(b_base + kk * 32);
#pragma unroll
                for (int t = 0; t < T; ++t) {
                  const uint64_t ad = make_desc<64>(a_base + t * kABytes + kk * 32);
                  umma_f16(tmem_acc + (uint32_t)(t * p.acc_stride), ad, bd, idesc, acc_flag);
                  if (n_half != p.c_res)
                    umma_f16(tmem_acc + (uint32_t)(t * p.acc_stride + n_half), ad,
                             make_desc<64>(b_base + n_half * 64 + kk * 32), idesc, acc_flag);
                }
                acc_flag = 1;
              }
            }
            umma_commit(smem_u32(&s_empty[ring.s]));     // frees the stage when these MMAs retire
            ring.advance();
          }
        }
        umma_commit(smem_u32(&s_acc_full[ab]));          // accumulator complete -> epilogue
        ++used;
      }
    }
    tc_fence_before();
  } else if (warp == kProdWarps + 1) {
    // ================================================ B producer (TMA weight tiles)
    if (lane == 0) {
      const int n_half = p.c_res > 256 ? p.c_res / 2 : p.c_res;
      Ring ring(S);
      KMask nmask = blockIdx.x < p.n_tiles ? load_mask_t<T>(p, blockIdx.x) : KMask{0u, 0u, 0u, 0u};
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const KMask amask = nmask;
        if (tile + (int)gridDim.x < p.n_tiles) nmask = load_mask_t<T>(p, tile + gridDim.x);
        for (int k = next_active(amask, -1, p.kvol); k < p.kvol; k = next_active(amask, k, p.kvol)) {
          for (int c = 0; c < n_chunks; ++c) {
            if (ring.wraps > 0) mbar_wait(smem_u32(&s_empty[ring.s]), (ring.wraps - 1) & 1);
            const uint32_t bar = smem_u32(&s_full[ring.s]);
            const uint32_t b_base = smem_base + ring.s * p.stage_stride + kAStage;
            const bool wide = c < p.n64;
            const int rowb = wide ? 128 : 64;
            uint32_t tx = (p.dbg & 2) ? 0u : (uint32_t)(p.c_res * rowb);
            if (kTmaGather && !(p.dbg & 1)) tx += (uint32_t)(kRows * rowb);   // the gathered A rows
            mbar_arrive_expect_tx(bar, tx);
            const CUtensorMap* tm = wide ? &tm64 : &tm32;
            const int col = wide ? c * 64 : p.n64 * 64;
            if (!(p.dbg & 2)) {
              tma_load_2d(b_base, tm, col, k * p.c_res, bar);
              if (n_half != p.c_res) tma_load_2d(b_base + n_half * rowb, tm, col, k * p.c_res + n_half, bar);
            }
            ring.advance();
          }
        }
      }
    }
  } else {
    // =================================================================== epilogue
    const int q = warp & 3;                              // TMEM lane quarter this warp may read
    int used = 0;
    KMask nmask = blockIdx.x < p.n_tiles ? load_mask_t<T>(p, blockIdx.x) : KMask{0u, 0u, 0u, 0u};
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const KMask amask = nmask;
      if (tile + (int)gridDim.x < p.n_tiles) nmask = load_mask_t<T>(p, tile + gridDim.x);
      const bool any = amask.any();
      const int ab = p.acc_bufs == 2 ? (used & 1) : 0;
      const int turn = p.acc_bufs == 2 ? (used >> 1) : used;
      // destination rows first: their load overlaps the wait for the accumulator
      const int64_t r_t0 = (int64_t)tile * kRows + q * 32 + lane;
      const int64_t r_t1 = r_t0 + kTileM;
      const int64_t ro0 = (p.row_perm && r_t0 < p.n_rows) ? (int64_t)__ldg(p.row_perm + r_t0) : r_t0;
      const int64_t ro1 = (T == 2 && p.row_perm && r_t1 < p.n_rows) ? (int64_t)__ldg(p.row_perm + r_t1) : r_t1;
      if (any) {
        mbar_wait_backoff(smem_u32(&s_acc_full[ab]), turn & 1);   // a whole main loop away: sleep
        tc_fence_after();
      }
#pragma unroll 1
      for (int t = 0; t < T; ++t) {
      const int64_t r = t == 0 ? r_t0 : r_t1;
      const int64_t r_out = t == 0 ? ro0 : ro1;
      const uint32_t t_lane = tmem_base + (uint32_t)((ab * T + t) * p.acc_stride) + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < p.c_res; c0 += 16) {
        uint32_t v[16];
        if (any) {
          tmem_ld16(t_lane + (uint32_t)c0, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0u;
        }
        if (r < p.n_rows && !(p.dbg & 8)) {
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
      }
      if (any) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_acc_empty[ab]));   // accumulator may be overwritten
        ++used;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kProdWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
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

// `in` as a 2-D tensor [n_src rows, c_red columns] with a one-row box of box_cols channels: the shape
// tile::gather4 wants (four such rows per instruction).
static bool make_row_map(CUtensorMap* tm, const void* in, int64_t n_src, int c_red, int box_cols) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)c_red, (cuuint64_t)n_src};
  cuuint64_t strides[1] = {(cuuint64_t)c_red * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, 1u};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(in), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int T, bool G>
static cudaError_t launch_variant(const Params& p, const CUtensorMap& tm64, const CUtensorMap& tm32,
                                  const CUtensorMap& ta64, const CUtensorMap& ta32, int grid, size_t smem,
                                  cudaStream_t st) {
  static size_t opted_in = 0;                      // per instantiation
  if (smem > opted_in) {
    cudaError_t e = cudaFuncSetAttribute(gather_gemm_tc3_kernel<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    opted_in = smem;
  }
  gather_gemm_tc3_kernel<T, G><<<grid, 128 * T + 192, smem, st>>>(p, tm64, tm32, ta64, ta32);
  return cudaSuccess;
}

}  // namespace tc3

// shared with the weight-gradient kernel (conv_tc.cu)
bool tc_make_row_map(CUtensorMap* tm, const void* base, int64_t rows, int cols, int box_cols) {
  return tc3::make_row_map(tm, base, rows, cols, box_cols);
}

// wt: [K][c_res][c_red] fp16 (K-major B operand); nbr and tile_mask must be non-null
int launch_gather_gemm_tc3(const void* in, int64_t n_src, const void* wt, int k, int c_red, int c_res, int flip_k,
                           const int32_t* nbr, const uint32_t* tile_mask, const int32_t* row_perm,
                           int64_t n_rows, const void* bias, void* out, cudaStream_t st) {
  using namespace tc3;
  Params p;
  p.in = reinterpret_cast<const __half*>(in);
  p.nbr = nbr;
  p.tile_mask = tile_mask;
  p.row_perm = row_perm;
  p.bias = reinterpret_cast<const __half*>(bias);
  p.out = reinterpret_cast<__half*>(out);
  p.n_rows = n_rows;
  p.n_tiles128 = (int)ceil_div(n_rows, kTileM);
  p.kvol = k;
  p.c_red = c_red;
  p.c_res = c_res;
  p.flip_k = flip_k;
  p.n64 = c_red / 64;
  p.tail32 = (c_red % 64) ? 1 : 0;
  const int n_half = c_res > 256 ? c_res / 2 : c_res;
  CUtensorMap tm64, tm32;
  memset(&tm64, 0, sizeof(tm64));
  memset(&tm32, 0, sizeof(tm32));
  if (p.n64) B2S_REQUIRE(make_weight_map(&tm64, wt, k, c_res, c_red, 64, n_half), B2S_ERR_CUDA,
                         "b2s_conv_gather_gemm: cuTensorMapEncodeTiled failed (64-wide)");
  if (p.tail32) B2S_REQUIRE(make_weight_map(&tm32, wt, k, c_res, c_red, 32, n_half), B2S_ERR_CUDA,
                            "b2s_conv_gather_gemm: cuTensorMapEncodeTiled failed (32-wide)");
  // accumulator layout in TMEM: power-of-two column stride per accumulator
  int stride = 32;
  while (stride < c_res) stride <<= 1;
  p.acc_stride = stride;
  // T = 2 row tiles per CTA share every weight tile.  Measured (profiles/r1_conv_microbench_v3.txt):
  // a win only for the narrowest layers (C_res <= 32: 175 -> 122 us); for C_res >= 64 two
  // independent 128-row CTAs per SM are faster than one 256-row CTA (the gather warps' instruction
  // stream per (tile, offset) step is the limiter, and the union of two tiles' masks adds steps).
  int T = stride <= 32 ? 2 : 1;
  {
    const char* et = getenv("B2S_TC_T");
    if (et && (atoi(et) == 1 || atoi(et) == 2) && stride * atoi(et) <= 512) T = atoi(et);
  }
  if (p.n_tiles128 < 2 * sm_count()) T = 1;            // small levels: keep every SM busy
  p.acc_bufs = stride * T * 2 <= 512 ? 2 : 1;
  p.tmem_cols = stride * T * p.acc_bufs;
  p.n_tiles = (p.n_tiles128 + T - 1) / T;
  const int rowb_max = p.n64 ? 128 : 64;
  p.stage_stride = (T * kABytes + ((c_res + 7) / 8) * 8 * rowb_max + 1023) & ~1023;
  // two CTAs per SM when both the ring (>= 3 stages) and the TMEM columns fit twice
  const int budget2 = 111 * 1024, budget1 = 222 * 1024;
  int ctas_per_sm = 1;
  int stages = budget2 / p.stage_stride;
  if (stages >= 3 && p.tmem_cols <= 256) {
    ctas_per_sm = 2;
  } else {
    stages = budget1 / p.stage_stride;
  }
  if (stages > 8) stages = 8;
  p.dbg = 0;
  {
    const char* ed = getenv("B2S_TC3_DBG");
    if (ed) p.dbg = atoi(ed);
  }
  {
    const char* es = getenv("B2S_TC_STAGES");
    if (es && atoi(es) >= 2 && atoi(es) <= 8 && atoi(es) * p.stage_stride <= budget1) {
      stages = atoi(es);
      ctas_per_sm = (stages * p.stage_stride <= budget2 && p.tmem_cols <= 256) ? 2 : 1;
    }
  }
  B2S_REQUIRE(stages >= 2, B2S_ERR_UNSUPPORTED, "b2s_conv_gather_gemm: tile does not fit (C=%d)", c_res);
  p.stages = stages;
  const size_t smem = (size_t)stages * p.stage_stride + 1024;
  // experimental: rows fetched by the TMA unit (tile::gather4) instead of cp.async, see the kernel
  static const bool gather4 = [] {
    const char* e = getenv("B2S_TC_GATHER4");
    return e && e[0] == '1';
  }();
  CUtensorMap ta64, ta32;
  memset(&ta64, 0, sizeof(ta64));
  memset(&ta32, 0, sizeof(ta32));
  if (gather4) {
    if (p.n64) B2S_REQUIRE(make_row_map(&ta64, in, n_src, c_red, 64), B2S_ERR_CUDA,
                           "b2s_conv_gather_gemm: cuTensorMapEncodeTiled failed (rows, 64-wide)");
    if (p.tail32) B2S_REQUIRE(make_row_map(&ta32, in, n_src, c_red, 32), B2S_ERR_CUDA,
                              "b2s_conv_gather_gemm: cuTensorMapEncodeTiled failed (rows, 32-wide)");
  }
  int grid = sm_count() * ctas_per_sm;
  if (grid > p.n_tiles) grid = p.n_tiles;
  cudaError_t e;
  if (T == 2) e = gather4 ? launch_variant<2, true>(p, tm64, tm32, ta64, ta32, grid, smem, st)
                          : launch_variant<2, false>(p, tm64, tm32, ta64, ta32, grid, smem, st);
  else e = gather4 ? launch_variant<1, true>(p, tm64, tm32, ta64, ta32, grid, smem, st)
                   : launch_variant<1, false>(p, tm64, tm32, ta64, ta32, grid, smem, st);
  B2S_REQUIRE(e == cudaSuccess, B2S_ERR_CUDA, "b2s_conv_gather_gemm: cannot opt in to %zu B smem: %s", smem,
              cudaGetErrorString(e));
  return B2S_OK;
}

}  // namespace b2s
