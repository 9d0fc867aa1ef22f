// Sparse convolution gather-GEMM, tensor-core family, revision 4: STEP-TABLE driven persistent CTAs
// (fp16, sm_100a).  Replaces the forward / input-gradient loops of the reference
// (TS/backend/convolution/convolution_cuda.cu:53-278: per offset gather -> cuBLAS -> scatter).
//
// Why (profiles/r1_conv_ablation.txt, r2_*): revision 3 spent 70 % of its time in the instruction
// streams of its roles with every copy, MMA and store switched off - per (tile, offset) step each of the
// four gather warps walked the tile's offset mask, loaded the map entries of its 32 rows, spread them
// with 12 shuffles, and all four roles repeated the mask walk.  Revision 4 moves that work out of the
// kernel: b2s_tile_steps (coords_kmap.cu) compacts, once per kernel map, the ACTIVE (tile, offset) steps
// into a table
//     step_start[tile] .. step_start[tile + 1]      step ids of a tile, offsets ascending
//     step_rows[step][TR]                            source row (or -1) of every tile row, stored in the
//                                                    order the gather lanes consume it
// that all ~12-24 convolutions of a level reuse.  A gather warp's step is then: two 16-byte loads
// (prefetched three steps ahead), one multiply + one cp.async per 16 bytes.  MMA issuer and epilogue
// need only the step COUNT of a tile, the weight-tile producer walks the one-word mask.
//
// Also new: 96-channel reductions (64 + 32) are ONE pipeline stage (six MMAs behind one barrier pair),
// the epilogue can accumulate per-channel sum / sum of squares of the rows it writes (batch-norm
// statistics without another pass over [N, C]), and its wait backs off exponentially.
//
// Warp roles (6 + 4T warps): 0..4T-1 gather producers (cp.async 16 B, zero-fill, SW128 / SW64 tiles),
// 4T MMA issuer (+ TMEM alloc), 4T+1 TMA weight-tile producer, 4T+2.. epilogue (TMEM -> fp16 rows).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace b2s {
namespace tc4 {
using namespace tc;

constexpr int kABytes = kTileM * 128;      // 128 rows x 64 channels
constexpr int kATail = kTileM * 64;        // 128 rows x 32 channels

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
// wait of a role that is a whole tile away from its event: poll, then sleep with doubling back-off
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, uint32_t cap_ns) {
  uint32_t ns = 32;
  while (true) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(kWaitHintNs)
        : "memory");
    if (done) break;
    if (cap_ns) {
      __nanosleep(ns);
      if (ns < cap_ns) ns <<= 1;
    }
  }
}

// critical-path wait: plain try_wait spin when hint_ns == 0, else try_wait with a suspend-time hint
__device__ __forceinline__ void mbar_wait_h(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  if (hint_ns == 0) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP_S:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE_S;\n\t"
        "bra WAIT_LOOP_S;\n\t"
        "WAIT_DONE_S:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP_H:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra WAIT_DONE_H;\n\t"
        "bra WAIT_LOOP_H;\n\t"
        "WAIT_DONE_H:\n\t"
        "}" ::"r"(bar),
        "r"(parity), "r"(hint_ns)
        : "memory");
  }
}

struct Params {
  const __half* in;            // [n_src, c_red]
  const int32_t* step_rows;    // [steps][128 T] source rows in lane order (see b2s_tile_steps)
  const int32_t* step_start;   // [n_tiles + 1]
  const uint32_t* tile_mask;   // [n_tiles][words] active offsets of a tile (stored offset order)
  const int32_t* row_perm;     // out row of launch row j, or nullptr
  const __half* bias;          // [c_res] or nullptr
  __half* out;                 // [n_rows, c_res]
  double* bn_sums;             // [2][c_res] += (sum, sum of squares) of the rows written, or nullptr
  int64_t n_rows;
  int n_tiles;                 // CTA tiles (128 T rows each)
  int kvol, words, c_red, c_res, flip_k;
  int n64, tail32, merge_tail; // c_red = 64 n64 + 32 tail32; merge: tail shares the last wide stage
  int stages, stage_stride;
  int acc_stride, acc_bufs, tmem_cols;
  int dbg;                     // ablation (B2S_TC4_DBG): 1 no gathers, 2 no weight TMA, 4 no MMAs, 8 no stores, 16 no epilogue body
  int wait_ns;                 // suspend-time hint of the pipeline waits (B2S_TC4_WAIT_NS; 0 = spin)
  int sleep_ns;                // cap of the epilogue's back-off sleep (B2S_TC4_SLEEP_NS; 0 = no sleep)
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

// kMinBlocks: CTAs per SM the register allocation must allow (3 caps the kernel at 68 registers per thread)
template <int T, int kMinBlocks>
__global__ void __launch_bounds__(128 * T + 192, kMinBlocks) gather_gemm_tc4_kernel(
    const Params p, const __grid_constant__ CUtensorMap tm64, const __grid_constant__ CUtensorMap tm32) {
  constexpr int kProdWarps = 4 * T;
  constexpr int kRows = kTileM * T;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t s_full[8];
  __shared__ __align__(8) uint64_t s_empty[8];
  __shared__ __align__(8) uint64_t s_acc_full[2];
  __shared__ __align__(8) uint64_t s_acc_empty[2];
  __shared__ uint32_t s_tmem;
  __shared__ float s_stat[2 * 512];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.stages;
  const int n_chunks = p.merge_tail ? p.n64 : p.n64 + p.tail32;
  // smem layout of a stage: [A wide: T x 16 KiB][A tail: T x 8 KiB (merged stages only)][B wide][B tail]
  const uint32_t a_tail_off = p.merge_tail ? (uint32_t)(T * kABytes) : 0u;
  const uint32_t b_off = (uint32_t)(T * kABytes) + (p.merge_tail ? (uint32_t)(T * kATail) : 0u);
  const uint32_t b_tail_off = p.merge_tail ? b_off + (uint32_t)(p.c_res * 128) : b_off;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&s_full[s]), 128 * T + 1);            // gather threads + the TMA expect_tx
      mbar_init(smem_u32(&s_empty[s]), 1);                     // one tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&s_acc_full[a]), 1);                  // tcgen05.commit after the last stage
      mbar_init(smem_u32(&s_acc_empty[a]), 4);                 // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (p.bn_sums)
    for (int i = tid; i < 2 * p.c_res; i += blockDim.x) s_stat[i] = 0.f;
  if (warp == kProdWarps) tmem_alloc(smem_u32(&s_tmem), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  if (warp < kProdWarps) {
    // ================================================================ A producers
    // lane (q = lane >> 3, j = lane & 7) copies the 16-byte column j of tile rows warp*32 + i*4 + q, i = 0..7
    // (wide chunk) and column j & 3 of rows i = 4*(j >> 2) .. +3 (32-channel tail); the table stores the eight
    // source rows of (warp, q) contiguously, so a lane's bookkeeping per step is two 16-byte loads.
    Ring ring(S);
    const int q = lane >> 3, j8 = lane & 7;
    const char* base128 = reinterpret_cast<const char*>(p.in) + j8 * 16;
    const char* base64 = reinterpret_cast<const char*>(p.in) + p.n64 * 128 + (j8 & 3) * 16;
    const uint32_t row_bytes = (uint32_t)p.c_red * 2u;
    uint32_t off128[8], off64[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = warp * 32 + i * 4 + q;
      off128[i] = (uint32_t)(row >> 7) * kABytes + swz<128>(row & 127, j8);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int row = warp * 32 + ((j8 >> 2) * 4 + t) * 4 + q;
      off64[t] = (uint32_t)(row >> 7) * kATail + swz<64>(row & 127, j8 & 3);
    }
    const int4* table = reinterpret_cast<const int4*>(p.step_rows) + (warp * 32 + q * 8) / 4;
    constexpr int kQuads = kRows / 4;                          // int4 per step

    // flattened walk over this CTA's steps; `f` runs three steps ahead of the gathers
    struct Cursor { int tile, s, s_end; };
    auto open_tile = [&](Cursor& c) {
      while (c.tile < p.n_tiles) {
        c.s = __ldg(p.step_start + c.tile);
        c.s_end = __ldg(p.step_start + c.tile + 1);
        if (c.s < c.s_end) return;
        c.tile += gridDim.x;
      }
    };
    auto advance = [&](Cursor& c) {
      if (++c.s == c.s_end) {
        c.tile += gridDim.x;
        open_tile(c);
      }
    };
    auto fetch = [&](const Cursor& c, int4& lo, int4& hi) {
      if (c.tile < p.n_tiles) {
        const int4* src = table + (int64_t)c.s * kQuads;
        lo = __ldg(src);
        hi = __ldg(src + 1);
      }
    };
    Cursor f{(int)blockIdx.x, 0, 0};
    open_tile(f);
    int4 lo0, hi0, lo1, hi1, lo2, hi2;
    lo0 = hi0 = lo1 = hi1 = lo2 = hi2 = make_int4(-1, -1, -1, -1);
    bool live0 = f.tile < p.n_tiles;
    fetch(f, lo0, hi0);
    if (live0) advance(f);
    bool live1 = live0 && f.tile < p.n_tiles;
    fetch(f, lo1, hi1);
    if (live1) advance(f);
    bool live2 = live1 && f.tile < p.n_tiles;
    fetch(f, lo2, hi2);
    if (live2) advance(f);
    while (live0) {
      int4 lo3 = make_int4(-1, -1, -1, -1), hi3 = lo3;
      const bool live3 = live2 && f.tile < p.n_tiles;
      fetch(f, lo3, hi3);
      if (live3) advance(f);
      const int r[8] = {lo0.x, lo0.y, lo0.z, lo0.w, hi0.x, hi0.y, hi0.z, hi0.w};
      const int4 tl = (j8 >> 2) ? hi0 : lo0;
      const int rt[4] = {tl.x, tl.y, tl.z, tl.w};
      for (int c = 0; c < n_chunks; ++c) {
        if (ring.wraps > 0) mbar_wait_h(smem_u32(&s_empty[ring.s]), (ring.wraps - 1) & 1, p.wait_ns);
        const uint32_t a_base = smem_base + ring.s * p.stage_stride;
        if (!(p.dbg & 1)) {
          if (c < p.n64) {
            const char* b = base128 + c * 128;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const bool ok = r[i] >= 0;
              cp_async16(a_base + off128[i], b + (ok ? (uint32_t)r[i] * row_bytes : 0u), ok ? 16u : 0u);
            }
          }
          if (p.tail32 && (p.merge_tail ? c == p.n64 - 1 : c == p.n64)) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const bool ok = rt[t] >= 0;
              cp_async16(a_base + a_tail_off + off64[t], base64 + (ok ? (uint32_t)rt[t] * row_bytes : 0u),
                         ok ? 16u : 0u);
            }
          }
        }
        cp_async_mbar_arrive_noinc(smem_u32(&s_full[ring.s]));
        ring.advance();
      }
      lo0 = lo1; hi0 = hi1; lo1 = lo2; hi1 = hi2; lo2 = lo3; hi2 = hi3;
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
      int s_next = blockIdx.x < p.n_tiles ? __ldg(p.step_start + blockIdx.x) : 0;
      int e_next = blockIdx.x < p.n_tiles ? __ldg(p.step_start + blockIdx.x + 1) : 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int n_steps = e_next - s_next;
        if (tile + (int)gridDim.x < p.n_tiles) {        // next tile's bounds load during this tile
          s_next = __ldg(p.step_start + tile + gridDim.x);
          e_next = __ldg(p.step_start + tile + gridDim.x + 1);
        }
        if (n_steps == 0) continue;
        const int ab = p.acc_bufs == 2 ? (used & 1) : 0;
        const int turn = p.acc_bufs == 2 ? (used >> 1) : used;   // uses of this accumulator before
        if (turn > 0) mbar_wait_h(smem_u32(&s_acc_empty[ab]), (turn - 1) & 1, p.wait_ns);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (uint32_t)(ab * T * p.acc_stride);
        uint32_t acc_flag = 0;
        for (int st = 0; st < n_steps; ++st) {
          for (int c = 0; c < n_chunks; ++c) {
            mbar_wait_h(smem_u32(&s_full[ring.s]), ring.wraps & 1, p.wait_ns);
            // rows were written through the generic proxy (cp.async): make them visible to the async proxy
            // the MMAs read through (PTX memory model; measured cost: none, see profiles/r2_conv_tc4.txt)
            fence_proxy_async();
            tc_fence_after();
            const uint32_t a_base = smem_base + ring.s * p.stage_stride;
            if (!(p.dbg & 4)) {
              if (c < p.n64) {
                const uint32_t b_base = a_base + b_off;
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
              }
              if (p.tail32 && (p.merge_tail ? c == p.n64 - 1 : c == p.n64)) {
                const uint32_t at_base = a_base + a_tail_off, bt_base = a_base + b_tail_off;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                  const uint64_t bd = make_desc<64>(bt_base + kk * 32);
#pragma unroll
                  for (int t = 0; t < T; ++t) {
                    const uint64_t ad = make_desc<64>(at_base + t * kATail + kk * 32);
                    umma_f16(tmem_acc + (uint32_t)(t * p.acc_stride), ad, bd, idesc, acc_flag);
                    if (n_half != p.c_res)
                      umma_f16(tmem_acc + (uint32_t)(t * p.acc_stride + n_half), ad,
                               make_desc<64>(bt_base + n_half * 64 + kk * 32), idesc, acc_flag);
                  }
                  acc_flag = 1;
                }
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
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const uint32_t* tm = p.tile_mask + (int64_t)tile * p.words;
        for (int w = 0; w < p.words; ++w) {
          uint32_t bits = __ldg(tm + w);
          while (bits) {
            const int ks = w * 32 + __ffs(bits) - 1;     // stored offset index, ascending = table order
            bits &= bits - 1;
            const int k = p.flip_k ? p.kvol - 1 - ks : ks;
            for (int c = 0; c < n_chunks; ++c) {
              if (ring.wraps > 0) mbar_wait_h(smem_u32(&s_empty[ring.s]), (ring.wraps - 1) & 1, p.wait_ns);
              const uint32_t bar = smem_u32(&s_full[ring.s]);
              const uint32_t a_base = smem_base + ring.s * p.stage_stride;
              const bool wide = c < p.n64;
              const bool tail = p.tail32 && (p.merge_tail ? c == p.n64 - 1 : c == p.n64);
              uint32_t tx = 0;
              if (!(p.dbg & 2)) tx = (wide ? (uint32_t)(p.c_res * 128) : 0u) + (tail ? (uint32_t)(p.c_res * 64) : 0u);
              mbar_arrive_expect_tx(bar, tx);
              if (!(p.dbg & 2)) {
                if (wide) {
                  tma_load_2d(a_base + b_off, &tm64, c * 64, k * p.c_res, bar);
                  if (n_half != p.c_res)
                    tma_load_2d(a_base + b_off + n_half * 128, &tm64, c * 64, k * p.c_res + n_half, bar);
                }
                if (tail) {
                  tma_load_2d(a_base + b_tail_off, &tm32, p.n64 * 64, k * p.c_res, bar);
                  if (n_half != p.c_res)
                    tma_load_2d(a_base + b_tail_off + n_half * 64, &tm32, p.n64 * 64, k * p.c_res + n_half, bar);
                }
              }
              ring.advance();
            }
          }
        }
      }
    }
  } else {
    // =================================================================== epilogue
    const int q = warp & 3;                              // TMEM lane quarter this warp may read
    int used = 0;
    int s_next = blockIdx.x < p.n_tiles ? __ldg(p.step_start + blockIdx.x) : 0;
    int e_next = blockIdx.x < p.n_tiles ? __ldg(p.step_start + blockIdx.x + 1) : 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const bool any = e_next > s_next;
      if (tile + (int)gridDim.x < p.n_tiles) {
        s_next = __ldg(p.step_start + tile + gridDim.x);
        e_next = __ldg(p.step_start + tile + gridDim.x + 1);
      }
      const int ab = p.acc_bufs == 2 ? (used & 1) : 0;
      const int turn = p.acc_bufs == 2 ? (used >> 1) : used;
      // destination rows first: their load overlaps the wait for the accumulator
      const int64_t r_t0 = (int64_t)tile * kRows + q * 32 + lane;
      const int64_t r_t1 = r_t0 + kTileM;
      const int64_t ro0 = (p.row_perm && r_t0 < p.n_rows) ? (int64_t)__ldg(p.row_perm + r_t0) : r_t0;
      const int64_t ro1 = (T == 2 && p.row_perm && r_t1 < p.n_rows) ? (int64_t)__ldg(p.row_perm + r_t1) : r_t1;
      if (any) {
        mbar_wait_sleep(smem_u32(&s_acc_full[ab]), turn & 1, (uint32_t)p.sleep_ns);   // a whole main loop away
        tc_fence_after();
      }
#pragma unroll 1
      for (int t = 0; t < T && !(p.dbg & 16); ++t) {
        const int64_t r = t == 0 ? r_t0 : r_t1;
        const int64_t r_out = t == 0 ? ro0 : ro1;
        const bool live = r < p.n_rows;
        const uint32_t t_lane = tmem_base + (uint32_t)((ab * T + t) * p.acc_stride) + ((uint32_t)(q * 32) << 16);
        // 16 columns per TMEM load (the 32-column form was measured: more registers, spills, +6 % kernel time);
        // fp32 -> fp16 two at a time; the epilogue warps share issue slots with the gather warps (ablation: no
        // epilogue body = -19 % kernel time), so every instruction here counts
        for (int c0 = 0; c0 < p.c_res;) {
          constexpr int w = 16;
          uint32_t v[16];
          if (any) {
            tmem_ld16(t_lane + (uint32_t)c0, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0u;
          }
          uint32_t hp[8];                                    // packed half2
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float f0 = __uint_as_float(v[2 * j]), f1 = __uint_as_float(v[2 * j + 1]);
            if (p.bias && 2 * j < w) {
              f0 += __half2float(__ldg(p.bias + c0 + 2 * j));
              f1 += __half2float(__ldg(p.bias + c0 + 2 * j + 1));
            }
            const __half2 h2 = __floats2half2_rn(f0, f1);
            hp[j] = *reinterpret_cast<const uint32_t*>(&h2);
          }
          if (live && !(p.dbg & 8)) {
            uint4* dst = reinterpret_cast<uint4*>(p.out + r_out * p.c_res + c0);
            dst[0] = make_uint4(hp[0], hp[1], hp[2], hp[3]);
            dst[1] = make_uint4(hp[4], hp[5], hp[6], hp[7]);
          }
          if (p.bn_sums) {
            // per-column sum / sum of squares of the fp16 values just written, over the warp's 32 rows:
            // halving exchange (16 + 8 + 4 + 2 + 1 column slots survive per lane), then one shared-memory
            // atomic per column from the lanes that end up owning it
            {
              constexpr int h0 = 0;
              float a[16], b[16];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint32_t u = hp[j];
                const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&u));
                const float x0 = live ? f2.x : 0.f, x1 = live ? f2.y : 0.f;
                a[2 * j] = x0;
                a[2 * j + 1] = x1;
                b[2 * j] = x0 * x0;
                b[2 * j + 1] = x1 * x1;
              }
#pragma unroll
              for (int width = 8, bit = 16; width >= 1; width >>= 1, bit >>= 1) {
                const bool upper = (lane & bit) != 0;
#pragma unroll
                for (int j = 0; j < width; ++j) {
                  const float sa = upper ? a[j] : a[j + width], sb = upper ? b[j] : b[j + width];
                  const float ka = upper ? a[j + width] : a[j], kb = upper ? b[j + width] : b[j];
                  a[j] = ka + __shfl_xor_sync(0xffffffffu, sa, bit);
                  b[j] = kb + __shfl_xor_sync(0xffffffffu, sb, bit);
                }
              }
              a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
              b[0] += __shfl_xor_sync(0xffffffffu, b[0], 1);
              if ((lane & 1) == 0) {
                const int col = c0 + h0 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 +
                                ((lane >> 1) & 1);
                atomicAdd(&s_stat[col], a[0]);
                atomicAdd(&s_stat[p.c_res + col], b[0]);
              }
            }
          }
          c0 += w;
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
  if (p.bn_sums)
    for (int i = tid; i < 2 * p.c_res; i += blockDim.x)
      if (s_stat[i] != 0.f) atomicAdd(p.bn_sums + i, (double)s_stat[i]);
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

// a [rows, cols] fp16 tensor with a one-row box of box_cols channels: the shape tile::gather4 wants (four such
// rows per instruction); used by the weight-gradient kernel's TMA variant
static bool make_row_map(CUtensorMap* tm, const void* base, int64_t rows, int cols, int box_cols) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, 1u};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int T, int kMinBlocks>
static cudaError_t launch_variant(const Params& p, const CUtensorMap& tm64, const CUtensorMap& tm32, int grid,
                                  size_t smem, cudaStream_t st) {
  static size_t opted_in = 0;                      // per instantiation
  if (smem > opted_in) {
    cudaError_t e = cudaFuncSetAttribute(gather_gemm_tc4_kernel<T, kMinBlocks>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    opted_in = smem;
  }
  gather_gemm_tc4_kernel<T, kMinBlocks><<<grid, 128 * T + 192, smem, st>>>(p, tm64, tm32);
  return cudaSuccess;
}

}  // namespace tc4

bool tc_make_row_map(CUtensorMap* tm, const void* base, int64_t rows, int cols, int box_cols) {
  return tc4::make_row_map(tm, base, rows, cols, box_cols);
}

// Row tiles per CTA tile the kernel will use for this shape; the caller builds the step table with
// tile_rows = 128 * T (b2s_conv_tile_rows).
int tc4_tile_rows(int c_res, int64_t n_rows) {
  int stride = 32;
  while (stride < c_res) stride <<= 1;
  // T = 2 (256-row CTA tiles sharing each weight tile) was a win for C_res <= 32 in revision 3 (175 -> 122 us);
  // with step tables the 128-row form wins everywhere (batch 16, 32 -> 32: 289 -> 216 us, profiles/r2_tile_order.txt)
  // - kept behind B2S_TC_T=2 for experiments
  int T = 1;
  {
    const char* et = getenv("B2S_TC_T");
    if (et && (atoi(et) == 1 || atoi(et) == 2) && stride * atoi(et) <= 512) T = atoi(et);
  }
  if ((n_rows + 127) / 128 < 2 * sm_count()) T = 1;            // small levels: keep every SM busy
  return 128 * T;
}

// wt: [K][c_res][c_red] fp16 (K-major B operand); the step table was built with `tile_rows` rows per tile
int launch_gather_gemm_tc4(const void* in, int64_t n_src, const void* wt, int k, int c_red, int c_res, int flip_k,
                           const int32_t* step_rows, const int32_t* step_start, const uint32_t* tile_mask,
                           int tile_rows, const int32_t* row_perm, int64_t n_rows, const void* bias, void* out,
                           double* bn_sums, cudaStream_t st) {
  using namespace tc4;
  (void)n_src;
  B2S_REQUIRE(tile_rows == 128 || tile_rows == 256, B2S_ERR_INVALID, "b2s_conv_gather_gemm: tile_rows %d", tile_rows);
  const int T = tile_rows / 128;
  Params p;
  p.in = reinterpret_cast<const __half*>(in);
  p.step_rows = step_rows;
  p.step_start = step_start;
  p.tile_mask = tile_mask;
  p.row_perm = row_perm;
  p.bias = reinterpret_cast<const __half*>(bias);
  p.out = reinterpret_cast<__half*>(out);
  p.bn_sums = bn_sums;
  p.n_rows = n_rows;
  p.n_tiles = (int)ceil_div(n_rows, (int64_t)tile_rows);
  p.kvol = k;
  p.words = (k + 31) / 32;
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
  int stride = 32;
  while (stride < c_res) stride <<= 1;
  p.acc_stride = stride;
  B2S_REQUIRE(stride * T <= 512, B2S_ERR_UNSUPPORTED, "b2s_conv_gather_gemm: C_res %d with %d-row tiles", c_res,
              tile_rows);
  const int crp = ((c_res + 7) / 8) * 8;
  const int rowb_max = p.n64 ? 128 : 64;
  const int plain_stride = (T * kABytes + crp * rowb_max + 1023) & ~1023;
  const int merged_stride = (T * (kABytes + kATail) + crp * 192 + 1023) & ~1023;
  bool may_merge = p.n64 == 1 && p.tail32;
  {
    const char* em = getenv("B2S_TC4_MERGE");
    if (em && em[0] == '0') may_merge = false;
  }
  // CTAs per SM.  Every CTA is one independent gather -> MMA -> epilogue chain, and the number of chains per SM,
  // not the ring depth, is what hides the hand-shake and gather latencies (weight gradient: 1 x 4 stages ->
  // 2 x 2 stages = -27 %, profiles/r2_wgrad_ctas.txt), so take the most CTAs that fit with >= 2 stages each:
  // shared memory 228 KiB per SM, per CTA 1 KiB reserved + ~5 KiB static + 1 KiB alignment slack; TMEM 512
  // columns per SM (a CTA needs T accumulators, double-buffered when they fit); three CTAs also need the
  // 68-register variant.  96-channel reductions merge the 32-channel tail into the 64-channel stage when two
  // such stages fit.
  static const int max_ctas = [] {
    const char* e = getenv("B2S_TC4_CTAS");
    int v = e ? atoi(e) : 2;
    return v < 1 ? 1 : (v > 3 ? 3 : v);
  }();
  int ctas_per_sm = 1, stages = 0;
  p.merge_tail = 0;
  p.acc_bufs = 1;
  for (int c = (T == 1 ? max_ctas : (max_ctas > 2 ? 2 : max_ctas)); c >= 1; --c) {
    const int budget = (228 * 1024) / c - 7 * 1024;
    if (c * stride * T > 512 && c > 1) continue;
    const bool merge = may_merge && 2 * merged_stride <= budget;
    const int ss = merge ? merged_stride : plain_stride;
    const int st_fit = budget / ss;
    if (st_fit < 2 && c > 1) continue;
    ctas_per_sm = c;
    stages = st_fit;
    p.merge_tail = merge ? 1 : 0;
    p.stage_stride = ss;
    p.acc_bufs = (c * stride * T * 2 <= 512) ? 2 : 1;
    break;
  }
  p.tmem_cols = stride * T * p.acc_bufs;
  if (stages > 8) stages = 8;
  p.dbg = 0;
  p.wait_ns = 0;
  {
    const char* ed = getenv("B2S_TC4_DBG");
    if (ed) p.dbg = atoi(ed);
    const char* ew = getenv("B2S_TC4_WAIT_NS");
    if (ew) p.wait_ns = atoi(ew);
    const char* es = getenv("B2S_TC4_SLEEP_NS");
    p.sleep_ns = es ? atoi(es) : 256;
  }
  {
    const char* es = getenv("B2S_TC_STAGES");
    if (es && atoi(es) >= 2 && atoi(es) <= stages) stages = atoi(es);
  }
  B2S_REQUIRE(stages >= 2, B2S_ERR_UNSUPPORTED, "b2s_conv_gather_gemm: tile does not fit (C=%d)", c_res);
  p.stages = stages;
  const size_t smem = (size_t)stages * p.stage_stride + 1024;
  int grid = persistent_sms() * ctas_per_sm;
  if (grid > p.n_tiles) grid = p.n_tiles;
  cudaError_t e = T == 2 ? launch_variant<2, 1>(p, tm64, tm32, grid, smem, st)
                  : ctas_per_sm == 3 ? launch_variant<1, 3>(p, tm64, tm32, grid, smem, st)
                                     : launch_variant<1, 2>(p, tm64, tm32, grid, smem, st);
  B2S_REQUIRE(e == cudaSuccess, B2S_ERR_CUDA, "b2s_conv_gather_gemm: cannot opt in to %zu B smem: %s", smem,
              cudaGetErrorString(e));
  return B2S_OK;
}

}  // namespace b2s
