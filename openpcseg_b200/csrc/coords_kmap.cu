// Coordinate-level kernels: sorted-unique of hashes, strided-conv output coordinates,
// kernel-map (rule) construction.
//
// Integer / byte work, HBM- and L2-bound.  Algorithmic bytes per unit:
//   kmap build : 16 N_in (coords) + 12*slots (table) + 16 N_out + 4 K N_out (nbr_out)
//                [+ 4 K N_in (nbr_in)] + one 32 B sector per probe (K N_out probes, L2 hits)
//   pairs      : 4 K N_out read + 8 M written
// The global radix sort and the order-preserving compaction are CUB device primitives
// (toolkit headers); everything coordinate-specific is hand-written here.
#include <cub/cub.cuh>

#include "common.cuh"

namespace b2s {

// ------------------------------------------------------------------ key packing
// (batch, x, y, z) -> 64-bit key whose unsigned order is the lexicographic order of
// torch.unique(dim=0) on [b, x, y, z] rows (TS/nn/functional/downsample.py:49-51).
constexpr int kXyzBits = 18;
constexpr int kXyzBias = 1 << 17;
constexpr uint64_t kSentinelKey = ~0ULL;

__device__ __forceinline__ bool pack_ok(int x, int y, int z, int b) {
  const int lo = -kXyzBias, hi = kXyzBias - 2;
  return x >= lo && x <= hi && y >= lo && y <= hi && z >= lo && z <= hi && b >= 0 && b < 1024;
}
__device__ __forceinline__ uint64_t pack_key(int x, int y, int z, int b) {
  return ((uint64_t)(uint32_t)b << (3 * kXyzBits)) |
         ((uint64_t)(uint32_t)(x + kXyzBias) << (2 * kXyzBits)) |
         ((uint64_t)(uint32_t)(y + kXyzBias) << kXyzBits) | (uint64_t)(uint32_t)(z + kXyzBias);
}
__device__ __forceinline__ int4 unpack_key(uint64_t key) {
  const uint32_t m = (1u << kXyzBits) - 1;
  int4 c;
  c.z = (int)(key & m) - kXyzBias;
  c.y = (int)((key >> kXyzBits) & m) - kXyzBias;
  c.x = (int)((key >> (2 * kXyzBits)) & m) - kXyzBias;
  c.w = (int)(key >> (3 * kXyzBits));
  return c;
}

struct Int3 {
  int v[3];
};

// Fast path of spdownsample: snap to the coarse grid with float division + trunc
// (TS/nn/functional/downsample.py:25-28; float32 on purpose, like torch.div on int32).
__global__ void __launch_bounds__(256) snap_pack_kernel(const int4* __restrict__ coords, int64_t n,
                                                         Int3 step, uint64_t* __restrict__ keys,
                                                         int* __restrict__ range_flag) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = __ldg(coords + i);
    int x = (int)(truncf((float)c.x / (float)step.v[0]) * (float)step.v[0]);
    int y = (int)(truncf((float)c.y / (float)step.v[1]) * (float)step.v[1]);
    int z = (int)(truncf((float)c.z / (float)step.v[2]) * (float)step.v[2]);
    if (!pack_ok(x, y, z, c.w)) atomicOr(range_flag, 1);
    keys[i] = pack_key(x, y, z, c.w);
  }
}

__global__ void __launch_bounds__(256) coord_min_kernel(const int4* __restrict__ coords, int64_t n,
                                                         int* __restrict__ cmin) {
  int mx = INT_MAX, my = INT_MAX, mz = INT_MAX;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = __ldg(coords + i);
    mx = min(mx, c.x);
    my = min(my, c.y);
    mz = min(mz, c.z);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = min(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    my = min(my, __shfl_xor_sync(0xffffffffu, my, o));
    mz = min(mz, __shfl_xor_sync(0xffffffffu, mz, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(cmin + 0, mx);
    atomicMin(cmin + 1, my);
    atomicMin(cmin + 2, mz);
  }
}

// Offset `j` of get_kernel_offsets(size, stride) (TS/nn/utils/kernel.py:19-31):
// odd volume -> x fastest, even volume -> z fastest.
__host__ __device__ __forceinline__ void kernel_offset(int j, const Int3& size, const Int3& ts,
                                                        bool odd, int& ox, int& oy, int& oz) {
  int kx, ky, kz;
  if (odd) {
    kx = j % size.v[0];
    ky = (j / size.v[0]) % size.v[1];
    kz = j / (size.v[0] * size.v[1]);
  } else {
    kz = j % size.v[2];
    ky = (j / size.v[2]) % size.v[1];
    kx = j / (size.v[2] * size.v[1]);
  }
  ox = (kx - (size.v[0] + 1) / 2 + 1) * ts.v[0];
  oy = (ky - (size.v[1] + 1) / 2 + 1) * ts.v[1];
  oz = (kz - (size.v[2] + 1) / 2 + 1) * ts.v[2];
}

// Slow path of spdownsample (downsample.py:29-45): every input voxel proposes
// coord + offset[k]; candidates off the coarse grid or below the per-axis minimum
// get the sentinel key (sorted last, dropped after the unique).
__global__ void __launch_bounds__(256) expand_pack_kernel(
    const int4* __restrict__ coords, int64_t n, int kvol, Int3 size, Int3 ts, Int3 step,
    const int* __restrict__ cmin, uint64_t* __restrict__ keys, int* __restrict__ range_flag) {
  const bool odd = (kvol & 1) != 0;
  const int mnx = cmin[0], mny = cmin[1], mnz = cmin[2];
  const int64_t total = n * kvol;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / kvol;
    int j = (int)(t - i * kvol);
    int4 c = __ldg(coords + i);
    int ox, oy, oz;
    kernel_offset(j, size, ts, odd, ox, oy, oz);
    int x = c.x + ox, y = c.y + oy, z = c.z + oz;
    bool keep = (x % step.v[0] == 0) && (y % step.v[1] == 0) && (z % step.v[2] == 0) &&
                x >= mnx && y >= mny && z >= mnz;
    uint64_t key = kSentinelKey;
    if (keep) {
      if (!pack_ok(x, y, z, c.w)) atomicOr(range_flag, 1);
      key = pack_key(x, y, z, c.w);
    }
    keys[t] = key;
  }
}

__global__ void __launch_bounds__(256) unpack_kernel(const uint64_t* __restrict__ uniq,
                                                      const int64_t* __restrict__ n_uniq,
                                                      const int* __restrict__ range_flag,
                                                      int4* __restrict__ out,
                                                      int64_t* __restrict__ d_count) {
  int64_t cnt = *n_uniq;
  if (cnt > 0 && uniq[cnt - 1] == kSentinelKey) --cnt;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cnt;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = unpack_key(uniq[i]);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    d_count[0] = cnt;
    d_count[1] = *range_flag;
  }
}

// ------------------------------------------------------------------ kernel map
// One thread per output row; K probes from registers.  Row k of nbr_out is written
// coalesced; per-offset hit counts go warp-ballot -> shared -> one global atomic
// per (block, k).
__global__ void __launch_bounds__(256) kmap_probe_kernel(
    TableView table, const int4* __restrict__ out_coords, int64_t n_out, int64_t n_in,
    const int32_t* __restrict__ offsets, int kvol, int32_t* __restrict__ nbr_out,
    int32_t* __restrict__ nbr_in, int32_t* __restrict__ nbsizes, uint32_t* __restrict__ mask_out,
    uint32_t* __restrict__ mask_in) {
  const int words = (kvol + 31) >> 5;
  extern __shared__ int32_t s_mem[];
  int32_t* s_off = s_mem;             // 3*kvol
  int32_t* s_cnt = s_mem + 3 * kvol;  // kvol
  for (int t = threadIdx.x; t < 3 * kvol; t += blockDim.x) s_off[t] = offsets[t];
  for (int t = threadIdx.x; t < kvol; t += blockDim.x) s_cnt[t] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // whole warps stay in the loop so the ballots below are convergent
  const int64_t n_round = (n_out + 31) / 32 * 32;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n_round;
       o += (int64_t)gridDim.x * blockDim.x) {
    const bool live = o < n_out;
    int4 c = live ? __ldg(out_coords + o) : make_int4(0, 0, 0, 0);
    for (int k = 0; k < kvol; ++k) {
      int32_t hit = -1;
      if (live) {
        hit = table_find(table, coord_hash(c.x + s_off[3 * k], c.y + s_off[3 * k + 1],
                                           c.z + s_off[3 * k + 2], c.w));
        nbr_out[(int64_t)k * n_out + o] = hit;
        if (hit >= 0 && nbr_in) {
          nbr_in[(int64_t)k * n_in + hit] = (int32_t)o;
          if (mask_in) atomicOr(mask_in + (int64_t)(hit >> 7) * words + (k >> 5), 1u << (k & 31));
        }
      }
      unsigned m = __ballot_sync(0xffffffffu, hit >= 0);
      if (lane == 0 && m) {
        atomicAdd(s_cnt + k, __popc(m));
        // a warp's 32 rows lie inside one 128-row tile (o is warp-aligned)
        if (mask_out) atomicOr(mask_out + (o >> 7) * words + (k >> 5), 1u << (k & 31));
      }
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < kvol; t += blockDim.x)
    if (s_cnt[t]) atomicAdd(nbsizes + t, s_cnt[t]);
}

// one CTA per 128-row tile: bit k <=> some row of the tile has a neighbour for offset k
__global__ void __launch_bounds__(128) tile_mask_kernel(const int32_t* __restrict__ nbr, int kvol, int64_t n,
                                                         uint32_t* __restrict__ mask) {
  __shared__ uint32_t s_m[4];
  const int words = (kvol + 31) >> 5;
  if (threadIdx.x < 4) s_m[threadIdx.x] = 0;
  __syncthreads();
  const int64_t r = (int64_t)blockIdx.x * 128 + threadIdx.x;
  for (int k = 0; k < kvol; ++k) {
    const bool have = r < n && __ldg(nbr + (int64_t)k * n + r) >= 0;
    const unsigned b = __ballot_sync(0xffffffffu, have);
    if ((threadIdx.x & 31) == 0 && b) atomicOr(&s_m[k >> 5], 1u << (k & 31));
  }
  __syncthreads();
  if (threadIdx.x < words) mask[(int64_t)blockIdx.x * words + threadIdx.x] = s_m[threadIdx.x];
}

// Sort key that groups rows with the same neighbourhood pattern: the K presence bits of a row,
// rarest offset most significant (so rows that need a rare offset cluster into few tiles and all
// other tiles can skip it), followed by a coarse (z, x, y) code that keeps equal-pattern rows
// spatially close.  Measured on the synthetic scan (active (tile, offset) fraction): API/hash order
// 1.00, (z,x,y) order 0.61, this key 0.31 at stride 1; 0.54 -> 0.27 at stride 2.
__global__ void __launch_bounds__(256) tile_order_key_kernel(const int32_t* __restrict__ nbr, int kvol,
                                                              int64_t n, const int32_t* __restrict__ nbsizes,
                                                              const int4* __restrict__ coords, int shift,
                                                              int64_t* __restrict__ keys,
                                                              uint32_t* __restrict__ row_bits, int batch_major) {
  __shared__ int s_bit[32];
  if (threadIdx.x < kvol) {
    const int mine = nbsizes[threadIdx.x];
    int rank = 0;                                  // 0 = rarest offset
    for (int j = 0; j < kvol; ++j) {
      const int o = nbsizes[j];
      rank += (o < mine || (o == mine && j < (int)threadIdx.x)) ? 1 : 0;
    }
    s_bit[threadIdx.x] = kvol - 1 - rank;          // rarest -> most significant of the K bits
  }
  __syncthreads();
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n;
       r += (int64_t)gridDim.x * blockDim.x) {
    uint64_t m = 0;
    uint32_t bits = 0;
    for (int k = 0; k < kvol; ++k)
      if (__ldg(nbr + (int64_t)k * n + r) >= 0) {
        m |= 1ULL << s_bit[k];
        bits |= 1u << k;
      }
    if (row_bits) row_bits[r] = bits;
    const int4 c = __ldg(coords + r);
    if (batch_major) {
      // [batch 8][pattern 27][z 8][x 10][y 10]: tiles never mix scans, and the static round-robin of tiles over
      // CTAs keeps ~one scan's rows (18 MB at 96 channels) in flight - they stay in L2 across the ~4.7 offsets
      // that gather each of them; with the pattern as the leading key a tile mixed all scans of the batch and
      // every gather of a 16-scan batch (290 MB of rows) went to DRAM
      const uint64_t zc = (uint64_t)((c.z >> shift) & 0xFF), xc = (uint64_t)((c.x >> (shift + 1)) & 0x3FF),
                     yc = (uint64_t)((c.y >> (shift + 1)) & 0x3FF);
      keys[r] = (int64_t)(((uint64_t)(c.w & 0xFF) << 55) | (m << 28) | (zc << 20) | (xc << 10) | yc);
    } else {
      const uint64_t zc = (uint64_t)((c.z >> shift) & 0x3FF), xc = (uint64_t)((c.x >> shift) & 0x1FFF),
                     yc = (uint64_t)((c.y >> shift) & 0x1FFF);
      keys[r] = (int64_t)((m << 36) | (zc << 26) | (xc << 13) | yc);
    }
  }
}

// ---- step table of a gather map (consumed by conv_tc4.cu) --------------------------------------------
// pass 1: active-offset mask of every tile of `tile_rows` launch rows + its popcount
template <int TR>
__global__ void __launch_bounds__(TR) tile_steps_mask_kernel(const int32_t* __restrict__ nbr, int kvol, int64_t n,
                                                             const int32_t* __restrict__ perm,
                                                             const uint32_t* __restrict__ row_bits,
                                                             uint32_t* __restrict__ mask,
                                                             int32_t* __restrict__ counts) {
  __shared__ uint32_t s_m[4];
  const int words = (kvol + 31) >> 5;
  if (threadIdx.x < 4) s_m[threadIdx.x] = 0;
  __syncthreads();
  const int64_t j = (int64_t)blockIdx.x * TR + threadIdx.x;
  const int64_t r = j < n ? (perm ? (int64_t)__ldg(perm + j) : j) : -1;
  if (row_bits && kvol <= 32) {
    const uint32_t b = __reduce_or_sync(0xffffffffu, r >= 0 ? __ldg(row_bits + r) : 0u);
    if ((threadIdx.x & 31) == 0 && b) atomicOr(&s_m[0], b);
  } else {
    for (int k = 0; k < kvol; ++k) {
      const bool have = r >= 0 && __ldg(nbr + (int64_t)k * n + r) >= 0;
      const unsigned b = __ballot_sync(0xffffffffu, have);
      if ((threadIdx.x & 31) == 0 && b) atomicOr(&s_m[k >> 5], 1u << (k & 31));
    }
  }
  __syncthreads();
  if (threadIdx.x < words) mask[(int64_t)blockIdx.x * words + threadIdx.x] = s_m[threadIdx.x];
  if (threadIdx.x == 0) counts[blockIdx.x] = __popc(s_m[0]) + __popc(s_m[1]) + __popc(s_m[2]) + __popc(s_m[3]);
}

// pass 2: in-place exclusive scan of the per-tile counts (one CTA; a level has at most a few 10^4 tiles)
__global__ void __launch_bounds__(1024) tile_steps_scan_kernel(int32_t* __restrict__ counts, int n_tiles) {
  using Scan = cub::BlockScan<int, 1024>;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base <= n_tiles; base += 1024) {          // entry n_tiles receives the total
    const int i = base + threadIdx.x;
    const int v = i < n_tiles ? counts[i] : 0;
    int excl, total;
    Scan(tmp).ExclusiveSum(v, excl, total);
    const int carry = s_carry;
    if (i <= n_tiles) counts[i] = carry + excl;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + total;
    __syncthreads();
  }
}

// pass 3: the source rows of every active (tile, offset) step in the gather lanes' order
template <int TR>
__global__ void __launch_bounds__(TR) tile_steps_fill_kernel(const int32_t* __restrict__ nbr, int kvol, int64_t n,
                                                             const int32_t* __restrict__ perm,
                                                             const uint32_t* __restrict__ mask,
                                                             const int32_t* __restrict__ step_start,
                                                             int32_t* __restrict__ step_rows) {
  const int words = (kvol + 31) >> 5;
  const int t = threadIdx.x;
  const int64_t j = (int64_t)blockIdx.x * TR + t;
  const int64_t r = j < n ? (perm ? (int64_t)__ldg(perm + j) : j) : -1;
  // tile row w*32 + i*4 + q  ->  slot w*32 + q*8 + i
  const int pos = (t & ~31) + (t & 3) * 8 + ((t & 31) >> 2);
  int64_t s = __ldg(step_start + blockIdx.x);
  for (int w = 0; w < words; ++w) {
    uint32_t bits = __ldg(mask + (int64_t)blockIdx.x * words + w);
    while (bits) {
      const int k = w * 32 + __ffs(bits) - 1;
      bits &= bits - 1;
      step_rows[s * TR + pos] = r >= 0 ? __ldg(nbr + (int64_t)k * n + r) : -1;
      ++s;
    }
  }
}

struct PairOf {
  const int32_t* nbr;
  int32_t n_out;
  __host__ __device__ __forceinline__ int2 operator()(int f) const {
    int k = f / n_out;
    return make_int2(nbr[f], f - k * n_out);
  }
};
struct PairValid {
  __host__ __device__ __forceinline__ bool operator()(const int2& p) const { return p.x >= 0; }
};
using PairIter = cub::TransformInputIterator<int2, PairOf, cub::CountingInputIterator<int>>;

// Pairs in (chunk of launch rows, offset, launch row) order for the weight-gradient kernel: flat index
// g = (chunk * K + k) * chunk_rows + j'  ->  launch row j = chunk * chunk_rows + j', map column perm[j] (or j).
struct ChunkPairOf {
  const int32_t* nbr;
  const int32_t* perm;
  int32_t n, k, chunk_rows;
  __host__ __device__ __forceinline__ int2 operator()(int64_t g) const {
    const int64_t per_chunk = (int64_t)k * chunk_rows;
    const int64_t chunk = g / per_chunk, rem = g - chunk * per_chunk;
    const int kk = (int)(rem / chunk_rows);
    const int64_t j = chunk * chunk_rows + (rem - (int64_t)kk * chunk_rows);
    if (j >= n) return make_int2(-1, -1);
    const int32_t o = perm ? perm[j] : (int32_t)j;
    return make_int2(nbr[(int64_t)kk * n + o], o);
  }
};
using ChunkPairIter = cub::TransformInputIterator<int2, ChunkPairOf, cub::CountingInputIterator<int64_t>>;

// seg_sizes[chunk * K + k] = number of pairs of offset k whose launch row falls into the chunk
__global__ void __launch_bounds__(256) chunk_pair_count_kernel(const int32_t* __restrict__ nbr,
                                                                const int32_t* __restrict__ perm, int kvol,
                                                                int64_t n, int chunk_rows,
                                                                int32_t* __restrict__ seg_sizes) {
  const int64_t n_round = (n + 31) / 32 * 32;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_round; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = j < n ? (perm ? (int64_t)__ldg(perm + j) : j) : -1;
    const int chunk = (int)(j / chunk_rows);
    // chunk_rows is a multiple of 32, so a warp's rows share the chunk
    for (int k = 0; k < kvol; ++k) {
      const bool have = o >= 0 && __ldg(nbr + (int64_t)k * n + o) >= 0;
      const unsigned b = __ballot_sync(0xffffffffu, have);
      if ((threadIdx.x & 31) == 0 && b) atomicAdd(seg_sizes + chunk * kvol + k, __popc(b));
    }
  }
}

static size_t sort_temp_bytes(int64_t n) {
  size_t b = 0;
  cub::DeviceRadixSort::SortKeys((void*)nullptr, b, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                 (int)n);
  return b;
}
static size_t unique_temp_bytes(int64_t n) {
  size_t b = 0;
  cub::DeviceSelect::Unique((void*)nullptr, b, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                            (int64_t*)nullptr, (int)n);
  return b;
}

// Workspace layout shared by unique_i64 and downsample: [A: cap keys][B: cap keys]
// [scratch 256 B][cub temp].
struct UniqueWs {
  uint64_t* a;
  uint64_t* b;
  int64_t* n_uniq;
  int* flag;
  int* cmin;
  void* temp;
  size_t temp_bytes;
};
static size_t unique_ws_bytes(int64_t cap) {
  size_t t = sort_temp_bytes(cap);
  size_t u = unique_temp_bytes(cap);
  return align_up((size_t)cap * 8, 256) * 2 + 256 + align_up(t > u ? t : u, 256);
}
static UniqueWs carve(void* ws, int64_t cap, size_t ws_bytes) {
  UniqueWs w;
  char* p = reinterpret_cast<char*>(ws);
  size_t kb = align_up((size_t)cap * 8, 256);
  w.a = reinterpret_cast<uint64_t*>(p);
  w.b = reinterpret_cast<uint64_t*>(p + kb);
  char* s = p + 2 * kb;
  w.n_uniq = reinterpret_cast<int64_t*>(s);
  w.flag = reinterpret_cast<int*>(s + 8);
  w.cmin = reinterpret_cast<int*>(s + 16);
  w.temp = s + 256;
  w.temp_bytes = ws_bytes - (2 * kb + 256);
  return w;
}

static bool is_fast_path(const int32_t* stride, const int32_t* kernel) {
  for (int a = 0; a < 3; ++a)
    if (!(stride[a] == 1 || stride[a] == kernel[a])) return false;
  return true;
}

}  // namespace b2s

using namespace b2s;

extern "C" {

size_t b2s_unique_workspace_bytes(int64_t n) { return unique_ws_bytes(n < 1 ? 1 : n); }

int b2s_unique_i64(const int64_t* keys, int64_t n, int64_t* out, int64_t* d_count, void* ws,
                   size_t ws_bytes, b2s_stream_t stream) {
  B2S_REQUIRE(n >= 0 && d_count, B2S_ERR_INVALID, "b2s_unique_i64: bad argument");
  B2S_REQUIRE(n < (1LL << 31), B2S_ERR_UNSUPPORTED, "b2s_unique_i64: n >= 2^31");
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    cudaMemsetAsync(d_count, 0, sizeof(int64_t), st);
    return B2S_OK;
  }
  B2S_REQUIRE(keys && out && ws, B2S_ERR_INVALID, "b2s_unique_i64: null pointer");
  B2S_REQUIRE(ws_bytes >= unique_ws_bytes(n), B2S_ERR_WORKSPACE,
              "b2s_unique_i64: workspace needs %zu bytes", unique_ws_bytes(n));
  UniqueWs w = carve(ws, n, ws_bytes);
  size_t tb = w.temp_bytes;
  // signed ascending order == torch.unique order
  cub::DeviceRadixSort::SortKeys(w.temp, tb, keys, reinterpret_cast<int64_t*>(w.a), (int)n, 0, 64,
                                 st);
  tb = w.temp_bytes;
  cub::DeviceSelect::Unique(w.temp, tb, reinterpret_cast<const int64_t*>(w.a), out, d_count,
                            (int)n, st);
  B2S_CHECK_LAUNCH("b2s_unique_i64");
  return B2S_OK;
}

int64_t b2s_downsample_capacity(int64_t n, const int32_t* stride_host,
                                const int32_t* kernel_host) {
  if (is_fast_path(stride_host, kernel_host)) return n;
  return n * (int64_t)kernel_host[0] * kernel_host[1] * kernel_host[2];
}

size_t b2s_downsample_workspace_bytes(int64_t n, const int32_t* stride_host,
                                      const int32_t* kernel_host) {
  int64_t cap = b2s_downsample_capacity(n, stride_host, kernel_host);
  return unique_ws_bytes(cap < 1 ? 1 : cap);
}

int b2s_downsample_coords(const int32_t* coords, int64_t n, const int32_t* stride_host,
                          const int32_t* kernel_host, const int32_t* tensor_stride_host,
                          int32_t* out_coords, int64_t* d_count, void* ws, size_t ws_bytes,
                          b2s_stream_t stream) {
  B2S_REQUIRE(n >= 0 && stride_host && kernel_host && tensor_stride_host && d_count,
              B2S_ERR_INVALID, "b2s_downsample_coords: bad argument");
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    cudaMemsetAsync(d_count, 0, 2 * sizeof(int64_t), st);
    return B2S_OK;
  }
  for (int a = 0; a < 3; ++a)
    B2S_REQUIRE(stride_host[a] >= 1 && kernel_host[a] >= 1 && tensor_stride_host[a] >= 1,
                B2S_ERR_INVALID, "b2s_downsample_coords: non-positive stride/kernel");
  B2S_REQUIRE(coords && out_coords && ws, B2S_ERR_INVALID, "b2s_downsample_coords: null pointer");
  int64_t cap = b2s_downsample_capacity(n, stride_host, kernel_host);
  B2S_REQUIRE(cap < (1LL << 31), B2S_ERR_UNSUPPORTED, "b2s_downsample_coords: too many candidates");
  B2S_REQUIRE(ws_bytes >= unique_ws_bytes(cap), B2S_ERR_WORKSPACE,
              "b2s_downsample_coords: workspace needs %zu bytes", unique_ws_bytes(cap));
  UniqueWs w = carve(ws, cap, ws_bytes);
  Int3 step, size, ts;
  for (int a = 0; a < 3; ++a) {
    step.v[a] = stride_host[a] * tensor_stride_host[a];
    size.v[a] = kernel_host[a];
    ts.v[a] = tensor_stride_host[a];
  }
  cudaMemsetAsync(w.flag, 0, sizeof(int), st);
  const int4* c4 = reinterpret_cast<const int4*>(coords);
  if (is_fast_path(stride_host, kernel_host)) {
    snap_pack_kernel<<<grid_for(n, 256), 256, 0, st>>>(c4, n, step, w.a, w.flag);
  } else {
    cudaMemsetAsync(w.cmin, 0x7F, 3 * sizeof(int), st);
    coord_min_kernel<<<grid_for(n, 256), 256, 0, st>>>(c4, n, w.cmin);
    int kvol = size.v[0] * size.v[1] * size.v[2];
    expand_pack_kernel<<<grid_for(cap, 256), 256, 0, st>>>(c4, n, kvol, size, ts, step, w.cmin,
                                                           w.a, w.flag);
  }
  size_t tb = w.temp_bytes;
  cub::DeviceRadixSort::SortKeys(w.temp, tb, w.a, w.b, (int)cap, 0, 64, st);
  tb = w.temp_bytes;
  cub::DeviceSelect::Unique(w.temp, tb, w.b, w.a, w.n_uniq, (int)cap, st);
  unpack_kernel<<<grid_for(cap, 256), 256, 0, st>>>(w.a, w.n_uniq, w.flag,
                                                    reinterpret_cast<int4*>(out_coords), d_count);
  B2S_CHECK_LAUNCH("b2s_downsample_coords");
  return B2S_OK;
}

size_t b2s_kmap_workspace_bytes(int64_t n_in, int64_t n_out, int32_t k) {
  size_t table = align_up(b2s_table_bytes(n_in), 256);
  size_t sel = 0;
  int64_t total = (n_out < 1 ? 1 : n_out) * (int64_t)(k < 1 ? 1 : k);
  PairIter it(cub::CountingInputIterator<int>(0), PairOf{nullptr, 1});
  cub::DeviceSelect::If((void*)nullptr, sel, it, (int2*)nullptr, (int64_t*)nullptr, (int)total,
                        PairValid());
  return table > sel ? table : align_up(sel, 256);
}

int b2s_kmap_build(const int32_t* in_coords, int64_t n_in, const int32_t* out_coords,
                   int64_t n_out, const int32_t* offsets, int32_t k, int32_t* nbr_out,
                   int32_t* nbr_in, int32_t* nbsizes, uint32_t* tile_mask_out,
                   uint32_t* tile_mask_in, void* ws, size_t ws_bytes, b2s_stream_t stream) {
  B2S_REQUIRE(n_in >= 0 && n_out >= 0 && k >= 1 && k <= 2048, B2S_ERR_INVALID,
              "b2s_kmap_build: bad sizes (n_in=%lld n_out=%lld k=%d)", (long long)n_in,
              (long long)n_out, k);
  B2S_REQUIRE(nbsizes && offsets && ws, B2S_ERR_INVALID, "b2s_kmap_build: null pointer");
  B2S_REQUIRE(n_out * (int64_t)k < (1LL << 31) && n_in * (int64_t)k < (1LL << 31),
              B2S_ERR_UNSUPPORTED, "b2s_kmap_build: K*N >= 2^31");
  B2S_REQUIRE(ws_bytes >= b2s_table_bytes(n_in), B2S_ERR_WORKSPACE,
              "b2s_kmap_build: workspace needs %zu bytes", b2s_table_bytes(n_in));
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(nbsizes, 0, k * sizeof(int32_t), st);
  if (nbr_in && n_in > 0) cudaMemsetAsync(nbr_in, 0xFF, (size_t)k * n_in * sizeof(int32_t), st);
  const size_t words = (size_t)(k + 31) / 32;
  if (tile_mask_out && n_out > 0)
    cudaMemsetAsync(tile_mask_out, 0, (size_t)ceil_div(n_out, 128) * words * 4, st);
  if (tile_mask_in && n_in > 0)
    cudaMemsetAsync(tile_mask_in, 0, (size_t)ceil_div(n_in, 128) * words * 4, st);
  if (n_out == 0) return B2S_OK;
  B2S_REQUIRE(out_coords && nbr_out && (n_in == 0 || in_coords), B2S_ERR_INVALID,
              "b2s_kmap_build: null pointer");
  int rc = b2s_table_build_coords(in_coords, n_in, ws, ws_bytes, stream);
  if (rc != B2S_OK) return rc;
  TableView t = table_view(ws, n_in);
  kmap_probe_kernel<<<grid_for(n_out, 256), 256, 4 * k * sizeof(int32_t), st>>>(
      t, reinterpret_cast<const int4*>(out_coords), n_out, n_in, offsets, k, nbr_out, nbr_in,
      nbsizes, tile_mask_out, nbr_in ? tile_mask_in : nullptr);
  B2S_CHECK_LAUNCH("b2s_kmap_build");
  return B2S_OK;
}

int b2s_tile_order_key_bits(const int32_t* nbr, int32_t k, int64_t n, const int32_t* nbsizes,
                            const int32_t* coords, int32_t coord_shift, int64_t* keys, uint32_t* row_bits,
                            b2s_stream_t stream) {
  B2S_REQUIRE(k >= 1 && k <= 27 && n >= 0 && coord_shift >= 0 && coord_shift < 16, B2S_ERR_INVALID,
              "b2s_tile_order_key: bad sizes (kernel volume must be <= 27)");
  if (n == 0) return B2S_OK;
  B2S_REQUIRE(nbr && nbsizes && coords && keys, B2S_ERR_INVALID, "b2s_tile_order_key: null pointer");
  static const int batch_major = [] {
    const char* e = getenv("B2S_TILE_BATCH_MAJOR");
    return (e && e[0] == '1') ? 1 : 0;        // measured (profiles/r2_tile_order.txt): batch-major loses 7-28 %
  }();
  tile_order_key_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(
      nbr, k, n, nbsizes, reinterpret_cast<const int4*>(coords), coord_shift, keys, row_bits, batch_major);
  B2S_CHECK_LAUNCH("b2s_tile_order_key");
  return B2S_OK;
}

int b2s_tile_order_key(const int32_t* nbr, int32_t k, int64_t n, const int32_t* nbsizes,
                       const int32_t* coords, int32_t coord_shift, int64_t* keys, b2s_stream_t stream) {
  return b2s_tile_order_key_bits(nbr, k, n, nbsizes, coords, coord_shift, keys, nullptr, stream);
}

int b2s_tile_steps(const int32_t* nbr, int32_t k, int64_t n, const int32_t* perm, const uint32_t* row_bits,
                   int32_t tile_rows, uint32_t* tile_mask, int32_t* step_start, int32_t* step_rows,
                   b2s_stream_t stream) {
  B2S_REQUIRE(k >= 1 && k <= 128 && n >= 0 && (tile_rows == 128 || tile_rows == 256), B2S_ERR_INVALID,
              "b2s_tile_steps: bad sizes");
  B2S_REQUIRE(step_start, B2S_ERR_INVALID, "b2s_tile_steps: null pointer");
  cudaStream_t st = as_stream(stream);
  const int64_t tiles = ceil_div(n, (int64_t)tile_rows);
  B2S_REQUIRE(tiles * k < (1LL << 31), B2S_ERR_UNSUPPORTED, "b2s_tile_steps: too many steps");
  if (n == 0) {
    cudaMemsetAsync(step_start, 0, sizeof(int32_t), st);
    return B2S_OK;
  }
  B2S_REQUIRE(nbr && tile_mask && step_rows, B2S_ERR_INVALID, "b2s_tile_steps: null pointer");
  if (tile_rows == 128) {
    tile_steps_mask_kernel<128><<<(unsigned)tiles, 128, 0, st>>>(nbr, k, n, perm, row_bits, tile_mask, step_start);
  } else {
    tile_steps_mask_kernel<256><<<(unsigned)tiles, 256, 0, st>>>(nbr, k, n, perm, row_bits, tile_mask, step_start);
  }
  tile_steps_scan_kernel<<<1, 1024, 0, st>>>(step_start, (int)tiles);
  if (tile_rows == 128) {
    tile_steps_fill_kernel<128><<<(unsigned)tiles, 128, 0, st>>>(nbr, k, n, perm, tile_mask, step_start, step_rows);
  } else {
    tile_steps_fill_kernel<256><<<(unsigned)tiles, 256, 0, st>>>(nbr, k, n, perm, tile_mask, step_start, step_rows);
  }
  B2S_CHECK_LAUNCH("b2s_tile_steps");
  return B2S_OK;
}

int b2s_tile_mask(const int32_t* nbr, int32_t k, int64_t n, uint32_t* tile_mask, b2s_stream_t stream) {
  B2S_REQUIRE(k >= 1 && k <= 128 && n >= 0, B2S_ERR_INVALID, "b2s_tile_mask: bad sizes");
  if (n == 0) return B2S_OK;
  B2S_REQUIRE(nbr && tile_mask, B2S_ERR_INVALID, "b2s_tile_mask: null pointer");
  tile_mask_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, as_stream(stream)>>>(nbr, k, n, tile_mask);
  B2S_CHECK_LAUNCH("b2s_tile_mask");
  return B2S_OK;
}

int b2s_kmap_pairs(const int32_t* nbr_out, int32_t k, int64_t n_out, int32_t* nbmaps,
                   int64_t* d_total, void* ws, size_t ws_bytes, b2s_stream_t stream) {
  B2S_REQUIRE(k >= 1 && n_out >= 0 && d_total, B2S_ERR_INVALID, "b2s_kmap_pairs: bad argument");
  cudaStream_t st = as_stream(stream);
  if (n_out == 0) {
    cudaMemsetAsync(d_total, 0, sizeof(int64_t), st);
    return B2S_OK;
  }
  B2S_REQUIRE(nbr_out && nbmaps && ws, B2S_ERR_INVALID, "b2s_kmap_pairs: null pointer");
  int64_t total = n_out * (int64_t)k;
  B2S_REQUIRE(total < (1LL << 31), B2S_ERR_UNSUPPORTED, "b2s_kmap_pairs: K*N >= 2^31");
  PairIter it(cub::CountingInputIterator<int>(0), PairOf{nbr_out, (int32_t)n_out});
  size_t need = 0;
  cub::DeviceSelect::If((void*)nullptr, need, it, (int2*)nullptr, (int64_t*)nullptr, (int)total,
                        PairValid());
  B2S_REQUIRE(ws_bytes >= need, B2S_ERR_WORKSPACE, "b2s_kmap_pairs: workspace needs %zu bytes",
              need);
  cub::DeviceSelect::If(ws, need, it, reinterpret_cast<int2*>(nbmaps), d_total, (int)total,
                        PairValid(), st);
  B2S_CHECK_LAUNCH("b2s_kmap_pairs");
  return B2S_OK;
}

size_t b2s_kmap_pairs_chunked_workspace_bytes(int64_t n_out, int32_t k, int32_t n_chunks) {
  const int64_t chunk_rows = ((ceil_div(n_out < 1 ? 1 : n_out, (int64_t)(n_chunks < 1 ? 1 : n_chunks)) + 31) / 32) * 32;
  const int64_t total = chunk_rows * (int64_t)(n_chunks < 1 ? 1 : n_chunks) * (k < 1 ? 1 : k);
  size_t need = 0;
  ChunkPairIter it(cub::CountingInputIterator<int64_t>(0), ChunkPairOf{nullptr, nullptr, 1, 1, 32});
  cub::DeviceSelect::If((void*)nullptr, need, it, (int2*)nullptr, (int64_t*)nullptr, total, PairValid());
  return align_up(need, 256);
}

int b2s_kmap_pairs_chunked(const int32_t* nbr_out, int32_t k, int64_t n_out, const int32_t* perm,
                           int32_t n_chunks, int32_t* nbmaps, int32_t* seg_sizes, int64_t* d_total, void* ws,
                           size_t ws_bytes, b2s_stream_t stream) {
  B2S_REQUIRE(k >= 1 && n_out >= 0 && n_chunks >= 1 && d_total && seg_sizes, B2S_ERR_INVALID,
              "b2s_kmap_pairs_chunked: bad argument");
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(seg_sizes, 0, (size_t)n_chunks * k * sizeof(int32_t), st);
  if (n_out == 0) {
    cudaMemsetAsync(d_total, 0, sizeof(int64_t), st);
    return B2S_OK;
  }
  B2S_REQUIRE(nbr_out && nbmaps && ws, B2S_ERR_INVALID, "b2s_kmap_pairs_chunked: null pointer");
  B2S_REQUIRE(n_out * (int64_t)k < (1LL << 31), B2S_ERR_UNSUPPORTED, "b2s_kmap_pairs_chunked: K*N >= 2^31");
  const int64_t chunk_rows = ((ceil_div(n_out, (int64_t)n_chunks) + 31) / 32) * 32;
  const int64_t total = chunk_rows * (int64_t)n_chunks * k;
  chunk_pair_count_kernel<<<grid_for(n_out, 256), 256, 0, st>>>(nbr_out, perm, k, n_out, (int)chunk_rows, seg_sizes);
  ChunkPairIter it(cub::CountingInputIterator<int64_t>(0),
                   ChunkPairOf{nbr_out, perm, (int32_t)n_out, k, (int32_t)chunk_rows});
  size_t need = 0;
  cub::DeviceSelect::If((void*)nullptr, need, it, (int2*)nullptr, (int64_t*)nullptr, total, PairValid());
  B2S_REQUIRE(ws_bytes >= need, B2S_ERR_WORKSPACE, "b2s_kmap_pairs_chunked: workspace needs %zu bytes", need);
  cub::DeviceSelect::If(ws, need, it, reinterpret_cast<int2*>(nbmaps), d_total, total, PairValid(), st);
  B2S_CHECK_LAUNCH("b2s_kmap_pairs_chunked");
  return B2S_OK;
}

}  // extern "C"
