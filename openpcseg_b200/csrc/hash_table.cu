// sphash / kernel-hash / hash table / count kernels.
//
// Roofline: all HBM-streaming.  Algorithmic bytes: hash 16N + 8N; kernel-hash
// 16N + 8KN; table build 8N read + 12N slot writes; query 8Q + 8Q + one 32 B
// sector per probe.  One int4 (16 B) coordinate load per thread, one 8 B store.
#include <stdarg.h>

#include <stdlib.h>

#include "common.cuh"

namespace b2s {

static thread_local char g_err[512] = "ok";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

// SMs the persistent kernels (conv gather-GEMM, wgrad) may fill: all but `reserve`, so that a concurrent
// collective (NCCL's gradient all-reduce under DDP) finds free SMs instead of queueing behind a resident grid
static int g_sm_reserve = -1;
int persistent_sms() {
  if (g_sm_reserve < 0) {
    const char* e = getenv("B2S_SM_RESERVE");
    g_sm_reserve = e ? atoi(e) : 0;
    if (g_sm_reserve < 0) g_sm_reserve = 0;
  }
  const int n = sm_count() - g_sm_reserve;
  return n < 8 ? 8 : n;
}
void set_sm_reserve(int n) { g_sm_reserve = n < 0 ? 0 : n; }

__global__ void __launch_bounds__(256) hash_kernel(const int4* __restrict__ coords, int64_t n,
                                                    int64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = __ldg(coords + i);
    out[i] = coord_hash(c.x, c.y, c.z, c.w);
  }
}

// One thread per point, K hashes from registers; row k of the [K, N] output is
// written coalesced.
__global__ void __launch_bounds__(256) kernel_hash_kernel(const int4* __restrict__ coords,
                                                           int64_t n,
                                                           const int32_t* __restrict__ offsets,
                                                           int k, int64_t* __restrict__ out) {
  extern __shared__ int32_t s_off[];
  for (int t = threadIdx.x; t < 3 * k; t += blockDim.x) s_off[t] = offsets[t];
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = __ldg(coords + i);
    for (int j = 0; j < k; ++j)
      out[(int64_t)j * n + i] =
          coord_hash(c.x + s_off[3 * j], c.y + s_off[3 * j + 1], c.z + s_off[3 * j + 2], c.w);
  }
}

__global__ void __launch_bounds__(256) table_build_keys_kernel(const int64_t* __restrict__ keys,
                                                                int64_t n, TableView t) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    table_insert(t, __ldg(keys + i), (int32_t)i);
}

__global__ void __launch_bounds__(256) table_build_coords_kernel(const int4* __restrict__ coords,
                                                                  int64_t n, TableView t) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = __ldg(coords + i);
    table_insert(t, coord_hash(c.x, c.y, c.z, c.w), (int32_t)i);
  }
}

__global__ void __launch_bounds__(256) table_query_kernel(TableView t,
                                                           const int64_t* __restrict__ q,
                                                           int64_t nq, int64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nq;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (int64_t)table_find(t, __ldg(q + i));
}

__global__ void __launch_bounds__(256) count_kernel(const int32_t* __restrict__ idx, int64_t n,
                                                     int32_t* __restrict__ out, int64_t num) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int32_t v = __ldg(idx + i);
    if (v >= 0 && v < num) atomicAdd(out + v, 1);
  }
}

int grid_for(int64_t n, int threads) {
  int64_t blocks = ceil_div(n, threads);
  int64_t cap = (int64_t)sm_count() * 16;  // multiple of the SM count; grid-stride above it
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

int table_clear(const TableView& t, cudaStream_t st) {
  int64_t slots = (int64_t)t.mask + 1;
  cudaMemsetAsync(t.keys, 0xFF, slots * sizeof(int64_t), st);  // kEmptyKey
  cudaMemsetAsync(t.vals, 0x7F, slots * sizeof(int32_t), st);  // +inf for atomicMin
  return 0;
}

}  // namespace b2s

using namespace b2s;

extern "C" {

const char* b2s_last_error(void) { return g_err; }
int b2s_version(void) { return 200; }
void b2s_set_sm_reserve(int32_t n) { b2s::set_sm_reserve(n); }

int b2s_hash(const int32_t* coords, int64_t n, int64_t* out, b2s_stream_t stream) {
  B2S_REQUIRE(n >= 0, B2S_ERR_INVALID, "b2s_hash: n < 0");
  if (n == 0) return B2S_OK;
  B2S_REQUIRE(coords && out, B2S_ERR_INVALID, "b2s_hash: null pointer");
  hash_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const int4*>(coords), n, out);
  B2S_CHECK_LAUNCH("b2s_hash");
  return B2S_OK;
}

int b2s_kernel_hash(const int32_t* coords, int64_t n, const int32_t* offsets, int32_t k,
                    int64_t* out, b2s_stream_t stream) {
  B2S_REQUIRE(n >= 0 && k >= 0, B2S_ERR_INVALID, "b2s_kernel_hash: negative size");
  if (n == 0 || k == 0) return B2S_OK;
  B2S_REQUIRE(coords && offsets && out, B2S_ERR_INVALID, "b2s_kernel_hash: null pointer");
  B2S_REQUIRE(k <= 4096, B2S_ERR_UNSUPPORTED, "b2s_kernel_hash: kernel volume %d > 4096", k);
  kernel_hash_kernel<<<grid_for(n, 256), 256, 3 * k * sizeof(int32_t), as_stream(stream)>>>(
      reinterpret_cast<const int4*>(coords), n, offsets, k, out);
  B2S_CHECK_LAUNCH("b2s_kernel_hash");
  return B2S_OK;
}

int64_t b2s_table_slots(int64_t n) { return table_slots_for(n < 0 ? 0 : n); }
size_t b2s_table_bytes(int64_t n) { return (size_t)b2s_table_slots(n) * 12; }

int b2s_table_build(const int64_t* references, int64_t n, void* table, size_t table_bytes,
                    b2s_stream_t stream) {
  B2S_REQUIRE(n >= 0 && table, B2S_ERR_INVALID, "b2s_table_build: bad argument");
  B2S_REQUIRE(n < (1LL << 31), B2S_ERR_UNSUPPORTED, "b2s_table_build: n >= 2^31");
  B2S_REQUIRE(table_bytes >= b2s_table_bytes(n), B2S_ERR_WORKSPACE,
              "b2s_table_build: table needs %zu bytes", b2s_table_bytes(n));
  TableView t = table_view(table, n);
  table_clear(t, as_stream(stream));
  if (n > 0) {
    B2S_REQUIRE(references, B2S_ERR_INVALID, "b2s_table_build: null references");
    table_build_keys_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(references, n, t);
  }
  B2S_CHECK_LAUNCH("b2s_table_build");
  return B2S_OK;
}

int b2s_table_build_coords(const int32_t* coords, int64_t n, void* table, size_t table_bytes,
                           b2s_stream_t stream) {
  B2S_REQUIRE(n >= 0 && table, B2S_ERR_INVALID, "b2s_table_build_coords: bad argument");
  B2S_REQUIRE(n < (1LL << 31), B2S_ERR_UNSUPPORTED, "b2s_table_build_coords: n >= 2^31");
  B2S_REQUIRE(table_bytes >= b2s_table_bytes(n), B2S_ERR_WORKSPACE,
              "b2s_table_build_coords: table needs %zu bytes", b2s_table_bytes(n));
  TableView t = table_view(table, n);
  table_clear(t, as_stream(stream));
  if (n > 0) {
    B2S_REQUIRE(coords, B2S_ERR_INVALID, "b2s_table_build_coords: null coords");
    table_build_coords_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const int4*>(coords), n, t);
  }
  B2S_CHECK_LAUNCH("b2s_table_build_coords");
  return B2S_OK;
}

int b2s_table_query(const void* table, int64_t n_references, const int64_t* queries, int64_t nq,
                    int64_t* out, b2s_stream_t stream) {
  B2S_REQUIRE(table && nq >= 0 && n_references >= 0, B2S_ERR_INVALID,
              "b2s_table_query: bad argument");
  if (nq == 0) return B2S_OK;
  B2S_REQUIRE(queries && out, B2S_ERR_INVALID, "b2s_table_query: null pointer");
  TableView t = table_view(const_cast<void*>(table), n_references);
  table_query_kernel<<<grid_for(nq, 256), 256, 0, as_stream(stream)>>>(t, queries, nq, out);
  B2S_CHECK_LAUNCH("b2s_table_query");
  return B2S_OK;
}

int b2s_count(const int32_t* idx, int64_t n, int32_t* out, int64_t num, b2s_stream_t stream) {
  B2S_REQUIRE(n >= 0 && num >= 0, B2S_ERR_INVALID, "b2s_count: negative size");
  if (num == 0) return B2S_OK;
  B2S_REQUIRE(out, B2S_ERR_INVALID, "b2s_count: null out");
  cudaMemsetAsync(out, 0, num * sizeof(int32_t), as_stream(stream));
  if (n > 0) {
    B2S_REQUIRE(idx, B2S_ERR_INVALID, "b2s_count: null idx");
    count_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(idx, n, out, num);
  }
  B2S_CHECK_LAUNCH("b2s_count");
  return B2S_OK;
}

}  // extern "C"
