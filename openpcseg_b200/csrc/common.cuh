// Shared device/host helpers for libb2s (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b2s.h"

namespace b2s {

void set_error(const char* fmt, ...);

#define B2S_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      ::b2s::set_error(__VA_ARGS__);   \
      return (code);                   \
    }                                  \
  } while (0)

#define B2S_CHECK_LAUNCH(what)                                                      \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      ::b2s::set_error("%s: CUDA error %s", (what), cudaGetErrorString(e__));       \
      return B2S_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

static inline cudaStream_t as_stream(b2s_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Number of SMs of the current device (cached).  Grid sizes for the grid-stride
// kernels are multiples of it.
int sm_count();
// SMs the persistent (one CTA per SM slot) kernels fill: sm_count() minus the reserve set by
// b2s_set_sm_reserve / B2S_SM_RESERVE (room for concurrent collectives)
int persistent_sms();
void set_sm_reserve(int n);

constexpr int64_t kEmptyKey = -1;  // reserved table key (memset 0xFF); sphash values are < 2^60

// 60-bit folded FNV-1a over the four 32-bit words (x, y, z, b); bit-exact with the
// reference (TS/backend/hash/hash_cuda.cu:15-20).
__host__ __device__ __forceinline__ int64_t coord_hash(int x, int y, int z, int b) {
  uint64_t h = 14695981039346656037ULL;
  h = (h ^ (uint64_t)(uint32_t)x) * 1099511628211ULL;
  h = (h ^ (uint64_t)(uint32_t)y) * 1099511628211ULL;
  h = (h ^ (uint64_t)(uint32_t)z) * 1099511628211ULL;
  h = (h ^ (uint64_t)(uint32_t)b) * 1099511628211ULL;
  h = (h >> 60) ^ (h & 0x0FFFFFFFFFFFFFFFULL);
  return (int64_t)h;
}

// Slot scrambler (murmur3 finaliser) so arbitrary int64 keys spread over the table.
__host__ __device__ __forceinline__ uint32_t slot_hash(int64_t key) {
  uint64_t k = (uint64_t)key;
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return (uint32_t)k;
}

// Open-addressing table: `slots` keys followed by `slots` int32 values.
struct TableView {
  int64_t* keys;
  int32_t* vals;
  uint32_t mask;  // slots - 1 (slots is a power of two)
};
static inline int64_t table_slots_for(int64_t n) {
  int64_t s = 1024;
  while (s < 2 * n) s <<= 1;
  return s;
}
static inline TableView table_view(void* base, int64_t n_refs) {
  int64_t slots = table_slots_for(n_refs);
  TableView t;
  t.keys = reinterpret_cast<int64_t*>(base);
  t.vals = reinterpret_cast<int32_t*>(t.keys + slots);
  t.mask = (uint32_t)(slots - 1);
  return t;
}

__device__ __forceinline__ void table_insert(const TableView& t, int64_t key, int32_t val) {
  uint32_t s = slot_hash(key) & t.mask;
  while (true) {
    unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(t.keys + s),
                                        (unsigned long long)kEmptyKey, (unsigned long long)key);
    if (prev == (unsigned long long)kEmptyKey || prev == (unsigned long long)key) {
      atomicMin(t.vals + s, val);  // duplicates: smallest row wins (first-insert-wins)
      return;
    }
    s = (s + 1) & t.mask;
  }
}

__device__ __forceinline__ int32_t table_find(const TableView& t, int64_t key) {
  uint32_t s = slot_hash(key) & t.mask;
  while (true) {
    int64_t k = __ldg(t.keys + s);
    if (k == key) return __ldg(t.vals + s);
    if (k == kEmptyKey) return -1;
    s = (s + 1) & t.mask;
  }
}

// Launch helpers defined in hash_table.cu.
int grid_for(int64_t n, int threads);  // min(ceil(n/threads), 16 x SM count), >= 1
int table_clear(const TableView& t, cudaStream_t st);

template <typename T>
struct FeatIO;
template <>
struct FeatIO<float> {
  static __device__ __forceinline__ float load(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ void store(float* p, float v) { *p = v; }
};
template <>
struct FeatIO<__half> {
  static __device__ __forceinline__ float load(const __half* p) { return __half2float(__ldg(p)); }
  static __device__ __forceinline__ void store(__half* p, float v) { *p = __float2half_rn(v); }
};

}  // namespace b2s
