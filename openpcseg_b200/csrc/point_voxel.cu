// Point <-> voxel kernels: scatter-mean voxelize, trilinear devoxelize, the fused
// voxel_to_point map (8 corner probes + trilinear weights) and RPVNet's range-image ops.
//
// All HBM/L2-bound.  Rows are moved as 16-byte vectors (float4 / 8 x half) whenever the
// channel count allows; accumulation is always fp32.  Algorithmic bytes (e = element size):
//   voxelize fwd : (e C + 4) N_pts + 4 N_vox + e C N_vox
//   devoxelize   : (4 + e) 8 N_pts + e C 8 N_pts (gathers, mostly L2) + e C N_pts
#include "common.cuh"

namespace b2s {

template <typename T, int V>
struct Vec;
template <>
struct Vec<float, 4> {
  float4 raw;
  __device__ __forceinline__ void load(const float* p) { raw = __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = raw; }
  __device__ __forceinline__ float get(int i) const { return (&raw.x)[i]; }
  __device__ __forceinline__ void set(int i, float v) { (&raw.x)[i] = v; }
};
template <>
struct Vec<float, 1> {
  float raw;
  __device__ __forceinline__ void load(const float* p) { raw = __ldg(p); }
  __device__ __forceinline__ void store(float* p) const { *p = raw; }
  __device__ __forceinline__ float get(int) const { return raw; }
  __device__ __forceinline__ void set(int, float v) { raw = v; }
};
template <>
struct Vec<__half, 8> {
  uint4 raw;
  __device__ __forceinline__ void load(const __half* p) { raw = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void store(__half* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ float get(int i) const {
    return __half2float(reinterpret_cast<const __half*>(&raw)[i]);
  }
  __device__ __forceinline__ void set(int i, float v) {
    reinterpret_cast<__half*>(&raw)[i] = __float2half_rn(v);
  }
};
template <>
struct Vec<__half, 1> {
  __half raw;
  __device__ __forceinline__ void load(const __half* p) { raw = __ldg(p); }
  __device__ __forceinline__ void store(__half* p) const { *p = raw; }
  __device__ __forceinline__ float get(int) const { return __half2float(raw); }
  __device__ __forceinline__ void set(int, float v) { raw = __float2half_rn(v); }
};

// fp32 scatter-add of V consecutive channels (one 16 B vector red when V == 4)
template <int V>
__device__ __forceinline__ void red_add(float* dst, const float* v) {
  if constexpr (V % 4 == 0) {
#pragma unroll
    for (int i = 0; i < V; i += 4)
      atomicAdd(reinterpret_cast<float4*>(dst + i), make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) atomicAdd(dst + i, v[i]);
  }
}

// out/acc[idx[i], :] += feats[i, :] / count[idx[i]]      (TS/backend/voxelize/voxelize_cuda.cu:12-25)
template <typename T, int V>
__global__ void __launch_bounds__(256) voxelize_fwd_kernel(const T* __restrict__ feats,
                                                            const int32_t* __restrict__ idx,
                                                            const int32_t* __restrict__ counts,
                                                            int64_t n_pts, int64_t n_vox, int c,
                                                            float* __restrict__ acc) {
  const int groups = c / V;
  const int64_t total = n_pts * groups;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / groups;
    int ch = (int)(t - i * groups) * V;
    int32_t pos = __ldg(idx + i);
    if (pos < 0 || pos >= n_vox) continue;
    int32_t cnt = __ldg(counts + pos);
    if (cnt == 0) continue;
    Vec<T, V> v;
    v.load(feats + i * c + ch);
    float r[V];
    const float fc = (float)cnt;
#pragma unroll
    for (int j = 0; j < V; ++j) r[j] = v.get(j) / fc;
    red_add<V>(acc + (int64_t)pos * c + ch, r);
  }
}

// grad_pts[i, :] = grad_vox[idx[i], :] / count            (voxelize_cuda.cu:28-42; a pure gather)
template <typename T, int V>
__global__ void __launch_bounds__(256) voxelize_bwd_kernel(const T* __restrict__ grad_vox,
                                                            const int32_t* __restrict__ idx,
                                                            const int32_t* __restrict__ counts,
                                                            int64_t n_pts, int64_t n_vox, int c,
                                                            T* __restrict__ grad_pts) {
  const int groups = c / V;
  const int64_t total = n_pts * groups;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / groups;
    int ch = (int)(t - i * groups) * V;
    int32_t pos = __ldg(idx + i);
    int32_t cnt = (pos >= 0 && pos < n_vox) ? __ldg(counts + pos) : 0;
    Vec<T, V> v;
    if (cnt != 0) {
      v.load(grad_vox + (int64_t)pos * c + ch);
      const float fc = (float)cnt;
#pragma unroll
      for (int j = 0; j < V; ++j) v.set(j, v.get(j) / fc);
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) v.set(j, 0.f);
    }
    v.store(grad_pts + i * c + ch);
  }
}

// out[p, :] = sum_k w[p,k] * feats[idx[p,k], :]            (TS/backend/devoxelize/devoxelize_cuda.cu:11-34)
template <typename T, int V>
__global__ void __launch_bounds__(256) devoxelize_fwd_kernel(const T* __restrict__ feats,
                                                              const int32_t* __restrict__ idx,
                                                              const float* __restrict__ w,
                                                              int64_t n_pts, int c,
                                                              T* __restrict__ out) {
  const int groups = c / V;
  const int64_t total = n_pts * groups;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = t / groups;
    int ch = (int)(t - p * groups) * V;
    float accum[V];
#pragma unroll
    for (int j = 0; j < V; ++j) accum[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int32_t r = __ldg(idx + p * 8 + k);
      if (r < 0) continue;
      float wk = __ldg(w + p * 8 + k);
      if (wk == 0.f) continue;       // on-grid points: 7 of the 8 trilinear weights are exactly 0
      Vec<T, V> v;
      v.load(feats + (int64_t)r * c + ch);
#pragma unroll
      for (int j = 0; j < V; ++j) accum[j] = fmaf(wk, v.get(j), accum[j]);
    }
    Vec<T, V> o;
#pragma unroll
    for (int j = 0; j < V; ++j) o.set(j, accum[j]);
    o.store(out + p * c + ch);
  }
}

// acc[idx[p,k], :] += w[p,k] * grad_pts[p, :]              (devoxelize_cuda.cu:37-58)
template <typename T, int V>
__global__ void __launch_bounds__(256) devoxelize_bwd_kernel(const T* __restrict__ grad_pts,
                                                              const int32_t* __restrict__ idx,
                                                              const float* __restrict__ w,
                                                              int64_t n_pts, int c,
                                                              float* __restrict__ acc) {
  const int groups = c / V;
  const int64_t total = n_pts * groups;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = t / groups;
    int ch = (int)(t - p * groups) * V;
    Vec<T, V> g;
    g.load(grad_pts + p * c + ch);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int32_t r = __ldg(idx + p * 8 + k);
      if (r < 0) continue;
      float wk = __ldg(w + p * 8 + k);
      if (wk == 0.f) continue;
      float v[V];
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = wk * g.get(j);
      red_add<V>(acc + (int64_t)r * c + ch, v);
    }
  }
}

// The same scatter for CONTENDED maps (coarse strides: ~27 points per stride-16 voxel, every one of them adding
// into the same 8 corner rows - the plain kernel runs at the L2 atomic rate).  Points are visited in `order`
// (sorted by their corner-0 voxel, so neighbours in the order share all eight corners); a thread owns one
// V-channel group, walks kPts consecutive points and keeps one running sum per corner in registers, flushing a
// corner with a vector red only when its voxel changes: ~run-length times fewer atomics.
template <typename T, int V, int kPts>
__global__ void __launch_bounds__(128) devoxelize_bwd_sorted_kernel(const T* __restrict__ grad_pts,
                                                                     const int32_t* __restrict__ order,
                                                                     const int32_t* __restrict__ idx,
                                                                     const float* __restrict__ w, int64_t n_pts,
                                                                     int c, float* __restrict__ acc) {
  const int groups = c / V;
  const int slices = blockDim.x / groups;                   // point slices per block
  const int cg = threadIdx.x % groups, sl = threadIdx.x / groups;
  if (sl >= slices) return;
  const int64_t p0 = ((int64_t)blockIdx.x * slices + sl) * kPts;
  if (p0 >= n_pts) return;
  const int64_t p1 = p0 + kPts < n_pts ? p0 + kPts : n_pts;
  const int ch = cg * V;
  float run[8][V];
  int32_t cur[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    cur[k] = -1;
#pragma unroll
    for (int j = 0; j < V; ++j) run[k][j] = 0.f;
  }
  for (int64_t q = p0; q < p1; ++q) {
    const int64_t p = __ldg(order + q);
    Vec<T, V> g;
    g.load(grad_pts + p * c + ch);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int32_t r = __ldg(idx + p * 8 + k);
      const float wk = r >= 0 ? __ldg(w + p * 8 + k) : 0.f;
      if (r < 0 || wk == 0.f) continue;
      if (r != cur[k]) {
        if (cur[k] >= 0) red_add<V>(acc + (int64_t)cur[k] * c + ch, run[k]);
        cur[k] = r;
#pragma unroll
        for (int j = 0; j < V; ++j) run[k][j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < V; ++j) run[k][j] = fmaf(wk, g.get(j), run[k][j]);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (cur[k] >= 0) red_add<V>(acc + (int64_t)cur[k] * c + ch, run[k]);
}

__global__ void __launch_bounds__(256) f32_to_f16_kernel(const float* __restrict__ src, int64_t n,
                                                          __half* __restrict__ dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = __float2half_rn(src[i]);
}

// Trilinear weights of one point, fp32 arithmetic in the reference's order
// (TS/nn/functional/devoxelize.py:10-48).  `hit` tells which corners exist.
__device__ __forceinline__ void ti_weights(float x, float y, float z, float scale, const bool* hit,
                                           float* w) {
  float xf, yf, zf;
  if (scale != 1.f) {
    xf = floorf(x / scale) * scale;
    yf = floorf(y / scale) * scale;
    zf = floorf(z / scale) * scale;
  } else {
    xf = floorf(x);
    yf = floorf(y);
    zf = floorf(z);
  }
  const float xc = xf + scale, yc = yf + scale, zc = zf + scale;
  const float ax[2] = {xc - x, x - xf}, ay[2] = {yc - y, y - yf}, az[2] = {zc - z, z - zf};
  const float vol = scale * scale * scale;
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v = ax[(j >> 2) & 1] * ay[(j >> 1) & 1] * az[j & 1];
    if (scale != 1.f) v = v / vol;
    if (!hit[j]) v = 0.f;
    w[j] = v;
    sum += v;
  }
  const float den = sum + 1e-8f;
#pragma unroll
  for (int j = 0; j < 8; ++j) w[j] = w[j] / den;
}

// Fused voxel_to_point map: floor to the stride grid, probe the 8 corners
// (get_kernel_offsets(2, s): z fastest), weights.  minkunet/utils.py:73-81.
__global__ void __launch_bounds__(256) trilinear_map_kernel(const float4* __restrict__ pts,
                                                             int64_t n_pts, int stride,
                                                             TableView table,
                                                             int32_t* __restrict__ idx,
                                                             float* __restrict__ w) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_pts;
       p += (int64_t)gridDim.x * blockDim.x) {
    float4 q = __ldg(pts + p);
    const float fs = (float)stride;
    const int bx = (int)floorf(q.x / fs) * stride;
    const int by = (int)floorf(q.y / fs) * stride;
    const int bz = (int)floorf(q.z / fs) * stride;
    const int b = (int)q.w;
    int32_t r[8];
    bool hit[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      r[j] = table_find(table, coord_hash(bx + ((j >> 2) & 1) * stride, by + ((j >> 1) & 1) * stride,
                                          bz + (j & 1) * stride, b));
      hit[j] = r[j] >= 0;
    }
    float wt[8];
    ti_weights(q.x, q.y, q.z, fs, hit, wt);
    int4* ip = reinterpret_cast<int4*>(idx + p * 8);
    ip[0] = make_int4(r[0], r[1], r[2], r[3]);
    ip[1] = make_int4(r[4], r[5], r[6], r[7]);
    float4* wp = reinterpret_cast<float4*>(w + p * 8);
    wp[0] = make_float4(wt[0], wt[1], wt[2], wt[3]);
    wp[1] = make_float4(wt[4], wt[5], wt[6], wt[7]);
  }
}

__global__ void __launch_bounds__(256) ti_weights_kernel(const float4* __restrict__ pts,
                                                          int64_t n_pts,
                                                          const int64_t* __restrict__ idx,
                                                          float scale, float* __restrict__ w) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_pts;
       p += (int64_t)gridDim.x * blockDim.x) {
    float4 q = __ldg(pts + p);
    bool hit[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) hit[j] = __ldg(idx + (int64_t)j * n_pts + p) != -1;
    float wt[8];
    ti_weights(q.x, q.y, q.z, scale, hit, wt);
#pragma unroll
    for (int j = 0; j < 8; ++j) w[(int64_t)j * n_pts + p] = wt[j];
  }
}

// ------------------------------------------------------------ range-image ops (RPVNet)
__global__ void __launch_bounds__(256) map_count_kernel(const int32_t* __restrict__ pxpy,
                                                         int64_t n, int h, int w,
                                                         int32_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int bs = pxpy[3 * i], px = pxpy[3 * i + 1], py = pxpy[3 * i + 2];
    if (px >= 0 && py >= 0) atomicAdd(out + ((int64_t)bs * h + py) * w + px, 1);
  }
}

__global__ void __launch_bounds__(256) denselize_fwd_kernel(const float* __restrict__ feats,
                                                             const int32_t* __restrict__ pxpy,
                                                             const int32_t* __restrict__ cmap,
                                                             int64_t n, int c, int h, int w,
                                                             float* __restrict__ dense) {
  const int64_t total = n * c;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / c;
    int j = (int)(t - i * c);
    int bs = pxpy[3 * i], px = pxpy[3 * i + 1], py = pxpy[3 * i + 2];
    int64_t cell = (int64_t)py * w + px;
    int64_t pos = (int64_t)bs * h * w + cell;
    if (pos < 0) continue;
    int cnt = cmap[pos];
    if (cnt == 0) continue;
    atomicAdd(dense + ((int64_t)bs * c + j) * h * w + cell, feats[t] / (float)cnt);
  }
}

__global__ void __launch_bounds__(256) denselize_bwd_kernel(const float* __restrict__ gdense,
                                                             const int32_t* __restrict__ pxpy,
                                                             const int32_t* __restrict__ cmap,
                                                             int64_t n, int c, int h, int w,
                                                             float* __restrict__ gfeats) {
  const int64_t total = n * c;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / c;
    int j = (int)(t - i * c);
    int bs = pxpy[3 * i], px = pxpy[3 * i + 1], py = pxpy[3 * i + 2];
    int64_t cell = (int64_t)py * w + px;
    int64_t pos = (int64_t)bs * h * w + cell;
    int cnt = cmap[pos];
    gfeats[t] = cnt ? gdense[((int64_t)bs * c + j) * h * w + cell] / (float)cnt : 0.f;
  }
}

// ---------------------------------------------------------------- scatter-max (Cylinder3D)
// out[idx[i], j] = max_i feats[i, j]; arg[idx[i], j] = smallest i attaining it (for backward).
// Replaces torch_scatter.scatter_max in seg_utils.voxelize / initial_voxelize_max
// (tools/utils/common/seg_utils.py:172-188, cylinder_ts.py:24-43).  fp32 keys are mapped to
// order-preserving unsigned ints so one atomicMax per element suffices.
__device__ __forceinline__ uint32_t float_key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

template <typename T>
__global__ void __launch_bounds__(256) scatter_max_key_kernel(const T* __restrict__ feats,
                                                              const int64_t* __restrict__ idx, int64_t n,
                                                              int c, int64_t m, uint32_t* __restrict__ keys) {
  const int64_t total = n * c;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / c;
    const int j = (int)(t - i * c);
    const int64_t v = __ldg(idx + i);
    if (v < 0 || v >= m) continue;
    atomicMax(keys + v * c + j, float_key(FeatIO<T>::load(feats + t)));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) scatter_max_arg_kernel(const T* __restrict__ feats,
                                                              const int64_t* __restrict__ idx, int64_t n,
                                                              int c, int64_t m,
                                                              const uint32_t* __restrict__ keys,
                                                              int64_t* __restrict__ arg) {
  const int64_t total = n * c;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / c;
    const int j = (int)(t - i * c);
    const int64_t v = __ldg(idx + i);
    if (v < 0 || v >= m) continue;
    if (float_key(FeatIO<T>::load(feats + t)) == keys[v * c + j])
      atomicMin(reinterpret_cast<unsigned long long*>(arg + v * c + j), (unsigned long long)i);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) scatter_max_finish_kernel(const uint32_t* __restrict__ keys,
                                                                 int64_t* __restrict__ arg, int64_t total,
                                                                 int64_t n, T* __restrict__ out) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const bool empty = keys[t] == 0u;          // no row mapped here: torch_scatter leaves 0 / arg = n
    FeatIO<T>::store(out + t, empty ? 0.f : key_float(keys[t]));
    if (empty) arg[t] = n;
  }
}

template <typename T>
constexpr int vec_width() { return sizeof(T) == 4 ? 4 : 8; }

#define B2S_DISPATCH_VEC(T, c, ptrs_aligned, CALL)          \
  do {                                                      \
    constexpr int VW = vec_width<T>();                      \
    if ((c) % VW == 0 && (ptrs_aligned)) {                  \
      constexpr int V = VW;                                 \
      CALL;                                                 \
    } else {                                                \
      constexpr int V = 1;                                  \
      CALL;                                                 \
    }                                                       \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace b2s

using namespace b2s;

extern "C" {

int b2s_voxelize_fwd(int32_t dtype, const void* feats, const int32_t* idx, const int32_t* counts,
                     int64_t n_pts, int64_t n_vox, int32_t c, void* out, float* acc,
                     b2s_stream_t stream) {
  B2S_REQUIRE(n_pts >= 0 && n_vox >= 0 && c >= 1, B2S_ERR_INVALID, "b2s_voxelize_fwd: bad sizes");
  B2S_REQUIRE(dtype == B2S_F32 || dtype == B2S_F16, B2S_ERR_INVALID, "b2s_voxelize_fwd: dtype");
  if (n_vox == 0) return B2S_OK;
  B2S_REQUIRE(out && counts && (dtype == B2S_F32 || acc), B2S_ERR_INVALID,
              "b2s_voxelize_fwd: null pointer (fp16 needs the fp32 scratch)");
  cudaStream_t st = as_stream(stream);
  float* target = dtype == B2S_F32 ? reinterpret_cast<float*>(out) : acc;
  cudaMemsetAsync(target, 0, (size_t)n_vox * c * sizeof(float), st);
  if (n_pts > 0) {
    B2S_REQUIRE(feats && idx, B2S_ERR_INVALID, "b2s_voxelize_fwd: null pointer");
    bool al = aligned16(feats) && aligned16(target);
    if (dtype == B2S_F32) {
      B2S_DISPATCH_VEC(float, c, al, (voxelize_fwd_kernel<float, V><<<grid_for(n_pts * (c / V), 256), 256, 0, st>>>(
          reinterpret_cast<const float*>(feats), idx, counts, n_pts, n_vox, c, target)));
    } else {
      B2S_DISPATCH_VEC(__half, c, al, (voxelize_fwd_kernel<__half, V><<<grid_for(n_pts * (c / V), 256), 256, 0, st>>>(
          reinterpret_cast<const __half*>(feats), idx, counts, n_pts, n_vox, c, target)));
    }
  }
  if (dtype == B2S_F16)
    f32_to_f16_kernel<<<grid_for(n_vox * c, 256), 256, 0, st>>>(acc, n_vox * c,
                                                                reinterpret_cast<__half*>(out));
  B2S_CHECK_LAUNCH("b2s_voxelize_fwd");
  return B2S_OK;
}

int b2s_voxelize_bwd(int32_t dtype, const void* grad_vox, const int32_t* idx,
                     const int32_t* counts, int64_t n_pts, int64_t n_vox, int32_t c,
                     void* grad_pts, b2s_stream_t stream) {
  B2S_REQUIRE(n_pts >= 0 && n_vox >= 0 && c >= 1, B2S_ERR_INVALID, "b2s_voxelize_bwd: bad sizes");
  B2S_REQUIRE(dtype == B2S_F32 || dtype == B2S_F16, B2S_ERR_INVALID, "b2s_voxelize_bwd: dtype");
  if (n_pts == 0) return B2S_OK;
  B2S_REQUIRE(grad_pts && idx && (n_vox == 0 || (grad_vox && counts)), B2S_ERR_INVALID,
              "b2s_voxelize_bwd: null pointer");
  cudaStream_t st = as_stream(stream);
  bool al = aligned16(grad_vox) && aligned16(grad_pts);
  if (dtype == B2S_F32) {
    B2S_DISPATCH_VEC(float, c, al, (voxelize_bwd_kernel<float, V><<<grid_for(n_pts * (c / V), 256), 256, 0, st>>>(
        reinterpret_cast<const float*>(grad_vox), idx, counts, n_pts, n_vox, c,
        reinterpret_cast<float*>(grad_pts))));
  } else {
    B2S_DISPATCH_VEC(__half, c, al, (voxelize_bwd_kernel<__half, V><<<grid_for(n_pts * (c / V), 256), 256, 0, st>>>(
        reinterpret_cast<const __half*>(grad_vox), idx, counts, n_pts, n_vox, c,
        reinterpret_cast<__half*>(grad_pts))));
  }
  B2S_CHECK_LAUNCH("b2s_voxelize_bwd");
  return B2S_OK;
}

int b2s_devoxelize_fwd(int32_t dtype, const void* feats, const int32_t* idx, const float* weights,
                       int64_t n_pts, int64_t n_vox, int32_t c, void* out, b2s_stream_t stream) {
  B2S_REQUIRE(n_pts >= 0 && n_vox >= 0 && c >= 1, B2S_ERR_INVALID, "b2s_devoxelize_fwd: bad sizes");
  B2S_REQUIRE(dtype == B2S_F32 || dtype == B2S_F16, B2S_ERR_INVALID, "b2s_devoxelize_fwd: dtype");
  if (n_pts == 0) return B2S_OK;
  B2S_REQUIRE(out && idx && weights && (n_vox == 0 || feats), B2S_ERR_INVALID,
              "b2s_devoxelize_fwd: null pointer");
  cudaStream_t st = as_stream(stream);
  bool al = aligned16(feats) && aligned16(out);
  if (dtype == B2S_F32) {
    B2S_DISPATCH_VEC(float, c, al, (devoxelize_fwd_kernel<float, V><<<grid_for(n_pts * (c / V), 256), 256, 0, st>>>(
        reinterpret_cast<const float*>(feats), idx, weights, n_pts,
        c, reinterpret_cast<float*>(out))));
  } else {
    B2S_DISPATCH_VEC(__half, c, al, (devoxelize_fwd_kernel<__half, V><<<grid_for(n_pts * (c / V), 256), 256, 0, st>>>(
        reinterpret_cast<const __half*>(feats), idx, weights,
        n_pts, c, reinterpret_cast<__half*>(out))));
  }
  B2S_CHECK_LAUNCH("b2s_devoxelize_fwd");
  return B2S_OK;
}

int b2s_devoxelize_bwd(int32_t dtype, const void* grad_pts, const int32_t* idx,
                       const float* weights, int64_t n_pts, int64_t n_vox, int32_t c,
                       void* grad_vox, float* acc, b2s_stream_t stream) {
  B2S_REQUIRE(n_pts >= 0 && n_vox >= 0 && c >= 1, B2S_ERR_INVALID, "b2s_devoxelize_bwd: bad sizes");
  B2S_REQUIRE(dtype == B2S_F32 || dtype == B2S_F16, B2S_ERR_INVALID, "b2s_devoxelize_bwd: dtype");
  if (n_vox == 0) return B2S_OK;
  B2S_REQUIRE(grad_vox && (dtype == B2S_F32 || acc), B2S_ERR_INVALID,
              "b2s_devoxelize_bwd: null pointer (fp16 needs the fp32 scratch)");
  cudaStream_t st = as_stream(stream);
  float* target = dtype == B2S_F32 ? reinterpret_cast<float*>(grad_vox) : acc;
  cudaMemsetAsync(target, 0, (size_t)n_vox * c * sizeof(float), st);
  if (n_pts > 0) {
    B2S_REQUIRE(grad_pts && idx && weights, B2S_ERR_INVALID, "b2s_devoxelize_bwd: null pointer");
    bool al = aligned16(grad_pts) && aligned16(target);
    if (dtype == B2S_F32) {
      B2S_DISPATCH_VEC(float, c, al, (devoxelize_bwd_kernel<float, V><<<grid_for(n_pts * (c / V), 256), 256, 0, st>>>(
          reinterpret_cast<const float*>(grad_pts), idx, weights,
          n_pts, c, target)));
    } else {
      B2S_DISPATCH_VEC(__half, c, al, (devoxelize_bwd_kernel<__half, V><<<grid_for(n_pts * (c / V), 256), 256, 0, st>>>(
          reinterpret_cast<const __half*>(grad_pts), idx, weights,
          n_pts, c, target)));
    }
  }
  if (dtype == B2S_F16)
    f32_to_f16_kernel<<<grid_for(n_vox * c, 256), 256, 0, st>>>(acc, n_vox * c,
                                                                reinterpret_cast<__half*>(grad_vox));
  B2S_CHECK_LAUNCH("b2s_devoxelize_bwd");
  return B2S_OK;
}

int b2s_devoxelize_bwd_sorted(int32_t dtype, const void* grad_pts, const int32_t* order, const int32_t* idx,
                              const float* weights, int64_t n_pts, int64_t n_vox, int32_t c, void* grad_vox,
                              float* acc, b2s_stream_t stream) {
  B2S_REQUIRE(n_pts >= 0 && n_vox >= 0 && c >= 1, B2S_ERR_INVALID, "b2s_devoxelize_bwd_sorted: bad sizes");
  B2S_REQUIRE(dtype == B2S_F32 || dtype == B2S_F16, B2S_ERR_INVALID, "b2s_devoxelize_bwd_sorted: dtype");
  const int v = dtype == B2S_F16 ? 8 : 4;
  B2S_REQUIRE(c % v == 0 && c / v <= 128 && aligned16(grad_pts), B2S_ERR_UNSUPPORTED,
              "b2s_devoxelize_bwd_sorted: C=%d must be a multiple of the 16-byte vector (use b2s_devoxelize_bwd)", c);
  if (n_vox == 0) return B2S_OK;
  B2S_REQUIRE(grad_vox && (dtype == B2S_F32 || acc), B2S_ERR_INVALID,
              "b2s_devoxelize_bwd_sorted: null pointer (fp16 needs the fp32 scratch)");
  cudaStream_t st = as_stream(stream);
  float* target = dtype == B2S_F32 ? reinterpret_cast<float*>(grad_vox) : acc;
  cudaMemsetAsync(target, 0, (size_t)n_vox * c * sizeof(float), st);
  if (n_pts > 0) {
    B2S_REQUIRE(grad_pts && order && idx && weights, B2S_ERR_INVALID, "b2s_devoxelize_bwd_sorted: null pointer");
    constexpr int kPts = 32;
    const int groups = c / v;
    const int slices = 128 / groups;
    const unsigned grid = (unsigned)ceil_div(n_pts, (int64_t)slices * kPts);
    if (dtype == B2S_F16)
      devoxelize_bwd_sorted_kernel<__half, 8, kPts><<<grid, 128, 0, st>>>(
          reinterpret_cast<const __half*>(grad_pts), order, idx, weights, n_pts, c, target);
    else
      devoxelize_bwd_sorted_kernel<float, 4, kPts><<<grid, 128, 0, st>>>(
          reinterpret_cast<const float*>(grad_pts), order, idx, weights, n_pts, c, target);
  }
  if (dtype == B2S_F16)
    f32_to_f16_kernel<<<grid_for(n_vox * c, 256), 256, 0, st>>>(acc, n_vox * c,
                                                                reinterpret_cast<__half*>(grad_vox));
  B2S_CHECK_LAUNCH("b2s_devoxelize_bwd_sorted");
  return B2S_OK;
}

int b2s_scatter_max(int32_t dtype, const void* feats, const int64_t* idx, int64_t n, int32_t c, int64_t m,
                    void* out, int64_t* arg, uint32_t* keys, b2s_stream_t stream) {
  B2S_REQUIRE(dtype == B2S_F32 || dtype == B2S_F16, B2S_ERR_INVALID, "b2s_scatter_max: dtype");
  B2S_REQUIRE(n >= 0 && c >= 1 && m >= 0, B2S_ERR_INVALID, "b2s_scatter_max: bad sizes");
  if (m == 0) return B2S_OK;
  B2S_REQUIRE(out && arg && keys && (n == 0 || (feats && idx)), B2S_ERR_INVALID,
              "b2s_scatter_max: null pointer");
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(keys, 0, (size_t)m * c * sizeof(uint32_t), st);      // key 0 < every float key
  cudaMemsetAsync(arg, 0x7F, (size_t)m * c * sizeof(int64_t), st);     // +inf for atomicMin
  if (n > 0) {
    const int g = grid_for(n * c, 256);
    if (dtype == B2S_F32) {
      scatter_max_key_kernel<float><<<g, 256, 0, st>>>(reinterpret_cast<const float*>(feats), idx, n, c, m, keys);
      scatter_max_arg_kernel<float><<<g, 256, 0, st>>>(reinterpret_cast<const float*>(feats), idx, n, c, m, keys, arg);
    } else {
      scatter_max_key_kernel<__half><<<g, 256, 0, st>>>(reinterpret_cast<const __half*>(feats), idx, n, c, m, keys);
      scatter_max_arg_kernel<__half><<<g, 256, 0, st>>>(reinterpret_cast<const __half*>(feats), idx, n, c, m, keys, arg);
    }
  }
  const int g2 = grid_for(m * c, 256);
  if (dtype == B2S_F32)
    scatter_max_finish_kernel<float><<<g2, 256, 0, st>>>(keys, arg, m * c, n, reinterpret_cast<float*>(out));
  else
    scatter_max_finish_kernel<__half><<<g2, 256, 0, st>>>(keys, arg, m * c, n, reinterpret_cast<__half*>(out));
  B2S_CHECK_LAUNCH("b2s_scatter_max");
  return B2S_OK;
}

int b2s_trilinear_map(const float* pts, int64_t n_pts, int32_t stride, const void* table,
                      int64_t n_vox, int32_t* idx, float* w, b2s_stream_t stream) {
  B2S_REQUIRE(n_pts >= 0 && stride >= 1 && n_vox >= 0, B2S_ERR_INVALID,
              "b2s_trilinear_map: bad argument");
  if (n_pts == 0) return B2S_OK;
  B2S_REQUIRE(pts && table && idx && w, B2S_ERR_INVALID, "b2s_trilinear_map: null pointer");
  TableView t = table_view(const_cast<void*>(table), n_vox);
  trilinear_map_kernel<<<grid_for(n_pts, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(pts), n_pts, stride, t, idx, w);
  B2S_CHECK_LAUNCH("b2s_trilinear_map");
  return B2S_OK;
}

int b2s_ti_weights(const float* pts, int64_t n_pts, const int64_t* idx, float scale, float* w,
                   b2s_stream_t stream) {
  B2S_REQUIRE(n_pts >= 0 && scale > 0.f, B2S_ERR_INVALID, "b2s_ti_weights: bad argument");
  if (n_pts == 0) return B2S_OK;
  B2S_REQUIRE(pts && idx && w, B2S_ERR_INVALID, "b2s_ti_weights: null pointer");
  ti_weights_kernel<<<grid_for(n_pts, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(pts), n_pts, idx, scale, w);
  B2S_CHECK_LAUNCH("b2s_ti_weights");
  return B2S_OK;
}

int b2s_map_count(const int32_t* pxpy, int64_t n, int32_t b, int32_t h, int32_t w,
                  int32_t* count_map, b2s_stream_t stream) {
  B2S_REQUIRE(n >= 0 && b >= 1 && h >= 1 && w >= 1 && count_map, B2S_ERR_INVALID,
              "b2s_map_count: bad argument");
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(count_map, 0, (size_t)b * h * w * sizeof(int32_t), st);
  if (n > 0) {
    B2S_REQUIRE(pxpy, B2S_ERR_INVALID, "b2s_map_count: null pxpy");
    map_count_kernel<<<grid_for(n, 256), 256, 0, st>>>(pxpy, n, h, w, count_map);
  }
  B2S_CHECK_LAUNCH("b2s_map_count");
  return B2S_OK;
}

int b2s_denselize_fwd(const float* feats, const int32_t* pxpy, const int32_t* count_map,
                      int64_t n, int32_t c, int32_t b, int32_t h, int32_t w, float* dense,
                      b2s_stream_t stream) {
  B2S_REQUIRE(n >= 0 && c >= 1 && b >= 1 && h >= 1 && w >= 1 && dense && count_map,
              B2S_ERR_INVALID, "b2s_denselize_fwd: bad argument");
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(dense, 0, (size_t)b * c * h * w * sizeof(float), st);
  if (n > 0) {
    B2S_REQUIRE(feats && pxpy, B2S_ERR_INVALID, "b2s_denselize_fwd: null pointer");
    denselize_fwd_kernel<<<grid_for(n * c, 256), 256, 0, st>>>(feats, pxpy, count_map, n, c, h, w,
                                                               dense);
  }
  B2S_CHECK_LAUNCH("b2s_denselize_fwd");
  return B2S_OK;
}

int b2s_denselize_bwd(const float* grad_dense, const int32_t* pxpy, const int32_t* count_map,
                      int64_t n, int32_t c, int32_t b, int32_t h, int32_t w, float* grad_feats,
                      b2s_stream_t stream) {
  B2S_REQUIRE(n >= 0 && c >= 1 && b >= 1 && h >= 1 && w >= 1, B2S_ERR_INVALID,
              "b2s_denselize_bwd: bad argument");
  if (n == 0) return B2S_OK;
  B2S_REQUIRE(grad_dense && pxpy && count_map && grad_feats, B2S_ERR_INVALID,
              "b2s_denselize_bwd: null pointer");
  denselize_bwd_kernel<<<grid_for(n * c, 256), 256, 0, as_stream(stream)>>>(
      grad_dense, pxpy, count_map, n, c, h, w, grad_feats);
  B2S_CHECK_LAUNCH("b2s_denselize_bwd");
  return B2S_OK;
}

}  // extern "C"
