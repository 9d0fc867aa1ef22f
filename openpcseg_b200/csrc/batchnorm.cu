// Fused batch-norm kernels for row-major [N, C] voxel features (C = 32..512, N = 10^4..10^6):
// training-mode BatchNorm1d (+ optional residual add, + optional ReLU) in two passes forward
// (statistics, apply) and two passes backward (reduce, apply).  These are the "next" row N1
// of SURVEY.md section 8f: after the sparse convs, BN/ReLU/add are the remaining full [N, C]
// HBM passes of the voxel segmentors (reference: nn.BatchNorm1d via fapply,
// pcseg/model/segmentor/voxel/minkunet/minkunet.py:27-29, relu/add at :134-136).
//
// HBM-bound.  Algorithmic bytes (e = element size): stats e*N*C; apply (1 [+1 residual] + 1)
// e*N*C; bwd reduce 2-3 e*N*C; bwd apply 3-4 e*N*C + e*N*C (+ e*N*C residual grad).
// Thread layout: a thread owns one 16-byte channel group and walks rows, so per-channel
// coefficients stay in registers and every access is a coalesced 16-byte vector.
#include "common.cuh"

namespace b2s {

template <typename T>
struct VecT;
template <>
struct VecT<__half> {
  static constexpr int W = 8;
  uint4 raw;
  __device__ __forceinline__ void load(const __half* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__half* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ float get(int i) const { return __half2float(reinterpret_cast<const __half*>(&raw)[i]); }
  __device__ __forceinline__ void set(int i, float v) { reinterpret_cast<__half*>(&raw)[i] = __float2half_rn(v); }
};
template <>
struct VecT<float> {
  static constexpr int W = 4;
  float4 raw;
  __device__ __forceinline__ void load(const float* p) { raw = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = raw; }
  __device__ __forceinline__ float get(int i) const { return (&raw.x)[i]; }
  __device__ __forceinline__ void set(int i, float v) { (&raw.x)[i] = v; }
};

constexpr int kBnThreads = 256;

// sums[0][c] += sum_n x[n][c];  sums[1][c] += sum_n x[n][c]^2   (fp64 accumulators)
template <typename T>
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const T* __restrict__ x, int64_t n, int c,
                                                              double* __restrict__ sums) {
  constexpr int W = VecT<T>::W;
  extern __shared__ float s_part[];                 // [2][c]
  for (int t = threadIdx.x; t < 2 * c; t += blockDim.x) s_part[t] = 0.f;
  __syncthreads();
  const int groups = c / W;
  const int cg = threadIdx.x % groups;
  const int rows_per_block = blockDim.x / groups;
  const int rl = threadIdx.x / groups;
  float s[W], q[W];
#pragma unroll
  for (int j = 0; j < W; ++j) s[j] = q[j] = 0.f;
  if (rl < rows_per_block) {
    // four independent 16-byte loads in flight per thread (one per iteration left this pass at
    // 3.5 TB/s: latency-bound, profiles/r1_launches_final.txt)
    const int64_t step = (int64_t)gridDim.x * rows_per_block;
    int64_t r = (int64_t)blockIdx.x * rows_per_block + rl;
    for (; r + 3 * step < n; r += 4 * step) {
      VecT<T> v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u].load(x + (r + u * step) * c + cg * W);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < W; ++j) {
          const float f = v[u].get(j);
          s[j] += f;
          q[j] = fmaf(f, f, q[j]);
        }
    }
    for (; r < n; r += step) {
      VecT<T> v;
      v.load(x + r * c + cg * W);
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const float f = v.get(j);
        s[j] += f;
        q[j] = fmaf(f, f, q[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < W; ++j) {
      atomicAdd(&s_part[cg * W + j], s[j]);
      atomicAdd(&s_part[c + cg * W + j], q[j]);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * c; t += blockDim.x) atomicAdd(sums + t, (double)s_part[t]);
}

// mean / invstd / scale / shift from the sums; running statistics like nn.BatchNorm1d
// (momentum update with the unbiased variance).
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int64_t n_host, const double* __restrict__ n_dev,
                                   int c,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float eps, float momentum, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out, float* __restrict__ scale,
                                   float* __restrict__ shift) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  // n_dev: the (all-reduced) global row count of synchronised batch norm, else the local row count
  const double n = n_dev ? *n_dev : (double)n_host;
  const double inv_n = 1.0 / n;
  const double m = sums[ch] * inv_n;
  double var = sums[c + ch] * inv_n - m * m;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
  mean_out[ch] = (float)m;
  invstd_out[ch] = invstd;
  scale[ch] = g * invstd;
  shift[ch] = b - (float)m * g * invstd;
  if (running_mean) {
    const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
    running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)m;
    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
  }
}

// y = act(x * scale + shift [+ residual])
template <typename T>
__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(const T* __restrict__ x,
                                                              const T* __restrict__ residual, int64_t n,
                                                              int c, const float* __restrict__ scale,
                                                              const float* __restrict__ shift, int relu,
                                                              T* __restrict__ y) {
  constexpr int W = VecT<T>::W;
  const int groups = c / W;
  const int cg = threadIdx.x % groups;
  const int rows_per_block = blockDim.x / groups;
  const int rl = threadIdx.x / groups;
  if (rl >= rows_per_block) return;
  float a[W], b[W];
#pragma unroll
  for (int j = 0; j < W; ++j) {
    a[j] = __ldg(scale + cg * W + j);
    b[j] = __ldg(shift + cg * W + j);
  }
  for (int64_t r = (int64_t)blockIdx.x * rows_per_block + rl; r < n;
       r += (int64_t)gridDim.x * rows_per_block) {
    const int64_t off = r * c + cg * W;
    VecT<T> v, o;
    v.load(x + off);
    if (residual) {
      VecT<T> rs;
      rs.load(residual + off);
#pragma unroll
      for (int j = 0; j < W; ++j) {
        float f = fmaf(v.get(j), a[j], b[j]) + rs.get(j);
        o.set(j, relu ? fmaxf(f, 0.f) : f);
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) {
        float f = fmaf(v.get(j), a[j], b[j]);
        o.set(j, relu ? fmaxf(f, 0.f) : f);
      }
    }
    o.store(y + off);
  }
}

// ReLU mask of the forward pass: relu == 1 reads the saved output (y > 0); relu == 2 (no residual) recomputes
// it from x with the forward's own scale / shift and expression, so the backward pass reads one tensor less
// in each of its two kernels and the autograd graph does not keep y alive for it.
template <typename T>
__device__ __forceinline__ bool relu_open_from_x(float x, float a, float b) {
  const float f = fmaxf(fmaf(x, a, b), 0.f);
  if constexpr (sizeof(T) == 2) return __half2float(__float2half_rn(f)) > 0.f;
  else return f > 0.f;
}

// g = dy * (y > 0 if relu);  sums[0][c] += sum g;  sums[1][c] += sum g * xhat
template <typename T>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_reduce_kernel(
    const T* __restrict__ dy, const T* __restrict__ y, const T* __restrict__ x, int64_t n, int c,
    const float* __restrict__ mean, const float* __restrict__ invstd, int relu,
    const float* __restrict__ scale_shift, double* __restrict__ sums) {
  constexpr int W = VecT<T>::W;
  extern __shared__ float s_part[];
  for (int t = threadIdx.x; t < 2 * c; t += blockDim.x) s_part[t] = 0.f;
  __syncthreads();
  const int groups = c / W;
  const int cg = threadIdx.x % groups;
  const int rows_per_block = blockDim.x / groups;
  const int rl = threadIdx.x / groups;
  if (rl < rows_per_block) {
    float m[W], is[W], sg[W], sgx[W], sa[W], sb[W];
#pragma unroll
    for (int j = 0; j < W; ++j) {
      m[j] = __ldg(mean + cg * W + j);
      is[j] = __ldg(invstd + cg * W + j);
      sa[j] = relu == 2 ? __ldg(scale_shift + cg * W + j) : 0.f;
      sb[j] = relu == 2 ? __ldg(scale_shift + c + cg * W + j) : 0.f;
      sg[j] = sgx[j] = 0.f;
    }
    for (int64_t r = (int64_t)blockIdx.x * rows_per_block + rl; r < n;
         r += (int64_t)gridDim.x * rows_per_block) {
      const int64_t off = r * c + cg * W;
      VecT<T> vd, vx, vy;
      vd.load(dy + off);
      vx.load(x + off);
      if (relu == 1) vy.load(y + off);
#pragma unroll
      for (int j = 0; j < W; ++j) {
        float g = vd.get(j);
        if (relu == 1 && !(vy.get(j) > 0.f)) g = 0.f;
        if (relu == 2 && !relu_open_from_x<T>(vx.get(j), sa[j], sb[j])) g = 0.f;
        sg[j] += g;
        sgx[j] = fmaf(g, (vx.get(j) - m[j]) * is[j], sgx[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < W; ++j) {
      atomicAdd(&s_part[cg * W + j], sg[j]);
      atomicAdd(&s_part[c + cg * W + j], sgx[j]);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * c; t += blockDim.x) atomicAdd(sums + t, (double)s_part[t]);
}

// dx = gamma * invstd * (g - sum_g / n - xhat * sum_gx / n);  dres = g
template <typename T>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_apply_kernel(
    const T* __restrict__ dy, const T* __restrict__ y, const T* __restrict__ x, int64_t n, int c,
    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
    const double* __restrict__ sums, const double* __restrict__ n_dev, int relu,
    const float* __restrict__ scale_shift, T* __restrict__ dx, T* __restrict__ dres) {
  constexpr int W = VecT<T>::W;
  const int groups = c / W;
  const int cg = threadIdx.x % groups;
  const int rows_per_block = blockDim.x / groups;
  const int rl = threadIdx.x / groups;
  if (rl >= rows_per_block) return;
  float m[W], is[W], k0[W], k1[W], k2[W], sa[W], sb[W];
  const float inv_n = n_dev ? (float)(1.0 / *n_dev) : 1.f / (float)n;   // sums are global under sync BN
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const int ch = cg * W + j;
    m[j] = __ldg(mean + ch);
    is[j] = __ldg(invstd + ch);
    const float gi = (gamma ? __ldg(gamma + ch) : 1.f) * is[j];
    sa[j] = relu == 2 ? __ldg(scale_shift + ch) : 0.f;
    sb[j] = relu == 2 ? __ldg(scale_shift + c + ch) : 0.f;
    k0[j] = gi;                                         // * g
    k1[j] = gi * (float)(sums[ch] * (double)inv_n);     // mean of g
    k2[j] = gi * (float)(sums[c + ch] * (double)inv_n); // mean of g * xhat
  }
  for (int64_t r = (int64_t)blockIdx.x * rows_per_block + rl; r < n;
       r += (int64_t)gridDim.x * rows_per_block) {
    const int64_t off = r * c + cg * W;
    VecT<T> vd, vx, vy, o, og;
    vd.load(dy + off);
    vx.load(x + off);
    if (relu == 1) vy.load(y + off);
#pragma unroll
    for (int j = 0; j < W; ++j) {
      float g = vd.get(j);
      if (relu == 1 && !(vy.get(j) > 0.f)) g = 0.f;
      if (relu == 2 && !relu_open_from_x<T>(vx.get(j), sa[j], sb[j])) g = 0.f;
      const float xhat = (vx.get(j) - m[j]) * is[j];
      o.set(j, k0[j] * g - k1[j] - k2[j] * xhat);
      og.set(j, g);
    }
    o.store(dx + off);
    if (dres) og.store(dres + off);
  }
}

static int bn_grid(int64_t n, int c, int w) {
  const int rows_per_block = kBnThreads / (c / w);
  int64_t blocks = ceil_div(n, (int64_t)rows_per_block * 4);   // >= 4 rows per thread
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

static bool bn_shape_ok(int32_t dtype, int c) {
  const int w = dtype == B2S_F16 ? 8 : 4;
  return c % w == 0 && c / w <= kBnThreads && c >= w;
}

}  // namespace b2s

using namespace b2s;

extern "C" {

int b2s_bn_supported(int32_t dtype, int32_t c) {
  return (dtype == B2S_F16 || dtype == B2S_F32) && bn_shape_ok(dtype, c) ? 1 : 0;
}

int b2s_bn_forward(int32_t dtype, const void* x, const void* residual, int64_t n, int32_t c,
                   const float* gamma, const float* beta, float eps, float momentum,
                   float* running_mean, float* running_var, int32_t relu, void* y, float* mean,
                   float* invstd, float* scale_shift /*[2][c]*/, double* sums /*[2][c]*/,
                   b2s_stream_t stream) {
  return b2s_bn_forward_sums(dtype, x, residual, n, c, gamma, beta, eps, momentum, running_mean, running_var,
                             relu, y, mean, invstd, scale_shift, sums, 0, stream);
}

int b2s_bn_forward_sums(int32_t dtype, const void* x, const void* residual, int64_t n, int32_t c,
                        const float* gamma, const float* beta, float eps, float momentum,
                        float* running_mean, float* running_var, int32_t relu, void* y, float* mean,
                        float* invstd, float* scale_shift /*[2][c]*/, double* sums /*[2][c]*/,
                        int32_t sums_ready, b2s_stream_t stream) {
  B2S_REQUIRE(dtype == B2S_F32 || dtype == B2S_F16, B2S_ERR_INVALID, "b2s_bn_forward: dtype");
  B2S_REQUIRE(n >= 1 && c >= 1 && x && y && mean && invstd && scale_shift && sums, B2S_ERR_INVALID,
              "b2s_bn_forward: bad argument");
  B2S_REQUIRE(bn_shape_ok(dtype, c), B2S_ERR_UNSUPPORTED, "b2s_bn_forward: C=%d not a vector multiple", c);
  cudaStream_t st = as_stream(stream);
  const int w = dtype == B2S_F16 ? 8 : 4;
  const int grid = bn_grid(n, c, w);
  const size_t sh = 2 * c * sizeof(float);
  if (!sums_ready) {                     // else: accumulated by the producing conv's epilogue (bn_sums)
    cudaMemsetAsync(sums, 0, 2 * c * sizeof(double), st);
    if (dtype == B2S_F16)
      bn_stats_kernel<__half><<<grid, kBnThreads, sh, st>>>(reinterpret_cast<const __half*>(x), n, c, sums);
    else
      bn_stats_kernel<float><<<grid, kBnThreads, sh, st>>>(reinterpret_cast<const float*>(x), n, c, sums);
  }
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, st>>>(sums, n, sums_ready == 2 ? sums + 2 * c : nullptr, c, gamma,
                                                      beta, eps, momentum, running_mean,
                                                      running_var, mean, invstd, scale_shift,
                                                      scale_shift + c);
  if (dtype == B2S_F16)
    bn_apply_kernel<__half><<<grid, kBnThreads, 0, st>>>(
        reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(residual), n, c, scale_shift,
        scale_shift + c, relu, reinterpret_cast<__half*>(y));
  else
    bn_apply_kernel<float><<<grid, kBnThreads, 0, st>>>(
        reinterpret_cast<const float*>(x), reinterpret_cast<const float*>(residual), n, c, scale_shift,
        scale_shift + c, relu, reinterpret_cast<float*>(y));
  B2S_CHECK_LAUNCH("b2s_bn_forward");
  return B2S_OK;
}

int b2s_bn_stats(int32_t dtype, const void* x, int64_t n, int32_t c, double* sums, b2s_stream_t stream) {
  B2S_REQUIRE((dtype == B2S_F32 || dtype == B2S_F16) && n >= 1 && c >= 1 && x && sums, B2S_ERR_INVALID,
              "b2s_bn_stats: bad argument");
  B2S_REQUIRE(bn_shape_ok(dtype, c), B2S_ERR_UNSUPPORTED, "b2s_bn_stats: C=%d not a vector multiple", c);
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(sums, 0, 2 * c * sizeof(double), st);
  const int grid = bn_grid(n, c, dtype == B2S_F16 ? 8 : 4);
  const size_t sh = 2 * c * sizeof(float);
  if (dtype == B2S_F16)
    bn_stats_kernel<__half><<<grid, kBnThreads, sh, st>>>(reinterpret_cast<const __half*>(x), n, c, sums);
  else
    bn_stats_kernel<float><<<grid, kBnThreads, sh, st>>>(reinterpret_cast<const float*>(x), n, c, sums);
  B2S_CHECK_LAUNCH("b2s_bn_stats");
  return B2S_OK;
}

int b2s_bn_backward_reduce(int32_t dtype, const void* dy, const void* y, const void* x, int64_t n, int32_t c,
                           const float* mean, const float* invstd, int32_t relu, const float* scale_shift,
                           double* sums, b2s_stream_t stream) {
  B2S_REQUIRE(dtype == B2S_F32 || dtype == B2S_F16, B2S_ERR_INVALID, "b2s_bn_backward: dtype");
  B2S_REQUIRE(n >= 1 && c >= 1 && dy && x && mean && invstd && sums && (relu != 1 || y) &&
                  (relu != 2 || scale_shift) && relu >= 0 && relu <= 2, B2S_ERR_INVALID,
              "b2s_bn_backward: bad argument");
  B2S_REQUIRE(bn_shape_ok(dtype, c), B2S_ERR_UNSUPPORTED, "b2s_bn_backward: C=%d not a vector multiple", c);
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(sums, 0, 2 * c * sizeof(double), st);
  const int grid = bn_grid(n, c, dtype == B2S_F16 ? 8 : 4);
  const size_t sh = 2 * c * sizeof(float);
  if (dtype == B2S_F16)
    bn_bwd_reduce_kernel<__half><<<grid, kBnThreads, sh, st>>>(
        reinterpret_cast<const __half*>(dy), reinterpret_cast<const __half*>(y),
        reinterpret_cast<const __half*>(x), n, c, mean, invstd, relu, scale_shift, sums);
  else
    bn_bwd_reduce_kernel<float><<<grid, kBnThreads, sh, st>>>(
        reinterpret_cast<const float*>(dy), reinterpret_cast<const float*>(y),
        reinterpret_cast<const float*>(x), n, c, mean, invstd, relu, scale_shift, sums);
  B2S_CHECK_LAUNCH("b2s_bn_backward_reduce");
  return B2S_OK;
}

int b2s_bn_backward_apply(int32_t dtype, const void* dy, const void* y, const void* x, int64_t n, int32_t c,
                          const float* mean, const float* invstd, const float* gamma, int32_t relu,
                          const float* scale_shift, void* dx, void* dres, const double* sums,
                          const double* n_total, b2s_stream_t stream) {
  B2S_REQUIRE(dtype == B2S_F32 || dtype == B2S_F16, B2S_ERR_INVALID, "b2s_bn_backward: dtype");
  B2S_REQUIRE(n >= 1 && c >= 1 && dy && x && dx && mean && invstd && sums && (relu != 1 || y) &&
                  (relu != 2 || scale_shift) && relu >= 0 && relu <= 2, B2S_ERR_INVALID,
              "b2s_bn_backward: bad argument");
  B2S_REQUIRE(bn_shape_ok(dtype, c), B2S_ERR_UNSUPPORTED, "b2s_bn_backward: C=%d not a vector multiple", c);
  cudaStream_t st = as_stream(stream);
  const int grid = bn_grid(n, c, dtype == B2S_F16 ? 8 : 4);
  if (dtype == B2S_F16)
    bn_bwd_apply_kernel<__half><<<grid, kBnThreads, 0, st>>>(
        reinterpret_cast<const __half*>(dy), reinterpret_cast<const __half*>(y),
        reinterpret_cast<const __half*>(x), n, c, mean, invstd, gamma, sums, n_total, relu, scale_shift,
        reinterpret_cast<__half*>(dx), reinterpret_cast<__half*>(dres));
  else
    bn_bwd_apply_kernel<float><<<grid, kBnThreads, 0, st>>>(
        reinterpret_cast<const float*>(dy), reinterpret_cast<const float*>(y),
        reinterpret_cast<const float*>(x), n, c, mean, invstd, gamma, sums, n_total, relu, scale_shift,
        reinterpret_cast<float*>(dx), reinterpret_cast<float*>(dres));
  B2S_CHECK_LAUNCH("b2s_bn_backward_apply");
  return B2S_OK;
}

int b2s_bn_backward(int32_t dtype, const void* dy, const void* y, const void* x, int64_t n, int32_t c,
                    const float* mean, const float* invstd, const float* gamma, int32_t relu, void* dx,
                    void* dres, double* sums /*[2][c]: d_beta, d_gamma on return*/,
                    b2s_stream_t stream) {
  int rc = b2s_bn_backward_reduce(dtype, dy, y, x, n, c, mean, invstd, relu ? 1 : 0, nullptr, sums, stream);
  if (rc != B2S_OK) return rc;
  return b2s_bn_backward_apply(dtype, dy, y, x, n, c, mean, invstd, gamma, relu ? 1 : 0, nullptr, dx, dres, sums,
                               nullptr, stream);
}

}  // extern "C"
