// Sparse convolution, CUDA-core (SIMT) kernel family: exact fp32 FMA accumulation.
//
// This is the fp32 path (BASELINE tolerance 1e-5 rules out tf32 tensor cores) and the
// path for channel counts the tcgen05 kernels do not tile (C not a multiple of 16, e.g.
// the 4-channel stem).  Same output-stationary formulation as the tensor-core family
// (conv_tc.cu): each CTA owns 64 output rows, loops over the K kernel offsets, gathers the
// neighbour rows named by nbr[k][row] into shared memory and accumulates all offsets in
// registers -> one write per output element, no atomics, deterministic.
//
// Roofline: fp32 FMA pipe for wide layers, L2 gather bandwidth for narrow ones.
// Algorithmic bytes per launch (e = element size): e*C_red*M (gathered rows) +
// e*C_res*N_rows + 4*K*N_rows (map) + e*K*C_in*C_out.
#include "common.cuh"

namespace b2s {

constexpr int BM = 64, BN = 64, BK = 16;

template <typename T>
__global__ void __launch_bounds__(256) gather_gemm_simt_kernel(
    const T* __restrict__ in, const T* __restrict__ weight, const int32_t* __restrict__ nbr,
    const int32_t* __restrict__ row_perm, const T* __restrict__ bias, T* __restrict__ out, int64_t n_rows,
    int kvol, int c_in, int c_out, int transpose_w, int flip_k) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ int32_t s_row[BM];
  const int c_red = transpose_w ? c_out : c_in;
  const int c_res = transpose_w ? c_in : c_out;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int col0 = blockIdx.y * BN;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  for (int k = 0; k < kvol; ++k) {
    int have = 0;
    if (tid < BM) {
      int64_t r = row0 + tid;
      int32_t src = -1;
      if (r < n_rows) src = nbr ? __ldg(nbr + (int64_t)(flip_k ? kvol - 1 - k : k) * n_rows + r)
                                : (int32_t)r;
      s_row[tid] = src;
      have = src >= 0;
    }
    if (!__syncthreads_or(have)) continue;  // no row of this tile has neighbour k
    const T* wk = weight + (int64_t)k * c_in * c_out;
    for (int c0 = 0; c0 < c_red; c0 += BK) {
      {  // A tile: 64 rows x 16 channels, 4 per thread
        const int r = tid >> 2, kk0 = (tid & 3) * 4;
        const int32_t src = s_row[r];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int c = c0 + kk0 + j;
          As[kk0 + j][r] = (src >= 0 && c < c_red) ? FeatIO<T>::load(in + (int64_t)src * c_red + c) : 0.f;
        }
      }
      if (!transpose_w) {  // B[c][n] = W[k][c][n], n contiguous
        const int kk = tid >> 4, n0 = (tid & 15) * 4;
        const int c = c0 + kk;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int n = col0 + n0 + j;
          Bs[kk][n0 + j] = (c < c_red && n < c_res) ? FeatIO<T>::load(wk + (int64_t)c * c_out + n) : 0.f;
        }
      } else {  // B[c][n] = W[k][n][c], c contiguous
        const int n = tid >> 2, kk0 = (tid & 3) * 4;
        const int nn = col0 + n;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int c = c0 + kk0 + j;
          Bs[kk0 + j][n] = (c < c_red && nn < c_res) ? FeatIO<T>::load(wk + (int64_t)nn * c_out + c) : 0.f;
        }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[4], b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = As[kk][ty * 4 + j];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t r = row0 + ty * 4 + i;
    if (r >= n_rows) continue;
    const int64_t r_out = row_perm ? (int64_t)__ldg(row_perm + r) : r;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = col0 + tx * 4 + j;
      if (n < c_res) {
        float v = acc[i][j];
        if (bias) v += FeatIO<T>::load(bias + n);
        FeatIO<T>::store(out + r_out * c_res + n, v);
      }
    }
  }
}

// grad_w[k][ci][co] += sum over the pairs of offset k of in[i][ci] * grad_out[o][co].
// grid = (splits, K, ci-tiles * co-tiles); pair ranges come from the device-resident
// nbsizes, so no host synchronisation is needed.
template <typename T>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(
    const T* __restrict__ in, const T* __restrict__ gout, const int32_t* __restrict__ nbmaps,
    const int32_t* __restrict__ nbsizes, int64_t n_identity, int kvol, int c_in, int c_out,
    int swap_pairs, float* __restrict__ gw) {
  __shared__ float Xs[BK][BM + 4];
  __shared__ float Ys[BK][BN + 4];
  __shared__ int32_t s_i[BK], s_o[BK];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int k = blockIdx.y;
  const int co_tiles = (c_out + BN - 1) / BN;
  const int ci0 = (blockIdx.z / co_tiles) * BM, co0 = (blockIdx.z % co_tiles) * BN;
  int64_t start = 0, cnt;
  if (nbmaps) {
    for (int j = 0; j < k; ++j) start += __ldg(nbsizes + j);
    cnt = __ldg(nbsizes + k);
  } else {
    cnt = n_identity;
  }
  const int64_t per = (cnt + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * per;
  const int64_t hi = lo + per < cnt ? lo + per : cnt;
  if (lo >= hi) return;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (int64_t p0 = lo; p0 < hi; p0 += BK) {
    if (tid < BK) {
      int64_t p = p0 + tid;
      int32_t i = -1, o = -1;
      if (p < hi) {
        if (nbmaps) {
          int2 pr = __ldg(reinterpret_cast<const int2*>(nbmaps) + start + p);
          i = swap_pairs ? pr.y : pr.x;
          o = swap_pairs ? pr.x : pr.y;
        } else {
          i = o = (int32_t)p;
        }
      }
      s_i[tid] = i;
      s_o[tid] = o;
    }
    __syncthreads();
    {
      const int pp = tid >> 4, c4 = (tid & 15) * 4;
      const int32_t i = s_i[pp], o = s_o[pp];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int ci = ci0 + c4 + j, co = co0 + c4 + j;
        Xs[pp][c4 + j] = (i >= 0 && ci < c_in) ? FeatIO<T>::load(in + (int64_t)i * c_in + ci) : 0.f;
        Ys[pp][c4 + j] = (o >= 0 && co < c_out) ? FeatIO<T>::load(gout + (int64_t)o * c_out + co) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int pp = 0; pp < BK; ++pp) {
      float a[4], b[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) a[j] = Xs[pp][ty * 4 + j];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ys[pp][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* dst = gw + (int64_t)k * c_in * c_out;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int ci = ci0 + ty * 4 + i;
    if (ci >= c_in) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int co = co0 + tx * 4 + j;
      if (co < c_out) atomicAdd(dst + (int64_t)ci * c_out + co, acc[i][j]);
    }
  }
}

template <typename T>
int launch_gather_gemm_simt(const void* in, const void* weight, int k, int c_in, int c_out,
                            int transpose_w, int flip_k, const int32_t* nbr, const int32_t* row_perm,
                            int64_t n_rows, const void* bias, void* out, cudaStream_t st) {
  const int c_res = transpose_w ? c_in : c_out;
  dim3 grid((unsigned)ceil_div(n_rows, BM), (unsigned)ceil_div(c_res, BN));
  gather_gemm_simt_kernel<T><<<grid, 256, 0, st>>>(
      reinterpret_cast<const T*>(in), reinterpret_cast<const T*>(weight), nbr, row_perm,
      reinterpret_cast<const T*>(bias), reinterpret_cast<T*>(out), n_rows, k, c_in, c_out,
      transpose_w, flip_k);
  return 0;
}

template <typename T>
int launch_wgrad_simt(const void* in, const void* gout, const int32_t* nbmaps,
                      const int32_t* nbsizes, int64_t n_identity, int64_t n_pairs_bound, int k,
                      int c_in, int c_out, int swap_pairs, float* gw, cudaStream_t st) {
  int tiles = (int)(ceil_div(c_in, BM) * ceil_div(c_out, BN));
  // enough splits to fill the machine a few times over, capped by the work available
  int64_t want = ceil_div((int64_t)sm_count() * 4, (int64_t)k * tiles);
  int64_t by_work = ceil_div(n_pairs_bound, (int64_t)k * 256);
  int splits = (int)(want < by_work ? want : by_work);
  if (splits < 1) splits = 1;
  dim3 grid((unsigned)splits, (unsigned)k, (unsigned)tiles);
  wgrad_simt_kernel<T><<<grid, 256, 0, st>>>(reinterpret_cast<const T*>(in),
                                            reinterpret_cast<const T*>(gout), nbmaps, nbsizes,
                                            n_identity, k, c_in, c_out, swap_pairs, gw);
  return 0;
}

template int launch_gather_gemm_simt<float>(const void*, const void*, int, int, int, int, int,
                                            const int32_t*, const int32_t*, int64_t, const void*,
                                            void*, cudaStream_t);
template int launch_gather_gemm_simt<__half>(const void*, const void*, int, int, int, int, int,
                                             const int32_t*, const int32_t*, int64_t, const void*,
                                             void*, cudaStream_t);
template int launch_wgrad_simt<float>(const void*, const void*, const int32_t*, const int32_t*,
                                      int64_t, int64_t, int, int, int, int, float*, cudaStream_t);
template int launch_wgrad_simt<__half>(const void*, const void*, const int32_t*, const int32_t*,
                                       int64_t, int64_t, int, int, int, int, float*, cudaStream_t);

}  // namespace b2s
