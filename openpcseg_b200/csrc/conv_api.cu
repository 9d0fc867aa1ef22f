// C-ABI entry points of the sparse convolution; dispatches between the tensor-core
// (conv_tc.cu, fp16, tcgen05) and the CUDA-core (conv_simt.cu, exact fp32) kernel families.
#include <stdlib.h>

#include "common.cuh"

namespace b2s {
template <typename T>
int launch_gather_gemm_simt(const void* in, const void* weight, int k, int c_in, int c_out,
                            int transpose_w, int flip_k, const int32_t* nbr, const int32_t* row_perm,
                            int64_t n_rows, const void* bias, void* out, cudaStream_t st);
template <typename T>
int launch_wgrad_simt(const void* in, const void* gout, const int32_t* nbmaps,
                      const int32_t* nbsizes, int64_t n_identity, int64_t n_pairs_bound, int k,
                      int c_in, int c_out, int swap_pairs, float* gw, cudaStream_t st);

// conv_tc.cu
bool tc_gather_gemm_supported(int c_red, int c_res);
int launch_gather_gemm_tc(const void* in, int64_t n_src, const void* weight, int weight_kmajor, int k, int c_in,
                          int c_out, int transpose_w, int flip_k, const int32_t* nbr, const uint32_t* tile_mask,
                          const int32_t* step_rows, const int32_t* step_start, int tile_rows,
                          const int32_t* row_perm, int64_t n_rows, const void* bias, void* out, double* bn_sums,
                          void* ws, size_t ws_bytes, cudaStream_t st);
size_t tc_gather_gemm_workspace(int k, int c_in, int c_out);
void launch_weight_to_kmajor(const void* w, int k, int c_in, int c_out, void* out, cudaStream_t st);
void launch_weights_refresh(const void* desc, int n, int64_t total_units, cudaStream_t st);
int tc4_tile_rows(int c_res, int64_t n_rows);
bool tc_wgrad_supported(int c_in, int c_out);
int launch_wgrad_tc(const void* in, int64_t n_in, const void* gout, int64_t n_out, const int32_t* nbmaps,
                    const int32_t* nbsizes, int n_seg, int64_t n_identity, int64_t n_pairs_bound, int k,
                    int c_in, int c_out, int swap_pairs, float* gw, cudaStream_t st);

static bool force_simt() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B2S_FORCE_SIMT");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
}  // namespace b2s

using namespace b2s;

extern "C" {

size_t b2s_conv_workspace_bytes(int32_t dtype, int64_t n_rows, int32_t c_in, int32_t c_out,
                                int32_t k) {
  (void)n_rows;
  if (dtype != B2S_F16) return 0;
  return tc_gather_gemm_workspace(k, c_in, c_out);
}

int32_t b2s_conv_steps_supported(int32_t dtype, int64_t n_src, int32_t c_red, int32_t c_res) {
  return dtype == B2S_F16 && !force_simt() && tc_gather_gemm_supported(c_red, c_res) &&
         n_src * (int64_t)c_red * 2 < (int64_t)0xFFFFFF00LL;
}

int32_t b2s_conv_tile_rows(int32_t c_res, int64_t n_rows) { return tc4_tile_rows(c_res, n_rows); }

int b2s_weight_to_kmajor(const void* weight, int32_t k, int32_t c_in, int32_t c_out, void* out,
                         b2s_stream_t stream) {
  B2S_REQUIRE(weight && out && k >= 1 && c_in >= 1 && c_out >= 1, B2S_ERR_INVALID, "b2s_weight_to_kmajor: bad argument");
  launch_weight_to_kmajor(weight, k, c_in, c_out, out, as_stream(stream));
  B2S_CHECK_LAUNCH("b2s_weight_to_kmajor");
  return B2S_OK;
}

int b2s_weights_refresh(const b2s_weight_desc* desc, int32_t n, int64_t total_units, b2s_stream_t stream) {
  B2S_REQUIRE(desc && n >= 1 && total_units >= 1 && total_units < (int64_t)0x7FFFFFFF, B2S_ERR_INVALID,
              "b2s_weights_refresh: bad argument");
  launch_weights_refresh(desc, n, total_units, as_stream(stream));
  B2S_CHECK_LAUNCH("b2s_weights_refresh");
  return B2S_OK;
}

int b2s_conv_gather_gemm(int32_t dtype, const void* in, int64_t n_src, const void* weight,
                         int32_t k, int32_t c_in, int32_t c_out, int32_t transpose_w,
                         int32_t flip_k, const int32_t* nbr, const uint32_t* tile_mask,
                         const int32_t* row_perm, int64_t n_rows, const void* bias, void* out, void* ws,
                         size_t ws_bytes, b2s_stream_t stream) {
  return b2s_conv_gather_gemm_steps(dtype, in, n_src, weight, 0, k, c_in, c_out, transpose_w, flip_k, nbr,
                                    tile_mask, nullptr, nullptr, 0, row_perm, n_rows, bias, out, nullptr, ws,
                                    ws_bytes, stream);
}

int b2s_conv_gather_gemm_steps(int32_t dtype, const void* in, int64_t n_src, const void* weight,
                               int32_t weight_kmajor, int32_t k, int32_t c_in, int32_t c_out,
                               int32_t transpose_w, int32_t flip_k, const int32_t* nbr,
                               const uint32_t* tile_mask, const int32_t* step_rows,
                               const int32_t* step_start, int32_t tile_rows, const int32_t* row_perm,
                               int64_t n_rows, const void* bias, void* out, double* bn_sums, void* ws,
                               size_t ws_bytes, b2s_stream_t stream) {
  B2S_REQUIRE(dtype == B2S_F32 || dtype == B2S_F16, B2S_ERR_INVALID, "b2s_conv_gather_gemm: dtype");
  B2S_REQUIRE(k >= 1 && c_in >= 1 && c_out >= 1 && n_rows >= 0 && n_src >= 0, B2S_ERR_INVALID,
              "b2s_conv_gather_gemm: bad sizes");
  if (n_rows == 0) return B2S_OK;
  B2S_REQUIRE(nbr || step_rows || k == 1, B2S_ERR_INVALID,
              "b2s_conv_gather_gemm: nbr == NULL (identity map) needs k == 1");
  B2S_REQUIRE(nbr || step_rows || !row_perm, B2S_ERR_INVALID, "b2s_conv_gather_gemm: row_perm needs a gather map");
  B2S_REQUIRE(nbr || step_rows || n_rows <= n_src, B2S_ERR_INVALID,
              "b2s_conv_gather_gemm: identity map with n_rows > n_src");
  B2S_REQUIRE(in && weight && out, B2S_ERR_INVALID, "b2s_conv_gather_gemm: null pointer");
  B2S_REQUIRE(n_rows < (1LL << 31) && n_src < (1LL << 31), B2S_ERR_UNSUPPORTED,
              "b2s_conv_gather_gemm: more than 2^31 rows");
  cudaStream_t st = as_stream(stream);
  const int c_red = transpose_w ? c_out : c_in, c_res = transpose_w ? c_in : c_out;
  const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(weight) |
                         reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (dtype == B2S_F16 && !force_simt() && aligned && tc_gather_gemm_supported(c_red, c_res)) {
    int rc = launch_gather_gemm_tc(in, n_src, weight, weight_kmajor, k, c_in, c_out, transpose_w, flip_k, nbr,
                                   tile_mask, step_rows, step_start, tile_rows, row_perm, n_rows, bias, out,
                                   bn_sums, ws, ws_bytes, st);
    if (rc != B2S_OK) return rc;
    B2S_CHECK_LAUNCH("b2s_conv_gather_gemm");
    return B2S_OK;
  }
  // (for the input gradient the "K-major" operand IS the parameter layout the CUDA-core kernels read)
  B2S_REQUIRE(!step_rows && !bn_sums && (!weight_kmajor || transpose_w), B2S_ERR_UNSUPPORTED,
              "b2s_conv_gather_gemm: step tables / bn_sums / K-major weights need the tensor-core kernels "
              "(check b2s_conv_steps_supported)");
  if (dtype == B2S_F16) {
    launch_gather_gemm_simt<__half>(in, weight, k, c_in, c_out, transpose_w, flip_k, nbr, row_perm,
                                    n_rows, bias, out, st);
  } else {
    launch_gather_gemm_simt<float>(in, weight, k, c_in, c_out, transpose_w, flip_k, nbr, row_perm,
                                   n_rows, bias, out, st);
  }
  B2S_CHECK_LAUNCH("b2s_conv_gather_gemm");
  return B2S_OK;
}

int b2s_conv_wgrad(int32_t dtype, const void* in, int64_t n_in, const void* grad_out,
                   int64_t n_out, int32_t k, int32_t c_in, int32_t c_out, const int32_t* nbmaps,
                   const int32_t* nbsizes, int32_t swap_pairs, float* grad_w, void* ws,
                   size_t ws_bytes, b2s_stream_t stream) {
  (void)ws;
  (void)ws_bytes;
  return b2s_conv_wgrad_segments(dtype, in, n_in, grad_out, n_out, k, c_in, c_out, nbmaps, nbsizes, k, swap_pairs,
                                 grad_w, stream);
}

int b2s_conv_wgrad_segments(int32_t dtype, const void* in, int64_t n_in, const void* grad_out,
                            int64_t n_out, int32_t k, int32_t c_in, int32_t c_out, const int32_t* nbmaps,
                            const int32_t* nbsizes, int32_t n_seg, int32_t swap_pairs, float* grad_w,
                            b2s_stream_t stream) {
  B2S_REQUIRE(dtype == B2S_F32 || dtype == B2S_F16, B2S_ERR_INVALID, "b2s_conv_wgrad: dtype");
  B2S_REQUIRE(k >= 1 && c_in >= 1 && c_out >= 1 && n_in >= 0 && n_out >= 0 && grad_w,
              B2S_ERR_INVALID, "b2s_conv_wgrad: bad argument");
  B2S_REQUIRE((nbmaps && nbsizes) || k == 1, B2S_ERR_INVALID,
              "b2s_conv_wgrad: pair list required unless k == 1 (identity)");
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(grad_w, 0, (size_t)k * c_in * c_out * sizeof(float), st);
  if (n_in == 0 || n_out == 0) return B2S_OK;
  B2S_REQUIRE(in && grad_out, B2S_ERR_INVALID, "b2s_conv_wgrad: null pointer");
  // upper bound of the number of pairs (only used to size the grid)
  const int64_t rows = swap_pairs ? n_in : n_out;
  const int64_t bound = nbmaps ? rows * k : (n_in < n_out ? n_in : n_out);
  const int64_t n_identity = n_in < n_out ? n_in : n_out;
  // the tensor-core kernel addresses rows by 32-bit byte offsets
  const bool small = n_in * (int64_t)c_in * 2 < 0xFFFFFF00LL && n_out * (int64_t)c_out * 2 < 0xFFFFFF00LL;
  if (dtype == B2S_F16 && !force_simt() && small && tc_wgrad_supported(c_in, c_out)) {
    int rc = launch_wgrad_tc(in, n_in, grad_out, n_out, nbmaps, nbsizes, nbmaps ? n_seg : k, n_identity, bound, k,
                             c_in, c_out, swap_pairs, grad_w, st);
    if (rc != B2S_OK) return rc;
    B2S_CHECK_LAUNCH("b2s_conv_wgrad");
    return B2S_OK;
  }
  B2S_REQUIRE(n_seg == k, B2S_ERR_UNSUPPORTED,
              "b2s_conv_wgrad_segments: segmented pair lists need the tensor-core kernel");
  if (dtype == B2S_F16) {
    launch_wgrad_simt<__half>(in, grad_out, nbmaps, nbsizes, n_identity, bound, k, c_in, c_out,
                              swap_pairs, grad_w, st);
  } else {
    launch_wgrad_simt<float>(in, grad_out, nbmaps, nbsizes, n_identity, bound, k, c_in, c_out,
                             swap_pairs, grad_w, st);
  }
  B2S_CHECK_LAUNCH("b2s_conv_wgrad");
  return B2S_OK;
}

}  // extern "C"
