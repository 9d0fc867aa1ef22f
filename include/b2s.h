/*
 * b2s.h - C ABI of libb2s.so, the B200-native (sm_100a) sparse-voxel backend.
 *
 * This is the drop-in boundary for the hot path of PJLab-ADG/OpenPCSeg: it
 * replaces the pybind11 module `torchsparse.backend` of the torchsparse 1.4.0
 * bundled in the reference (`package/torchsparse.zip`, written TS/ below =
 * `torchsparse/torchsparse/` inside the zip; the 20 bound functions are listed
 * at TS/backend/pybind_cuda.cpp:18-39) plus RPVNet's `range_utils` ops
 * (pcseg/model/segmentor/fusion/rpvnet/range_lib/range_utils/src/
 * rangelib_bindings_gpu.cpp:7-12).
 *
 * Conventions
 *  - plain pointers + sizes, no torch / ATen types.  Every pointer is a DEVICE
 *    pointer unless the parameter name ends in `_host`.
 *  - every entry point is stream-ordered on `stream` (a cudaStream_t), never
 *    synchronises the device, never allocates: outputs and workspaces are
 *    provided by the caller; `*_bytes` functions size the workspaces.
 *  - return value: 0 = B2S_OK, otherwise a b2s_status; b2s_last_error() gives
 *    a thread-local message.  Nothing throws.
 *  - row-major everywhere.  Coordinates are int32 [N,4] = (x, y, z, batch)
 *    exactly like the reference (TS/tensor.py:10-24).
 *  - `dtype` selects the feature element type: B2S_F32 or B2S_F16.
 */
#ifndef B2S_H_
#define B2S_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b2s_stream_t; /* cudaStream_t */

enum b2s_status {
  B2S_OK = 0,
  B2S_ERR_INVALID = 1,     /* bad argument (shape, null pointer, unsupported dtype) */
  B2S_ERR_WORKSPACE = 2,   /* workspace too small */
  B2S_ERR_CUDA = 3,        /* CUDA runtime error at launch */
  B2S_ERR_UNSUPPORTED = 4  /* shape not supported by this kernel family */
};
enum b2s_dtype { B2S_F32 = 0, B2S_F16 = 1 };

const char* b2s_last_error(void);
int b2s_version(void);
/* Leave `n` SMs free of the persistent convolution grids (default 0, or env B2S_SM_RESERVE): under
 * data-parallel training NCCL's all-reduce CTAs then start at once instead of queueing behind a grid that
 * fills every SM slot for the length of a kernel.  Process-wide, not stream-ordered.                  */
void b2s_set_sm_reserve(int32_t n);

/* ---------------------------------------------------------------- hashing ---
 * replaces hash_cuda / kernel_hash_cuda (TS/backend/hash/hash_cuda.cu:67-84).
 * out[i]       = fold60(fnv1a_words(x, y, z, b))
 * out[k*n + i] = same for (x+off[k][0], y+off[k][1], z+off[k][2], b)          */
int b2s_hash(const int32_t* coords, int64_t n, int64_t* out, b2s_stream_t stream);
int b2s_kernel_hash(const int32_t* coords, int64_t n, const int32_t* offsets, int32_t k,
                    int64_t* out, b2s_stream_t stream);

/* ------------------------------------------------------------- hash table ---
 * replaces hash_query_cuda (TS/backend/others/query_cuda.cu:9-56 and the cuckoo
 * table in TS/backend/hashmap/hashmap_cuda.cu).  Open addressing, 64-bit keys,
 * value = row index of the key in `references`; duplicate keys resolve to the
 * smallest row index (the CPU reference's first-insert-wins,
 * TS/backend/others/query_cpu.cpp:22-26).  The key -1 is reserved (sphash values are < 2^60).
 * out[i] = index or -1 (the value sphashquery returns, TS/nn/functional/query.py:32). */
int64_t b2s_table_slots(int64_t n_references);
size_t b2s_table_bytes(int64_t n_references);
int b2s_table_build(const int64_t* references, int64_t n, void* table, size_t table_bytes,
                    b2s_stream_t stream);
int b2s_table_build_coords(const int32_t* coords, int64_t n, void* table, size_t table_bytes,
                           b2s_stream_t stream); /* keys = b2s_hash(coords), fused */
int b2s_table_query(const void* table, int64_t n_references, const int64_t* queries, int64_t nq,
                    int64_t* out, b2s_stream_t stream);

/* replaces count_cuda (TS/backend/others/count_cuda.cu:25-31); zero-fills out first. */
int b2s_count(const int32_t* idx, int64_t n, int32_t* out, int64_t num, b2s_stream_t stream);

/* -------------------------------------------------- sorted unique / coords ---
 * b2s_unique_i64: ascending unique of int64 keys (what torch.unique(pc_hash)
 * does in pcseg/model/segmentor/voxel/minkunet/utils.py:17).  d_count[0]
 * receives the number of unique keys.
 * b2s_downsample_coords: output coordinates of a strided conv, sorted by
 * (batch, x, y, z) and unique - replaces spdownsample
 * (TS/nn/functional/downsample.py:11-52), both the snap-to-grid fast path and
 * the kernel-expansion slow path.  `out_coords` must hold n rows (fast path) or
 * n*kernel_volume rows (slow path); d_count[0] = rows written, d_count[1] = 0 or
 * 1 if a coordinate fell outside the packable range (|xyz| < 2^17, 0 <= b < 1024). */
size_t b2s_unique_workspace_bytes(int64_t n);
int b2s_unique_i64(const int64_t* keys, int64_t n, int64_t* out, int64_t* d_count, void* ws,
                   size_t ws_bytes, b2s_stream_t stream);
int64_t b2s_downsample_capacity(int64_t n, const int32_t* stride_host,
                                const int32_t* kernel_host);
size_t b2s_downsample_workspace_bytes(int64_t n, const int32_t* stride_host,
                                      const int32_t* kernel_host);
int b2s_downsample_coords(const int32_t* coords, int64_t n, const int32_t* stride_host,
                          const int32_t* kernel_host, const int32_t* tensor_stride_host,
                          int32_t* out_coords, int64_t* d_count, void* ws, size_t ws_bytes,
                          b2s_stream_t stream);

/* ------------------------------------------------------------- kernel map ---
 * Fused replacement of sphash + kernel-hash + sphashquery + nonzero
 * (TS/nn/functional/conv.py:156-176).  Pair convention of the reference:
 * in_coord = out_coord + offset[k].
 *   nbr_out [K, n_out] int32 : input row feeding output row o through W[k], or -1
 *   nbr_in  [K, n_in ] int32 : output row fed by input row i through W[k], or -1
 *                              (may be NULL)
 *   nbsizes [K] int32        : pairs per offset
 *   tile_mask_out / tile_mask_in (optional, may be NULL): uint32 [ceil(n/128)][ceil(K/32)],
 *                              bit k of tile t set <=> some row of the 128-row tile t of
 *                              nbr_out / nbr_in has a neighbour for offset k (lets the
 *                              convolution skip whole (tile, offset) steps without scanning)
 * b2s_kmap_pairs then emits the reference-format pair list: nbmaps int32 [M,2] =
 * (in, out), grouped by k ascending, out ascending inside a group; d_total[0]=M.
 * `nbmaps` must hold K*n_out rows.                                            */
size_t b2s_kmap_workspace_bytes(int64_t n_in, int64_t n_out, int32_t k);
int b2s_kmap_build(const int32_t* in_coords, int64_t n_in, const int32_t* out_coords,
                   int64_t n_out, const int32_t* offsets, int32_t k, int32_t* nbr_out,
                   int32_t* nbr_in, int32_t* nbsizes, uint32_t* tile_mask_out,
                   uint32_t* tile_mask_in, void* ws, size_t ws_bytes, b2s_stream_t stream);
int b2s_kmap_pairs(const int32_t* nbr_out, int32_t k, int64_t n_out, int32_t* nbmaps,
                   int64_t* d_total, void* ws, size_t ws_bytes, b2s_stream_t stream);
/* The same pair list in (chunk, offset, row) order for the weight gradient: the n_out rows - taken in the order
 * perm (NULL: as stored) - are cut into n_chunks equal ranges and all K offsets of a range are adjacent, so the
 * X / dY rows of a range are reused from L2 across its offsets instead of being re-read from HBM once per offset
 * (with perm = a spatial order of the rows also across neighbouring rows).  seg_sizes int32 [n_chunks * K]
 * (segment c*K + k = pairs of offset k in range c) replaces nbsizes in b2s_conv_wgrad_segments.          */
size_t b2s_kmap_pairs_chunked_workspace_bytes(int64_t n_out, int32_t k, int32_t n_chunks);
int b2s_kmap_pairs_chunked(const int32_t* nbr_out, int32_t k, int64_t n_out, const int32_t* perm,
                           int32_t n_chunks, int32_t* nbmaps, int32_t* seg_sizes, int64_t* d_total, void* ws,
                           size_t ws_bytes, b2s_stream_t stream);
/* active-offset masks of the 128-row tiles of an arbitrary gather map nbr [K, n] (e.g. a
 * column-permuted copy of nbr_out): tile_mask uint32 [ceil(n/128)][ceil(K/32)].           */
int b2s_tile_mask(const int32_t* nbr, int32_t k, int64_t n, uint32_t* tile_mask, b2s_stream_t stream);
/* int64 sort key per row of a submanifold map (k <= 27): neighbourhood bit pattern, rarest offset
 * most significant, then a coarse (z, x, y) code of coords >> coord_shift.  Sorting rows by it and
 * passing the order as row_perm groups rows with equal patterns into the same 128-row tiles.      */
int b2s_tile_order_key(const int32_t* nbr, int32_t k, int64_t n, const int32_t* nbsizes,
                       const int32_t* coords, int32_t coord_shift, int64_t* keys, b2s_stream_t stream);
/* Same, and also emits row_bits uint32 [n]: bit kk of row r <=> nbr[kk][r] >= 0 (k <= 32), which
 * b2s_tile_steps uses to build tile masks with one read per row.                                     */
int b2s_tile_order_key_bits(const int32_t* nbr, int32_t k, int64_t n, const int32_t* nbsizes,
                            const int32_t* coords, int32_t coord_shift, int64_t* keys, uint32_t* row_bits,
                            b2s_stream_t stream);
/* Step table of a gather map for the tensor-core convolution (conv_tc4.cu).  Launch row j (j < n) is map
 * column perm[j] (perm == NULL: j); tiles are `tile_rows` (128 | 256) consecutive launch rows.  Outputs:
 *   tile_mask  uint32 [tiles][ceil(k/32)]  active offsets of a tile;
 *   step_start int32 [tiles + 1]           exclusive prefix of the per-tile active-offset counts;
 *   step_rows  int32 [<= k * tiles * tile_rows]  for step s = step_start[t] + (rank of offset kk among the
 *              active offsets of tile t), tile row w*32 + i*4 + q is stored at s*tile_rows + w*32 + q*8 + i
 *              (= nbr[kk][launch row] or -1): the eight rows one gather lane copies are contiguous.
 * Everything stays on the device (no size leaves it); row_bits (optional, k <= 32) as emitted by
 * b2s_tile_order_key_bits.  The capacity k*tiles*tile_rows of step_rows is the caller's to provide.   */
int b2s_tile_steps(const int32_t* nbr, int32_t k, int64_t n, const int32_t* perm, const uint32_t* row_bits,
                   int32_t tile_rows, uint32_t* tile_mask, int32_t* step_start, int32_t* step_rows,
                   b2s_stream_t stream);

/* ------------------------------------------------------------ convolution ---
 * replaces convolution_forward_cuda / convolution_backward_cuda
 * (TS/backend/convolution/convolution_cuda.cu:53-165, :167-278).
 *
 * b2s_conv_gather_gemm: out[r, :] = sum_k in[nbr[kk(k)][r], :] * B_k   for r < n_rows
 *   where nbr is [K, n_rows] (rows with -1 contribute nothing), kk(k) = flip_k ?
 *   K-1-k : k, and B_k = weight[k] ([c_in, c_out]) when transpose_w == 0
 *   (forward: c_red = c_in, c_res = c_out) or weight[k]^T when transpose_w == 1
 *   (input gradient: c_red = c_out, c_res = c_in).  fp32 accumulation over all
 *   offsets, one write per output row, no atomics, deterministic.
 *   `in` has n_src rows of c_red channels; `out` n_rows x c_res; optional bias[c_res];
 *   optional tile_mask = the b2s_kmap_build / b2s_tile_mask mask that belongs to `nbr` (NULL:
 *   scanned); optional row_perm int32 [n_rows]: result row j of the launch is written to
 *   out[row_perm[j]] - lets the caller group rows with similar neighbourhoods into the same
 *   128-row tile (column j of `nbr` then describes original row row_perm[j]).
 * b2s_conv_wgrad: grad_w[k] = sum over pairs of offset k of in[i]^T * grad_out[o],
 *   pairs from b2s_kmap_pairs (device-resident sizes, no host sync).  grad_w is
 *   fp32 [K, c_in, c_out] and is zero-filled by the call.
 * `weight` has the feature dtype; k = 1 expresses the 1x1 / dense case with
 * nbr == NULL (identity map).                                                  */
size_t b2s_conv_workspace_bytes(int32_t dtype, int64_t n_rows, int32_t c_in, int32_t c_out,
                                int32_t k);
int b2s_conv_gather_gemm(int32_t dtype, const void* in, int64_t n_src, const void* weight,
                         int32_t k, int32_t c_in, int32_t c_out, int32_t transpose_w,
                         int32_t flip_k, const int32_t* nbr, const uint32_t* tile_mask,
                         const int32_t* row_perm, int64_t n_rows, const void* bias, void* out, void* ws,
                         size_t ws_bytes, b2s_stream_t stream);
/* The same contraction driven by a STEP TABLE (b2s_tile_steps) instead of the [K, n] gather map - the
 * production path of the fp16 tensor-core kernel (conv_tc4.cu): the table lists only the active (tile, offset)
 * steps, with the source rows in the order the gather lanes consume them, and is shared by every convolution
 * that uses the kernel map.  Extras: weight_kmajor != 0 - `weight` already is this pass's K-major operand
 * [K][c_res][c_red] (b2s_weight_to_kmajor output for the forward pass; the input gradient's is the parameter
 * layout itself), so no per-call transpose and no workspace; bn_sums (fp64 [2][c_res], caller-zeroed) receives
 * += per-channel sum and sum of squares of the fp16 rows written (batch-norm statistics from the epilogue).
 * With step_rows == NULL it behaves exactly like b2s_conv_gather_gemm.
 * b2s_conv_steps_supported: whether (dtype, sizes) can take a step table; b2s_conv_tile_rows: the tile_rows
 * (128 or 256) the kernel wants for this result width - the table must be built with it.              */
int32_t b2s_conv_steps_supported(int32_t dtype, int64_t n_src, int32_t c_red, int32_t c_res);
int32_t b2s_conv_tile_rows(int32_t c_res, int64_t n_rows);
int b2s_weight_to_kmajor(const void* weight_f16, int32_t k, int32_t c_in, int32_t c_out, void* out_f16,
                         b2s_stream_t stream);
/* The fp16 operand copies of MANY fp32 master weights in one launch - what the reference does per conv call through
 * custom_fwd(cast_inputs=torch.half) (TS/nn/functional/conv.py:19), done once per optimizer update for every
 * registered parameter.  desc: DEVICE array of n descriptors; entry i owns the 32 x 32 tiles
 * [unit_start, unit_start + k * ceil(c_in / 32) * ceil(c_out / 32)) of the launch (exclusive prefix sums, ascending);
 * total_units = their sum.  cast_f16 / kmajor_f16 may be NULL.                                           */
typedef struct b2s_weight_desc {
  const float* src;     /* fp32 [k][c_in][c_out]                                        */
  void* cast_f16;       /* fp16 [k][c_in][c_out]: operand of the input-gradient pass     */
  void* kmajor_f16;     /* fp16 [k][c_out][c_in]: operand of the forward pass            */
  int32_t k, c_in, c_out, unit_start;
} b2s_weight_desc;
int b2s_weights_refresh(const b2s_weight_desc* desc, int32_t n, int64_t total_units, b2s_stream_t stream);
int b2s_conv_gather_gemm_steps(int32_t dtype, const void* in, int64_t n_src, const void* weight,
                               int32_t weight_kmajor, int32_t k, int32_t c_in, int32_t c_out,
                               int32_t transpose_w, int32_t flip_k, const int32_t* nbr,
                               const uint32_t* tile_mask, const int32_t* step_rows,
                               const int32_t* step_start, int32_t tile_rows, const int32_t* row_perm,
                               int64_t n_rows, const void* bias, void* out, double* bn_sums, void* ws,
                               size_t ws_bytes, b2s_stream_t stream);
int b2s_conv_wgrad(int32_t dtype, const void* in, int64_t n_in, const void* grad_out,
                   int64_t n_out, int32_t k, int32_t c_in, int32_t c_out, const int32_t* nbmaps,
                   const int32_t* nbsizes, int32_t swap_pairs, float* grad_w, void* ws,
                   size_t ws_bytes, b2s_stream_t stream);
/* b2s_conv_wgrad over a segmented pair list (b2s_kmap_pairs_chunked): n_seg = n_chunks * K consecutive segments
 * of seg_sizes[s] pairs, segment s belonging to offset s % K.  n_seg <= 1024.                           */
int b2s_conv_wgrad_segments(int32_t dtype, const void* in, int64_t n_in, const void* grad_out,
                            int64_t n_out, int32_t k, int32_t c_in, int32_t c_out, const int32_t* nbmaps,
                            const int32_t* seg_sizes, int32_t n_seg, int32_t swap_pairs, float* grad_w,
                            b2s_stream_t stream);

/* --------------------------------------------------------- point <-> voxel ---
 * replaces voxelize_{forward,backward}_cuda (TS/backend/voxelize/voxelize_cuda.cu:44-80)
 * and devoxelize_{forward,backward}_cuda (TS/backend/devoxelize/devoxelize_cuda.cu:61-98).
 * `acc` is an fp32 scratch [n_vox, c] needed when dtype == B2S_F16 for the
 * scatter-adds (NULL for B2S_F32: the output itself is the accumulator).       */
int b2s_voxelize_fwd(int32_t dtype, const void* feats, const int32_t* idx, const int32_t* counts,
                     int64_t n_pts, int64_t n_vox, int32_t c, void* out, float* acc,
                     b2s_stream_t stream);
int b2s_voxelize_bwd(int32_t dtype, const void* grad_vox, const int32_t* idx,
                     const int32_t* counts, int64_t n_pts, int64_t n_vox, int32_t c,
                     void* grad_pts, b2s_stream_t stream);
int b2s_devoxelize_fwd(int32_t dtype, const void* feats, const int32_t* idx /*[n_pts,8]*/,
                       const float* weights /*[n_pts,8] fp32*/, int64_t n_pts,
                       int64_t n_vox, int32_t c, void* out, b2s_stream_t stream);
int b2s_devoxelize_bwd(int32_t dtype, const void* grad_pts, const int32_t* idx,
                       const float* weights, int64_t n_pts, int64_t n_vox, int32_t c,
                       void* grad_vox, float* acc, b2s_stream_t stream);
/* b2s_devoxelize_bwd for contended maps (many points per voxel: coarse strides).  `order` int32 [n_pts] is a
 * permutation of the points sorted by their corner-0 voxel (idx[:, 0]); consecutive points then share corners
 * and their contributions are summed in registers before ONE vector red per (corner, run).  Same result up to
 * fp32 summation order.  Needs C to be a multiple of the 16-byte vector (8 fp16 / 4 fp32 channels).        */
int b2s_devoxelize_bwd_sorted(int32_t dtype, const void* grad_pts, const int32_t* order, const int32_t* idx,
                              const float* weights, int64_t n_pts, int64_t n_vox, int32_t c, void* grad_vox,
                              float* acc, b2s_stream_t stream);

/* Scatter-max of point rows into voxel rows (Cylinder3D's torch_scatter.scatter_max,
 * tools/utils/common/seg_utils.py:172-188, pcseg/model/segmentor/voxel/cylinder3d/
 * cylinder_ts.py:24-43): out[idx[i], j] = max over i of feats[i, j]; arg int64 [m, c] = the
 * smallest contributing row (n for empty voxels, whose out is 0) - what backward needs.
 * keys uint32 [m, c] is scratch.                                                             */
int b2s_scatter_max(int32_t dtype, const void* feats, const int64_t* idx, int64_t n, int32_t c, int64_t m,
                    void* out, int64_t* arg, uint32_t* keys, b2s_stream_t stream);

/* Fused voxel_to_point map (pcseg/model/segmentor/voxel/minkunet/utils.py:73-81 +
 * calc_ti_weights, TS/nn/functional/devoxelize.py:10-48): for every point, the 8
 * corner voxel rows (or -1) at `stride` and the fp32 trilinear weights.
 * pts: fp32 [n_pts,4] = (x, y, z, batch) in voxel units; table built over the
 * voxel coords with b2s_table_build_coords.  idx int32 [n_pts,8], w fp32 [n_pts,8]. */
int b2s_trilinear_map(const float* pts, int64_t n_pts, int32_t stride, const void* table,
                      int64_t n_vox, int32_t* idx, float* w, b2s_stream_t stream);
/* calc_ti_weights alone: idx is the reference-layout int64 [8, n_pts]; w fp32 [8, n_pts]. */
int b2s_ti_weights(const float* pts, int64_t n_pts, const int64_t* idx, float scale, float* w,
                   b2s_stream_t stream);

/* ------------------------------------------------- fused batch norm ("next" N1) ---
 * Training-mode BatchNorm1d over the rows of [n, c] voxel features, fused with an optional
 * residual add and ReLU:  y = act(bn(x) [+ residual])  - what the voxel segmentors apply
 * after every sparse conv (nn.BatchNorm1d through fapply, minkunet.py:27-29; add + ReLU of
 * the residual block, minkunet.py:134-136).  gamma/beta/running_* are fp32 [c] (running_*
 * may be NULL); mean/invstd fp32 [c] are saved for backward; scale_shift fp32 [2][c] and
 * sums fp64 [2][c] are scratch.  Backward: dy, y (needed when relu), x -> dx (and dres = the
 * gradient of the residual input, may be NULL); on return sums[0][c] = d_beta, sums[1][c] =
 * d_gamma.  c must be a multiple of the 16-byte vector (8 for fp16, 4 for fp32).            */
int b2s_bn_supported(int32_t dtype, int32_t c);
int b2s_bn_forward(int32_t dtype, const void* x, const void* residual, int64_t n, int32_t c,
                   const float* gamma, const float* beta, float eps, float momentum,
                   float* running_mean, float* running_var, int32_t relu, void* y, float* mean,
                   float* invstd, float* scale_shift, double* sums, b2s_stream_t stream);
/* b2s_bn_forward with sums_ready != 0: `sums` already holds the per-channel sum / sum of squares of x (the
 * bn_sums output of b2s_conv_gather_gemm_steps) - the statistics pass over [n, c] is skipped.          */
int b2s_bn_forward_sums(int32_t dtype, const void* x, const void* residual, int64_t n, int32_t c,
                        const float* gamma, const float* beta, float eps, float momentum,
                        float* running_mean, float* running_var, int32_t relu, void* y, float* mean,
                        float* invstd, float* scale_shift, double* sums, int32_t sums_ready,
                        b2s_stream_t stream);
/* Pieces of the two calls above for SYNCHRONISED batch norm (the reference's IF_DIST: True): the caller
 * all-reduces the fp64 sums between them.  b2s_bn_stats: sums[2][c] = per-channel (sum, sum of squares) of x.
 * b2s_bn_forward_sums with sums_ready == 2: sums has 2c + 1 entries, the last one the GLOBAL row count.
 * b2s_bn_backward_reduce: sums[2][c] = (sum dy', sum dy' * xhat) of the local rows (= d_beta, d_gamma);
 * b2s_bn_backward_apply: dx (and dres) from all-reduced sums and the global count n_total (device fp64; NULL:
 * the local n).                                                                                          */
int b2s_bn_stats(int32_t dtype, const void* x, int64_t n, int32_t c, double* sums, b2s_stream_t stream);
/* relu: 0 none, 1 mask from the saved output y, 2 (no residual) mask recomputed from x with the forward's
 * scale_shift fp32 [2][c] (b2s_bn_forward's output) - y may then be NULL and is not read.             */
int b2s_bn_backward_reduce(int32_t dtype, const void* dy, const void* y, const void* x, int64_t n, int32_t c,
                           const float* mean, const float* invstd, int32_t relu, const float* scale_shift,
                           double* sums, b2s_stream_t stream);
int b2s_bn_backward_apply(int32_t dtype, const void* dy, const void* y, const void* x, int64_t n, int32_t c,
                          const float* mean, const float* invstd, const float* gamma, int32_t relu,
                          const float* scale_shift, void* dx, void* dres, const double* sums,
                          const double* n_total, b2s_stream_t stream);
int b2s_bn_backward(int32_t dtype, const void* dy, const void* y, const void* x, int64_t n, int32_t c,
                    const float* mean, const float* invstd, const float* gamma, int32_t relu, void* dx,
                    void* dres, double* sums, b2s_stream_t stream);

/* --------------------------------------------------------- range-image ops ---
 * RPVNet: replaces map_count_forward / denselize_forward / denselize_backward
 * (range_lib/range_utils/src/map_count_gpu.cu:5-14, denselize_gpu.cu:5-34).
 * pxpy int32 [n,3] = (batch, px, py); count_map int32 [B,1,H,W]; dense fp32 [B,C,H,W]. */
int b2s_map_count(const int32_t* pxpy, int64_t n, int32_t b, int32_t h, int32_t w,
                  int32_t* count_map, b2s_stream_t stream);
int b2s_denselize_fwd(const float* feats, const int32_t* pxpy, const int32_t* count_map,
                      int64_t n, int32_t c, int32_t b, int32_t h, int32_t w, float* dense,
                      b2s_stream_t stream);
int b2s_denselize_bwd(const float* grad_dense, const int32_t* pxpy, const int32_t* count_map,
                      int64_t n, int32_t c, int32_t b, int32_t h, int32_t w, float* grad_feats,
                      b2s_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B2S_H_ */
