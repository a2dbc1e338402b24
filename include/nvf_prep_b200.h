/*
 * nvf_prep_b200.h - C ABI of the components either side of the NVF decoder hot
 * path (SURVEY.md section 8f "next rows"), exported by the same shared library
 * as include/nvf_b200.h (libnvf_b200.so) and following the same conventions:
 * plain C, device pointers unless the name ends in _host, caller-owned memory
 * and workspace, asynchronous on `stream`, negative error codes decoded by
 * nvf_strerror(), nothing throws or exits across the boundary.
 */
#ifndef NVF_PREP_B200_H_
#define NVF_PREP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/*
 * Ground-truth grid + distance field builder.  Replaces util_get_grids.py:19-46
 * (cube_template + origin -> open3d KDTreeFlann.search_knn_vector_3d per grid
 * voxel -> dist = ||nearest - p|| as float64, gt_grid = (dist == 0) as uint8).
 *
 *   points     [n_points,3] int32 voxel coordinates of the cloud (pcd.points)
 *   origins    [n_blocks,3] int32 leaf origins ({fid}_l5_origins.txt); need not be
 *              multiples of 32
 *   max_cells  capacity of the internal sparse occupancy structure: an upper bound
 *              on the number of distinct (point >> 5) cells (n_blocks for aligned
 *              octree leaves, 8 * n_blocks always suffices when every point lies in a leaf)
 *   max_radius search radius in voxels per axis; 53 (= floor(31*sqrt(3)), the farthest a
 *              nearest point of a NON-EMPTY leaf can be) makes the result exact.  Values
 *              above 53 are clamped.
 *   gt_out     [n_blocks,1,32,32,32] uint8 or NULL
 *   dist64_out [n_blocks,1,32,32,32] float64 or NULL   (the {fid}_l5_dist.npy payload)
 *   dist32_out [n_blocks,1,32,32,32] float32 or NULL   (= dist64.float(), what
 *              LoadedVoxelDataset.__getitem__ feeds the loss, utils/dataloader.py:170-171)
 *   d2_out     [n_blocks,32768] uint16 squared distances or NULL (0xFFFF = none found)
 *   status_out [1] int32 device word, OR of NVF_GRID_STATUS_* (0 = exact result)
 *   workspace  nvf_grids_workspace_bytes(max_cells) bytes
 */
enum {
  NVF_GRID_STATUS_CELL_OVERFLOW = 1, /* more occupied cells than max_cells (or coordinates beyond +-2^25) */
  NVF_GRID_STATUS_NOT_FOUND = 2      /* some voxel has no point within max_radius (empty leaf)            */
};
#define NVF_GRID_MAX_RADIUS 53
int nvf_grids_workspace_bytes(int64_t max_cells, size_t* bytes_out);
int nvf_build_grids(const int32_t* points, int64_t n_points, const int32_t* origins, int64_t n_blocks,
                    int64_t max_cells, int32_t max_radius, uint8_t* gt_out, double* dist64_out, float* dist32_out,
                    uint16_t* d2_out, int32_t* status_out, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Minibatch gather for device-resident datasets.  Replaces the DataLoader batch assembly +
 * `.to(device)` + `emb[indices]` of the weight loop (LoadedVoxelDataset.__getitem__,
 * utils/dataloader.py:163-172; NVFPCC.py:151-158) when gt / dist / emb already live in HBM:
 * ONE launch copies the rows `idx` of the three tensors into the step's static input buffers.
 *   emb_all [n_rows, emb_floats], gt_all / dist_all [n_rows, 32768] float32, idx [n] int64 (device),
 *   emb_out [n, emb_floats], gt_out / dist_out [n, 32768].
 *   status  optional device int32 (sticky): bit 0 is set when an index lies outside [0, n_rows) - that row is
 *           then read from row 0 instead of out of bounds (the indices live on the device, so the host cannot
 *           check them without a synchronisation; callers read the word when they read their results).
 */
int nvf_gather_batch(const float* emb_all, const float* gt_all, const float* dist_all, const int64_t* idx,
                     int64_t n, int64_t n_rows, int32_t emb_floats, float* emb_out, float* gt_out, float* dist_out,
                     int32_t* status, void* stream);

/*
 * Latent bitstream: the arithmetic code of the rounded latents.  Replaces the
 * `./module_arithmeticcoding e|d 1 1` subprocess of encode()/decode() (NVFPCC.py:446-477,
 * 588-607; module_arithmeticcoding.cpp:368-432) with an in-process call producing / consuming
 * the SAME byte stream (`latent_pack['latent_byte_stream']`).  HOST pointers throughout.
 *   symbols  [n] int16 in [0, 1024]: latent + 512 (NVFPCC.py:447,453)
 *   mu,sigma [n] float32 per-symbol Gaussian parameters (mu already offset by 512, :455-460)
 *   level1/2 number of low mantissa bits cleared in mu / sigma (the helper's argv[2], argv[3]; 1 and 1)
 *   out      capacity out_cap >= nvf_arith_encode_bound(n); *out_len = bytes written
 * nvf_arith_decode_host returns NVF_ERR_BITSTREAM when the stream cannot have been produced
 * with these parameters (the reference helper asserts in that case).
 */
int nvf_arith_encode_bound(int64_t n_symbols, size_t* bytes_out);
int nvf_arith_encode_host(const int16_t* symbols, const float* mu, const float* sigma, int64_t n, int level1,
                          int level2, uint8_t* out, size_t out_cap, size_t* out_len);
int nvf_arith_decode_host(const uint8_t* stream, size_t stream_len, const float* mu, const float* sigma, int64_t n,
                          int level1, int level2, int16_t* symbols_out);

/*
 * Weight bitstream: Huffman code of the 1/16-quantised kernels.  Replaces entropy_encode /
 * entropy_decode of util_code_quantized_weights.py:108-148 (a Python walk over a '0'/'1'
 * string, one slice per bit) for the codebook carried in the pack (`inv_codebook`).  HOST pointers.
 *   code_symbols [n_codes] int32, code_lengths [n_codes] (<= 64), code_bits [n_codes] codeword
 *   value, first bit of the codeword = most significant of its `length` bits; stream bits are
 *   packed MSB first and zero padded to a byte boundary.
 */
int nvf_huffman_encode_host(const int32_t* symbols, int64_t n, const int32_t* code_symbols,
                            const uint8_t* code_lengths, const uint64_t* code_bits, int32_t n_codes, uint8_t* out,
                            size_t out_cap, size_t* out_len);
int nvf_huffman_decode_host(const uint8_t* stream, size_t stream_len, const int32_t* code_symbols,
                            const uint8_t* code_lengths, const uint64_t* code_bits, int32_t n_codes,
                            int64_t n_symbols, int32_t* symbols_out);

#ifdef __cplusplus
}
#endif
#endif /* NVF_PREP_B200_H_ */
