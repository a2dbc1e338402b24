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

#ifdef __cplusplus
}
#endif
#endif /* NVF_PREP_B200_H_ */
