// b2s_obs.cu -- depth / segmentation raster and segmented point cloud (placeholder until the raster lands)
#include "b2s_dev.cuh"
void b2s_launch_render(const DWorld& W, cudaStream_t s) {}
void b2s_launch_point_cloud(const DWorld& W, uint64_t seed, cudaStream_t s) {}
void b2s_launch_staged(const DWorld& W, int n, cudaStream_t s, int64_t* launches) {
  for (int i = 0; i < n; ++i) b2s_launch_substeps(W, 1, MODE_RAW, 0, 0, 0, s);
  *launches = n;
}
