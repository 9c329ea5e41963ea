/* b2o_obs.cpp -- CPU oracle: depth/segmentation raster and segmented point cloud.
 * TEST INFRASTRUCTURE ONLY (see b2o_world.h). */
#include "b2o_world.h"
namespace b2o {
void render(World&, int) {}
void point_cloud(World&, int, uint64_t) {}
}
