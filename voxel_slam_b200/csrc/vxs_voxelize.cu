// Voxel-map construction on the GPU (placeholder translation unit; replaced by the real kernels in the next milestone).
#include "vxs_internal.h"
void vxs_voxelize_release(vxs_ctx*) {}
extern "C" int vxs_voxel_keys(vxs_ctx* ctx, const double*, int64_t, double, int64_t*, uint64_t*) { return vxs_fail(ctx, VXS_ERR_ARG, "vxs_voxel_keys: not built yet"); }
extern "C" int vxs_build_window_factor(vxs_ctx* ctx, const vxs_map_params*, const double*, const int64_t*, const double*, int, const double*, int64_t, vxs_factor*, vxs_voxel_id*, int64_t, int64_t*) { return vxs_fail(ctx, VXS_ERR_ARG, "vxs_build_window_factor: not built yet"); }
extern "C" int vxs_build_gba_factor(vxs_ctx* ctx, const vxs_map_params*, const float*, int, const int64_t*, const double*, int, vxs_factor*, vxs_voxel_id*, int64_t, int64_t*) { return vxs_fail(ctx, VXS_ERR_ARG, "vxs_build_gba_factor: not built yet"); }
extern "C" int vxs_hba_window(vxs_ctx* ctx, const vxs_map_params*, const vxs_map_params*, const float*, int, const int64_t*, double*, int, int, int, double*, double*, int*) { return vxs_fail(ctx, VXS_ERR_ARG, "vxs_hba_window: not built yet"); }
