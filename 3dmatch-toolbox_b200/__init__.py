"""b200 TSDF fusion: B200-native multi-frame TSDF fusion engine behind the drop-in surface of
3dmatch-toolbox's depth-fusion/demo (C ABI in include/tsdf_fusion.h, CUDA in csrc/).

The directory name is not a Python identifier; load it with `__graft_entry__.load_package()`.
"""
from . import binding, sharding, synth  # noqa: F401
from .binding import (  # noqa: F401
    ComputeTDF, Integrate, LoadMatrixFromFile, ReadDepth, SaveVoxelGrid2SurfacePointCloud, ShardedVolume, TsdfError, Volume,
    cam2base_from_world, depth_from_u16, integrate_multi, invert_matrix, load_library, multiply_matrix, plan_z_cost, split_slabs, point_cloud_to_tdf, tdf_grid_from_points, write_bin, write_ply,
)
