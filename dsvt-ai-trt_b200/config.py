"""Static capacities and geometry: the reference's include/params.h and the Waymo-shape bench config."""
from dataclasses import dataclass, replace


@dataclass(frozen=True)
class HotPathConfig:
    # points2Features (params.h:24-41)
    max_points_num: int = 50000
    max_points_num_voxel_filter: int = 30000
    max_pillars_num: int = 10000
    max_num_points_per_voxel: int = 48
    point_feature_num: int = 4
    feature_num: int = 10
    x_min: float = -74.88
    x_max: float = 74.88
    y_min: float = -74.88
    y_max: float = 74.88
    z_min: float = -5.0
    z_max: float = 3.0
    voxel_x: float = 0.32
    voxel_y: float = 0.32
    voxel_z: float = 8.0
    grid_x: int = 468
    grid_y: int = 468
    grid_z: int = 1
    # dsvt input layer (params.h:52-70)
    win_shapes: tuple = ((12, 12, 1), (24, 24, 1))
    shifts: tuple = ((0, 0, 0), (6, 6, 0))
    max_voxel_num_per_win: int = 576
    max_win_num: int = 800
    voxel_num_set: int = 36
    # dsvt blocks (params.h:73-84)
    num_heads: int = 8
    channel_num: int = 192
    ffn_channel_num: int = 384
    num_blocks: int = 4
    pfn_channels: tuple = (96, 192)     # PFN_LAYER_0/1_OUT_CHANNEL (params.h:43-44): feature widths of the two scatter-max calls
    layer_norm_eps: float = 0.0          # effective value in the reference
    # post-processing (params.h:327-328)
    max_top_k: int = 500
    score_threshold: float = 0.3

    def with_(self, **kw):
        return replace(self, **kw)


REFERENCE = HotPathConfig()

# BASELINE.json configs[1]: 200k-point synthetic cloud; capacities per SURVEY.md 8(d)
WAYMO = HotPathConfig(max_points_num=320000, max_points_num_voxel_filter=320000, max_pillars_num=40000,
                      max_win_num=4096)
# pillar 0.30 x 0.30 variant named by BASELINE.json (grid 500)
WAYMO_030 = WAYMO.with_(voxel_x=0.30, voxel_y=0.30, x_min=-75.0, x_max=75.0, y_min=-75.0, y_max=75.0,
                        grid_x=500, grid_y=500)
