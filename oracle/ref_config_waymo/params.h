/*
 * Build-time configuration for compiling the REFERENCE's plugin sources at the bench's Waymo-shape
 * capacities (bench.py --impl reference).  The reference is configured exclusively through the
 * #defines of its include/params.h; its kernels size grids and index arrays from these
 * macros, so a 200k-point cloud needs a rebuilt configuration.  This file is placed FIRST on the
 * include path (oracle/build.py) and defines exactly the macros the plugin sources consume; the
 * reference sources themselves stay unmodified and in place.
 */
#ifndef PARAMS_H_
#define PARAMS_H_

#define THREADS_FOR_VOXEL 256

#define MAX_POINTS_NUM 320000
#define MAX_POINTS_NUM_1 320000
#define MAX_PILLARS_NUM 40000
#define POINTS_NUM_PER_VOXEL 48
#define FEATURES_NUM 10
#define ARRAY_LEN 200

#define MAX_VOXEL_NUM_PER_WIN 576
#define MAX_WIN_NUM 4096
#define VOXEL_NUM_SET 36

#define NUM_HEADS 8
#define GELU_A 0.5
#define GELU_B 0.7978845608028654
#define GELU_C 0.035677408136300125
#define POSEMBED_LAYBERS_OUT_FEATURES 192
#define SET_ATTENTION_0_0_GELU_OUT_CHANNEL 384

#define HM_TOP_K 500
#define LAST_DIMS 9

#endif
