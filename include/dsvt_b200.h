/*
 * dsvt_b200.h -- C ABI of the B200-native DSVT hot path.
 *
 * Every entry point replaces the device work of ONE reference TensorRT plugin
 * `enqueue` (or, for the set attention, of the TensorRT layer sub-graph built by
 * multHeadAttention()).  Tensors are plain device pointers with exactly the
 * reference's layouts; valid counts travel as device-side int32 (never synced
 * to the host); all work is enqueued on the caller's stream; nothing throws.
 * File:line citations are relative to the reference tree (jingyue202205/DSVT-AI-TRT
 * @15b31c3).
 *
 * The `batch` member of every params struct is an extension: the reference
 * kernels are batch-1.  With batch = B every tensor carries a
 * leading B dimension with the reference's static capacities as strides and the
 * count tensors hold B entries; B = 1 is the reference contract.
 *
 * Return value: DSVT_OK (0) on success, non-zero otherwise (the reference
 * abort()s on CUDA errors -- getSet.cu:8-19; we return a code instead).
 */
#ifndef DSVT_B200_H
#define DSVT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSVT_B200_ABI_VERSION 1

typedef struct CUstream_st* dsvt_stream_t; /* == cudaStream_t */

enum {
    DSVT_OK = 0,
    DSVT_ERR_INVALID_ARGUMENT = 1,
    DSVT_ERR_CUDA = 2,
    DSVT_ERR_WORKSPACE_TOO_SMALL = 3,
    DSVT_ERR_UNSUPPORTED = 4
};

/* version / diagnostics */
int dsvt_abi_version(void);
const char* dsvt_last_error(void);          /* thread-local, human readable */
int dsvt_device_sm_count(void);
/* number of kernel launches (incl. memset nodes) issued by this library in this process */
uint64_t dsvt_launch_count(void);

/* ------------------------------------------------------------------------ *
 * a1  Points2FeaturesPlugin::enqueue           plugins/src/points2Features.cu:896-990
 *     (kernels :669-705, :732-766, :792-865)
 * ------------------------------------------------------------------------ */
typedef struct dsvt_points2features_params {
    int32_t batch;
    int32_t max_points_num;              /* rows of `points` per frame          (MAX_POINTS_NUM 50000)   */
    int32_t max_points_num_voxel_filter; /* rows of `point_features` per frame  (MAX_POINTS_NUM_1 30000) */
    int32_t max_pillars_num;             /* (MAX_PILLARS_NUM 10000) */
    int32_t point_feature_num;           /* must be 4  */
    int32_t feature_num;                 /* must be 10 */
    int32_t max_num_points_per_voxel;    /* <= 64; reference 48 */
    float x_min, x_max, y_min, y_max, z_min, z_max;
    float voxel_x, voxel_y, voxel_z;
    int32_t grid_x, grid_y, grid_z;      /* grid_z must be 1 (pillars) */
    int32_t zero_tails;                  /* 1: rows beyond the valid counts are zero-filled (reference contract) */
} dsvt_points2features_params;

size_t dsvt_points2features_workspace_size(const dsvt_points2features_params* p);
/*
 * in : points [B,max_points_num,4] f32 ; points_size [B] i32
 * out: point_features [B,max_points_num_voxel_filter,10] f32
 *      point_index_in_voxel [B,max_pillars_num,max_num_points_per_voxel] i32 (row ids into point_features)
 *      coords [B,max_pillars_num,4] i32 = (0,0,y,x)
 *      point_num_in_voxel [B,max_pillars_num] i32 ; pillar_num [B] i32 ; point_num [B] i32
 * Order is canonical and deterministic: pillars ascending y*grid_x+x,
 * points of a pillar ascending input index (lowest 48 kept), rows pillar-major.
 */
int dsvt_points2features_launch(const dsvt_points2features_params* p,
                                const float* points, const int32_t* points_size,
                                float* point_features, int32_t* point_index_in_voxel, int32_t* coords,
                                int32_t* point_num_in_voxel, int32_t* pillar_num, int32_t* point_num,
                                void* workspace, size_t workspace_bytes, dsvt_stream_t stream);

/* ------------------------------------------------------------------------ *
 * (next #1) WindowPartitionPlugin::enqueue     plugins/src/windowPartition.cu:397-470 (kernel :278-381)
 * ------------------------------------------------------------------------ */
typedef struct dsvt_window_partition_params {
    int32_t batch;
    int32_t max_pillars_num;
    int32_t max_win_num;
    int32_t max_voxel_num_per_win;
    int32_t sparse_shape_x, sparse_shape_y, sparse_shape_z;
    int32_t win_shape_x, win_shape_y, win_shape_z;
    int32_t shift_x, shift_y, shift_z;
    int32_t zero_tails;
} dsvt_window_partition_params;

size_t dsvt_window_partition_workspace_size(const dsvt_window_partition_params* p);
/*
 * in : coords [B,max_pillars_num,4] i32 (0,z,y,x) ; voxel_num [B]
 * out: global_index [B,max_win_num,max_voxel_num_per_win] i32
 *      coors_in_win [B,max_win_num,max_voxel_num_per_win,3] i32 (z,y,x)
 *      voxel_num_in_win [B,max_win_num] i32 ; win_num [B] i32
 *      coors_in_win_2d [B,max_pillars_num,3] i32 ; coors_in_win_x_y [B,max_pillars_num,2] f32
 * Canonical order: windows ascending dense window index, voxels ascending voxel id.
 */
int dsvt_window_partition_launch(const dsvt_window_partition_params* p,
                                 const int32_t* coords, const int32_t* voxel_num,
                                 int32_t* global_index, int32_t* coors_in_win, int32_t* voxel_num_in_win,
                                 int32_t* win_num, int32_t* coors_in_win_2d, float* coors_in_win_x_y,
                                 void* workspace, size_t workspace_bytes, dsvt_stream_t stream);

/* ------------------------------------------------------------------------ *
 * a2  GetSetPlugin::enqueue                    plugins/src/getSet.cu:629-704
 *     (kernels :326-350, :369-422, :444-495, :517-567, :589-609)
 * ------------------------------------------------------------------------ */
typedef struct dsvt_get_set_params {
    int32_t batch;
    int32_t voxel_num_set;          /* 36 */
    int32_t max_win_num;            /* capacity of windows AND of sets (SURVEY.md A-6 vii) */
    int32_t max_voxel_num_per_win;  /* 576 */
    int32_t win_shape_x, win_shape_y, win_shape_z;
    int32_t num_heads;              /* NUM_HEADS 8 (params.h:73) */
    int32_t zero_tails;
} dsvt_get_set_params;

size_t dsvt_get_set_workspace_size(const dsvt_get_set_params* p);
/*
 * in : global_index [B,max_win_num,max_voxel_num_per_win] i32 ; coors_in_win [..,3] i32 (z,y,x)
 *      voxel_num_in_win [B,max_win_num] i32 ; win_num [B] i32
 * out: global_index_in_set [B,2,max_win_num,S] i32 (plane 0 Y-major order, plane 1 X-major)
 *      set_voxel_mask [B,2,max_win_num,S] f32 (0 or -FLT_MAX)
 *      set_num [B] i32
 *      mask_expand_0 / mask_expand_1 [B,max_win_num,num_heads,S] f32
 * Sets are numbered window-major in ascending window slot (canonical form of the
 * reference's atomicAdd race, getSet.cu:337).
 */
int dsvt_get_set_launch(const dsvt_get_set_params* p,
                        const int32_t* global_index, const int32_t* coors_in_win,
                        const int32_t* voxel_num_in_win, const int32_t* win_num,
                        int32_t* global_index_in_set, float* set_voxel_mask, int32_t* set_num,
                        float* mask_expand_0, float* mask_expand_1,
                        void* workspace, size_t workspace_bytes, dsvt_stream_t stream);

/* ------------------------------------------------------------------------ *
 * a4  GeluPlugin::enqueue                      plugins/src/gelu.cu:227-251 (kernel :201-211)
 * ------------------------------------------------------------------------ */
typedef struct dsvt_gelu_params {
    int32_t batch;
    int32_t max_pillars_num;
    int32_t channel_num;            /* 384 */
    int32_t zero_tails;
} dsvt_gelu_params;
/* in: x [B,max_pillars_num,C] f32 ; voxel_num [B] i32 -> out same shape */
int dsvt_gelu_launch(const dsvt_gelu_params* p, const float* x, const int32_t* voxel_num, float* out,
                     dsvt_stream_t stream);

/* ------------------------------------------------------------------------ *
 * a5  LayerNormPlugin::enqueue                 plugins/src/layerNorm.cu:357-402
 *     (kernels :261-279, :297-309, :326-338)
 * ------------------------------------------------------------------------ */
typedef struct dsvt_layer_norm_params {
    int32_t batch;
    int32_t max_pillars_num;
    int32_t channel_num;            /* 192 */
    float eps;                      /* the reference runs with 0.0 */
    int32_t zero_tails;
} dsvt_layer_norm_params;
/*
 * in: x [B,max_pillars_num,C] f32 ; voxel_num [B] ; gamma, beta [C] f32 (device)
 * residual: optional second addend with x's shape (NULL = reference behaviour).  When given the
 * kernel normalises x + residual, i.e. it absorbs the addElementWise(kSUM) that precedes every
 * LayerNorm in the reference graph (src/dsvt-ai-trt.cpp:669-697).
 */
int dsvt_layer_norm_launch(const dsvt_layer_norm_params* p, const float* x, const float* residual,
                           const int32_t* voxel_num, const float* gamma, const float* beta, float* out,
                           dsvt_stream_t stream);

/*
 * Chain of up to 3 consecutive LayerNorms whose only intermediate op is a residual add -- e.g. norm2 -> norm ->
 * residual_norm at the end of a DSVT block (src/dsvt-ai-trt.cpp:685-697, :750-756):
 *   y = LN_n( ... LN_2( LN_1(x + r_1) + r_2 ) ... + r_n ).   The row stays in registers between stages, so the
 * chain costs one read of x, one read per residual and ONE write instead of n read-write round trips.
 * Results are identical to n calls of dsvt_layer_norm_launch.  residual may be NULL per stage.
 */
typedef struct dsvt_ln_stage {
    const float* residual;   /* [B,max_pillars_num,C] or NULL */
    const float* gamma;      /* [C] device */
    const float* beta;       /* [C] device */
} dsvt_ln_stage;
int dsvt_layer_norm_chain_launch(const dsvt_layer_norm_params* p, const float* x, const int32_t* voxel_num,
                                 const dsvt_ln_stage* stages, int32_t n_stages, float* out, dsvt_stream_t stream);

/* ------------------------------------------------------------------------ *
 * a6  FilterBoxByScorePlugin::enqueue          plugins/src/filterBoxByScore.cu:328-379 (kernel :266-309)
 * ------------------------------------------------------------------------ */
typedef struct dsvt_filter_box_params {
    int32_t batch;
    int32_t max_top_k;              /* 500 */
    float x_min, x_max, y_min, y_max, z_min, z_max;
    float voxel_x, voxel_y, voxel_z;
    float score_threshold;
    int32_t zero_tails;
} dsvt_filter_box_params;
/*
 * in : scores [B,K] f32 ; classes, xs, ys [B,K] i32 ; center [B,K,2] ; center_z [B,K] ;
 *      angle [B,K] ; dim [B,K,3] f32
 * out: boxes [B,K,9] f32 (x,y,z,dx,dy,dz,angle,class,score), valid [B] i32.
 * Kept candidates are appended in ascending candidate index (canonical form of the
 * atomicAdd order, filterBoxByScore.cu:295).
 */
int dsvt_filter_box_launch(const dsvt_filter_box_params* p,
                           const float* scores, const int32_t* classes, const int32_t* xs, const int32_t* ys,
                           const float* center, const float* center_z, const float* angle, const float* dim,
                           float* boxes, int32_t* valid, dsvt_stream_t stream);

/* ------------------------------------------------------------------------ *
 * a3  multHeadAttention()                      src/dsvt-ai-trt.cpp:288-458
 *     (+ next #2: GetValueByIndex getValueByIndex.cu:282-303 and
 *      MapSetFeature2Voxel mapSetFeature2voxel.cu:258-275 fused in)
 * ------------------------------------------------------------------------ */
enum {
    DSVT_ATTN_FP32 = 0,      /* CUDA-core FP32 contractions: the FP32 configuration (tolerance 1e-3)                */
    /* 1 is not a set-attention precision: see DSVT_LINEAR_TF32 below */
    DSVT_ATTN_FP16 = 2,      /* tcgen05 kind::f16, FP16 operands, FP32 accumulate: the reference's USE_FP16
                                configuration (params.h:332; tolerance 1e-2); one fused kernel, QK^T / PV on tcgen05 */
    DSVT_ATTN_FP32_TC = 3,   /* FP32-accurate on tcgen05 (fused entry point, needs a workspace): per-voxel projection
                                GEMMs with every operand split into FP16 hi + lo and three kind::f16 MMAs per product
                                (2^-22 relative), FP32 accumulate; per-set QK^T / softmax / PV in FP32 on CUDA cores.
                                Same tolerance as DSVT_ATTN_FP32.  Domain: |x + pos| < 65504.                        */
    DSVT_ATTN_FP16_GEMM = 4  /* the pipeline of DSVT_ATTN_FP32_TC with single FP16 operands (tolerance 1e-2)         */
};

typedef struct dsvt_set_attention_params {
    int32_t batch;
    int32_t max_set_num;            /* MAX_WIN_NUM 800 */
    int32_t voxel_num_set;          /* 36 */
    int32_t channel_num;            /* 192 */
    int32_t num_heads;              /* 8 */
    int32_t max_pillars_num;        /* only for the fused entry point */
    int32_t axis_id;                /* plane of global_index_in_set used by the fused entry point */
    int32_t precision;              /* DSVT_ATTN_* */
    int32_t zero_tails;
} dsvt_set_attention_params;

/* Opaque, device-resident, pre-arranged copy of one attention layer's weights. */
typedef struct dsvt_attention_weights dsvt_attention_weights;
/* host pointers, PyTorch layouts: in_proj_weight [3C,C], in_proj_bias [3C], out_proj.weight [C,C], out_proj.bias [C]
 * (rows 0..C-1 = query, C..2C-1 = key, 2C..3C-1 = value; include/helper.h:367-433) */
dsvt_attention_weights* dsvt_attention_weights_create(int32_t channel_num, int32_t num_heads,
                                                      const float* in_proj_weight, const float* in_proj_bias,
                                                      const float* out_proj_weight, const float* out_proj_bias);
void dsvt_attention_weights_destroy(dsvt_attention_weights* w);

/* bytes of workspace the launch entry points need for p->precision; max_pillars_num == 0 selects the plugin-shaped form */
size_t dsvt_set_attention_workspace_size(const dsvt_set_attention_params* p);
/*
 * Plugin-shaped form (drop-in for the multHeadAttention() sub-graph):
 * in : q, k, v [B,max_set_num,S,C] f32 ; mask [B,max_set_num,num_heads,S] f32 (additive key mask)
 *      set_num [B] i32 or NULL (NULL = all max_set_num sets, as the reference graph does)
 * out: [B,max_set_num,S,C] f32
 * Precisions: DSVT_ATTN_FP32 (CUDA cores, no workspace) and the tensor-core pipeline DSVT_ATTN_FP32_TC / DSVT_ATTN_FP16_GEMM
 * (projections of the set slots on tcgen05, per-set core in FP32; workspace of dsvt_set_attention_workspace_size bytes with
 * max_pillars_num = 0; q, k, v 32-byte aligned).
 */
int dsvt_set_attention_launch(const dsvt_set_attention_params* p, const dsvt_attention_weights* w,
                              const float* q, const float* k, const float* v, const float* mask,
                              const int32_t* set_num, float* out,
                              void* workspace, size_t workspace_bytes, dsvt_stream_t stream);
/*
 * Fused form: q = k = x[idx] + pos[idx], v = x[idx]; result rows are scattered back to voxel rows.
 * in : x, pos [B,max_pillars_num,C] f32 ; global_index_in_set [B,2,max_set_num,S] i32 ;
 *      mask [B,max_set_num,num_heads,S] f32 ; set_num [B] ; voxel_num [B]
 * out: [B,max_pillars_num,C] f32 (rows >= voxel_num zero when zero_tails)
 */
int dsvt_set_attention_fused_launch(const dsvt_set_attention_params* p, const dsvt_attention_weights* w,
                                    const float* x, const float* pos, const int32_t* global_index_in_set,
                                    const float* mask, const int32_t* set_num, const int32_t* voxel_num,
                                    float* out, void* workspace, size_t workspace_bytes, dsvt_stream_t stream);

/*
 * Attention plan (GEMM-pipeline precisions DSVT_ATTN_FP32_TC / DSVT_ATTN_FP16_GEMM): the set partition of one
 * (frame, window partition, axis_id) in token order -- distinct tokens per set, their prefix sums and the voxel -> token
 * map.  It depends only on global_index_in_set / mask / set_num, i.e. on the GetSetPlugin outputs, so ONE plan serves
 * every attention layer that uses the partition (2 of the 8 layers each in the reference graph,
 * src/dsvt-ai-trt.cpp:653-1116).  dsvt_set_attention_fused_launch builds it internally on every call (stateless);
 * dsvt_set_attention_fused_planned_launch takes a prebuilt one (256-byte aligned, dsvt_set_attention_plan_size bytes).
 */
size_t dsvt_set_attention_plan_size(const dsvt_set_attention_params* p);
int dsvt_set_attention_plan_launch(const dsvt_set_attention_params* p, const int32_t* global_index_in_set,
                                   const float* mask, const int32_t* set_num, void* plan, size_t plan_bytes,
                                   dsvt_stream_t stream);
int dsvt_set_attention_fused_planned_launch(const dsvt_set_attention_params* p, const dsvt_attention_weights* w,
                                            const float* x, const float* pos, const int32_t* global_index_in_set,
                                            const float* mask, const int32_t* set_num, const int32_t* voxel_num,
                                            float* out, const void* plan, void* workspace, size_t workspace_bytes,
                                            dsvt_stream_t stream);

/*
 * The fused form followed by the encoder layer's first residual add + LayerNorm in the same launch sequence:
 *   out = LayerNorm(attention(x, pos) + residual) * gamma + beta      (norm1(src2 + src), src/dsvt-ai-trt.cpp:669-676;
 * addElementWise(kSUM) + LayerNormPlugin) -- the LayerNorm runs in the out-projection's epilogue, the attention output
 * itself never reaches memory (GEMM-pipeline precisions; the single-kernel precisions run attention + the row-wise LayerNorm
 * kernel in place, same results, one more launch).  plan may be NULL (rebuilt per call).  residual
 * [B,max_pillars_num,C], gamma / beta [C] on the device.  Arithmetic of the LayerNorm = dsvt_layer_norm_launch up to
 * the quotient by the standard deviation, which the epilogue evaluates as a reciprocal-multiply (<= 1 ulp per element).
 */
int dsvt_set_attention_fused_norm_launch(const dsvt_set_attention_params* p, const dsvt_attention_weights* w,
                                         const float* x, const float* pos, const int32_t* global_index_in_set,
                                         const float* mask, const int32_t* set_num, const int32_t* voxel_num,
                                         const float* residual, const float* gamma, const float* beta, float eps,
                                         float* out, const void* plan, void* workspace, size_t workspace_bytes,
                                         dsvt_stream_t stream);

/*
 * Instrumentation form of the two entry points above (GEMM-pipeline precisions, prebuilt plan): launches only the kernels
 * named in `stages` -- bit 0 the QKV projection GEMM, bit 1 the per-set core, bit 2 the out-projection GEMM (with the norm
 * epilogue when residual != NULL) -- on the workspace the earlier stages left behind, so that a caller can bracket each
 * kernel with its own CUDA events (bench.py's per-kernel roofline).  stages = 7 is the normal call.  No global state.
 */
int dsvt_set_attention_fused_stages_launch(const dsvt_set_attention_params* p, const dsvt_attention_weights* w,
                                           const float* x, const float* pos, const int32_t* global_index_in_set,
                                           const float* mask, const int32_t* set_num, const int32_t* voxel_num,
                                           const float* residual, const float* gamma, const float* beta, float eps,
                                           float* out, const void* plan, void* workspace, size_t workspace_bytes,
                                           int32_t stages, dsvt_stream_t stream);

/*
 * dsvt_set_attention_fused_stages_launch with the position embedding given as a TABLE: the embedding of a voxel is a function
 * of its cell (cx, cy) in its window only (the MLP's input is coors_in_win_x_y = (cx - win_x / 2, cy - win_y / 2),
 * src/dsvt-ai-trt.cpp:603-637, windowPartition.cu:358-359), so the MLP needs evaluating once per cell -- at most 12 x 12 or
 * 24 x 24 rows -- instead of once per voxel.  pos_table [win_x * win_y, 192] f32, row cy * win_x + cx (the MLP's output for that
 * cell, e.g. from dsvt_pos_embed_mlp_batch_launch on the list of cells); coors_in_win_2d [B,max_pillars_num,3] i32 =
 * WindowPartitionPlugin output 4 (cz, cy, cx).  Results = the per-voxel form with pos[v] = pos_table[cell(v)].
 * DSVT_ATTN_FP32_TC only.
 */
int dsvt_set_attention_fused_table_launch(const dsvt_set_attention_params* p, const dsvt_attention_weights* w,
                                          const float* x, const float* pos_table, const int32_t* coors_in_win_2d,
                                          int32_t win_shape_x, const int32_t* global_index_in_set, const float* mask,
                                          const int32_t* set_num, const int32_t* voxel_num, const float* residual,
                                          const float* gamma, const float* beta, float eps, float* out,
                                          const void* plan, void* workspace, size_t workspace_bytes,
                                          int32_t stages, dsvt_stream_t stream);

/* ------------------------------------------------------------------------ *
 * (next #4) dense linear layer  y = x * W^T + b on tcgen05 -- replaces the TensorRT FullyConnected
 * layers next to the plugins (fullyConnected_gelu_fullyConnected, src/dsvt-ai-trt.cpp:494-529).
 * W [N,K] row-major (PyTorch [out,in]), N % 64 == 0, K % 16 == 0; precision DSVT_LINEAR_TF32 (single-tile
 * tcgen05 kind::tf32) or DSVT_ATTN_FP16 (kind::f16); DSVT_ATTN_FP32_TC / DSVT_ATTN_FP16_GEMM: see below.
 * ------------------------------------------------------------------------ */
enum { DSVT_LINEAR_TF32 = 1 };   /* tcgen05 kind::tf32 operands, FP32 accumulate in TMEM -- dense linear layers only */
typedef struct dsvt_linear_weights dsvt_linear_weights;
dsvt_linear_weights* dsvt_linear_weights_create(int32_t N, int32_t K, const float* W, const float* b,
                                                int32_t precision);
void dsvt_linear_weights_destroy(dsvt_linear_weights* w);
/* x [M,K] f32 (device) -> y [M,N] f32 (device) */
int dsvt_linear_launch(const dsvt_linear_weights* w, const float* x, int32_t M, float* y, dsvt_stream_t stream);
/*
 * Weights created with precision DSVT_ATTN_FP32_TC (FP32-accurate: FP16 hi+lo split operands, 3 tcgen05 MMAs per product)
 * or DSVT_ATTN_FP16_GEMM run on the attention pipeline's persistent GEMM kernel; they need N % 192 == 0, K % 192 == 0
 * (the FFN 192->384->192 and PFN / pos-embed 192->192 layers).  This form takes the valid row count from the device
 * (rows [1] i32, as the plugins do), optionally applies GELU (activation = 1: the FFN's first linear + GeluPlugin in one
 * pass) and zero-fills rows [rows, max_rows).  x [max_rows,K] -> y [max_rows,N].
 */
int dsvt_linear_rows_launch(const dsvt_linear_weights* w, const float* x, const int32_t* rows, int32_t max_rows,
                            int32_t activation, float* y, int32_t zero_tails, dsvt_stream_t stream);
/*
 * Same layer with the input row given as the concatenation [x_lo (k_split columns) | x_hi (K - k_split columns)] of two
 * dense tensors, read in place: the PFN's addConcatenation (src/dsvt-ai-trt.cpp:583-587) is not materialised.  K == 192,
 * k_split a multiple of 32.  activation: 0 none, 1 GELU, 2 ReLU (a folded BatchNorm goes into W and b:
 * fullyConnectedBnLELU, src/dsvt-ai-trt.cpp:268-286).
 */
int dsvt_linear_rows_concat_launch(const dsvt_linear_weights* w, const float* x_lo, const float* x_hi, int32_t k_split,
                                   const int32_t* rows, int32_t max_rows, int32_t activation, float* y, int32_t zero_tails,
                                   dsvt_stream_t stream);

/*
 * Split-K form for N == 192, K = 192 * kb (kb <= 3; the FFN's second linear 384 -> 192): ONE launch, K block j writes its
 * partial product to y_parts[j] ([max_rows,192] each, rows beyond `rows` untouched); bias and the optional residual rows
 * `add` [max_rows,192] (the addElementWise(kSUM) behind the FFN, src/dsvt-ai-trt.cpp:685) go into part 0.  The result is
 * the sum of the parts: pass part 0 as x and part 1 as the residual of the LayerNorm that follows.
 */
int dsvt_linear_rows_splitk_launch(const dsvt_linear_weights* w, const float* x, const float* add, const int32_t* rows,
                                   int32_t max_rows, float* y_parts, dsvt_stream_t stream);

/*
 * A [*, K] -> [*, 192] layer (K = 192 or 384: the FFN's second linear) followed by up to three (residual add + LayerNorm)
 * stages in ONE kernel:  y0 = x W^T + b;  y = LN_s(y + stages[s].residual) for s < n_stages  -- the kSUM + LayerNormPlugin
 * pairs behind the FFN (norm2(src + src2), norm(src + x), the block's residual norm; src/dsvt-ai-trt.cpp:685-697, :750-756).
 * Rows beyond `rows` are zero-filled when zero_tails.  Same arithmetic per stage as dsvt_layer_norm_chain_launch.
 */
int dsvt_linear_rows_norm_launch(const dsvt_linear_weights* w, const float* x, const int32_t* rows, int32_t max_rows,
                                 const dsvt_ln_stage* stages, int32_t n_stages, float eps, float* y, int32_t zero_tails,
                                 dsvt_stream_t stream);

/*
 * The FFN of one encoder layer as ONE kernel (fullyConnected_gelu_fullyConnected, src/dsvt-ai-trt.cpp:494-529, + the
 * addElementWise / LayerNormPlugin pairs behind it, :685-697 and :750-756):
 *   y = LN_n( ... LN_1( gelu(x W1^T + b1) W2^T + b2 + r_1 ) ... + r_n )
 * fc1 = Linear(192 -> 384), fc2 = Linear(384 -> 192), both created with DSVT_ATTN_FP32_TC.  The 384-wide hidden rows stay on
 * the SM (second GEMM's A operand in tensor memory); results = dsvt_linear_rows_launch(activation GELU) followed by
 * dsvt_linear_rows_norm_launch.  x [max_rows,192] (32-B aligned), y [max_rows,192], rows [1] on the device (batch 1).
 */
int dsvt_ffn_fused_launch(const dsvt_linear_weights* fc1, const dsvt_linear_weights* fc2, const float* x,
                          const int32_t* rows, int32_t max_rows, const dsvt_ln_stage* stages, int32_t n_stages,
                          float eps, float* y, int32_t zero_tails, dsvt_stream_t stream);

/*
 * The tail of an encoder layer as ONE kernel: the attention's out-projection, norm1(attention + x) (src/dsvt-ai-trt.cpp:669-676),
 * the FFN (:494-529) and the LayerNorm chain behind it (:685-697, :750-756).  Call after
 * dsvt_set_attention_fused_stages_launch(..., stages = 3) (QKV projection + per-set core) with the SAME params, weights, plan and
 * workspace: the kernel reads the core's rows from the workspace.  `src` [max_pillars_num,192] receives norm1's output; the FIRST
 * LayerNorm stage's residual is `src` whatever stages[0].residual says (norm2(src + ffn)), the others are the caller's.
 * Results = dsvt_set_attention_fused_norm_launch followed by dsvt_ffn_fused_launch.  DSVT_ATTN_FP32_TC, batch 1.
 */
int dsvt_attention_tail_ffn_launch(const dsvt_set_attention_params* p, const dsvt_attention_weights* w, const void* plan,
                                   void* workspace, size_t workspace_bytes, const int32_t* voxel_num, const float* x,
                                   const float* norm1_gamma, const float* norm1_beta, float norm1_eps,
                                   const dsvt_linear_weights* fc1, const dsvt_linear_weights* fc2,
                                   const dsvt_ln_stage* stages, int32_t n_stages, float eps, float* src, float* y,
                                   dsvt_stream_t stream);

/*
 * Narrow first layers of the VFE / position-embedding MLPs (TensorRT FullyConnected + Scale + ReLU in the reference:
 * PFN layer 0 Linear(10 -> 96) src/dsvt-ai-trt.cpp:577, position embedding Linear(2 -> 192) :603-637 via :461-492):
 *   y = act((x W^T) * scale + shift),  W [N,K] row-major, K in [1, 16], N / 4 dividing 192, scale / shift [N] = the folded
 *   BatchNorm1d (NULL: 1 / 0).  x [B,max_rows,K] -> y [B,max_rows,N]; rows [B] on the device; rows beyond it zero-filled.
 */
typedef struct dsvt_small_linear dsvt_small_linear;
dsvt_small_linear* dsvt_small_linear_create(int32_t N, int32_t K, const float* W, const float* scale, const float* shift);
void dsvt_small_linear_destroy(dsvt_small_linear* w);
int dsvt_small_linear_launch(const dsvt_small_linear* w, const float* x, const int32_t* rows, int32_t batch,
                             int32_t max_rows, int32_t activation, float* y, int32_t zero_tails, dsvt_stream_t stream);

/*
 * The pillar feature net as ONE kernel (src/dsvt-ai-trt.cpp:571-590): PFN layer 0 (Linear(10 -> 96) + BatchNorm + ReLU),
 * TorchScatterMaxPlugin, the concatenation [point features | per-pillar max], PFN layer 1 (Linear(192 -> 192) + BatchNorm +
 * ReLU) and the second TorchScatterMaxPlugin, of which the graph reads output 1 only (:589):
 *   voxel_features [max_pillars_num, 192] = per-pillar max of relu([h0 | max(h0)] W1^T + b1),  h0 = relu(bn0(x W0^T)).
 * None of the four per-point tensors reaches memory.  pfn0 = dsvt_small_linear_create(96, 10, W0, scale, shift);
 * pfn1 = dsvt_linear_weights_create(192, 192, W1 (BatchNorm folded), b1, DSVT_ATTN_FP32_TC).  Inputs are
 * Points2FeaturesPlugin's outputs 0 (point rows [max_points_num, 10]), 1 (point_index_in_voxel; this library's voxeliser
 * emits a pillar's rows consecutively, which the kernel relies on), 4 (pillar count) and 5 (row count).  Results = the four
 * separate launches (dsvt_small_linear_launch, dsvt_torch_scatter_max_launch, dsvt_linear_rows_concat_launch,
 * dsvt_torch_scatter_max_launch).  Rows beyond the pillar count are zero-filled when zero_tails != 0.  Batch 1.
 */
size_t dsvt_vfe_fused_workspace_size(int32_t max_points_num, int32_t max_num_points_per_voxel);
int dsvt_vfe_fused_launch(const dsvt_small_linear* pfn0, const dsvt_linear_weights* pfn1, const float* point_features,
                          const int32_t* point_index_in_voxel, const int32_t* voxel_num, const int32_t* point_num,
                          int32_t max_points_num, int32_t max_pillars_num, int32_t max_num_points_per_voxel,
                          float* voxel_features, void* workspace, size_t workspace_bytes, int32_t zero_tails,
                          dsvt_stream_t stream);

/*
 * One position-embedding MLP (fullyConnectedBnLELU_fullyConnected, src/dsvt-ai-trt.cpp:461-492; 8 call sites :603-637) in one
 * kernel:  y = relu((x2 W1^T) * scale + shift) W2^T + b2,  x2 [max_rows,2] = the in-window coordinates (WindowPartitionPlugin
 * output 5), `first` = Linear(2 -> 192) + folded BatchNorm1d, `second` = Linear(192 -> 192) created with DSVT_ATTN_FP32_TC or
 * DSVT_ATTN_FP16_GEMM.  The 192-wide hidden rows are generated inside the second layer's GEMM and never reach memory; results
 * are bit-identical to dsvt_small_linear_launch (ReLU) followed by dsvt_linear_rows_launch.
 */
int dsvt_pos_embed_mlp_launch(const dsvt_small_linear* first, const dsvt_linear_weights* second, const float* x2,
                              const int32_t* rows, int32_t max_rows, float* y, int32_t zero_tails, dsvt_stream_t stream);
/* n (<= 8) position-embedding MLPs in ONE launch -- the eight MLPs of a frame (src/dsvt-ai-trt.cpp:603-637) depend on the window
 * coordinates only.  Arrays of n handles / device pointers (host arrays); results = n calls of dsvt_pos_embed_mlp_launch. */
int dsvt_pos_embed_mlp_batch_launch(const dsvt_small_linear* const* firsts, const dsvt_linear_weights* const* seconds,
                                    const float* const* x2s, int32_t n, const int32_t* rows, int32_t max_rows,
                                    float* const* ys, int32_t zero_tails, dsvt_stream_t stream);

/* ------------------------------------------------------------------------ *
 * (next #3) TorchScatterMaxPlugin::enqueue     plugins/src/torchScatterMax.cu:282-309 (kernel :201-262)
 *           Map2BevPlugin::enqueue             plugins/src/map2bev.cu:283-312 (kernel :250-265)
 * ------------------------------------------------------------------------ */
typedef struct dsvt_torch_scatter_max_params {
    int32_t batch;
    int32_t max_points_num;            /* rows of point_features (MAX_POINTS_NUM_1 30000) */
    int32_t max_pillars_num;
    int32_t feature_num;               /* 96 / 192 (PFN layer outputs); multiple of 4, <= 256 */
    int32_t max_num_points_per_voxel;  /* row stride of point_index_in_voxel (POINTS_NUM_PER_VOXEL 48) */
    int32_t zero_tails;
} dsvt_torch_scatter_max_params;
/*
 * in : point_features [B,max_points_num,F] f32 ; point_index_in_voxel [B,max_pillars_num,npv] i32 ;
 *      point_num_in_voxel [B,max_pillars_num] i32 ; voxel_num [B] i32 ;
 *      point_num [B] i32 or NULL -- the voxeliser's row count (Points2Features output 5).  With it only the rows
 *      [point_num, max_points_num) are zero-filled; without it the whole tensor is cleared first, as the reference does.
 * out: max_point_features [B,max_points_num,F] (every point row = its pillar's channel-wise max) ;
 *      max_voxel_features [B,max_pillars_num,F]
 */
int dsvt_torch_scatter_max_launch(const dsvt_torch_scatter_max_params* p, const float* point_features,
                                  const int32_t* point_index_in_voxel, const int32_t* point_num_in_voxel,
                                  const int32_t* voxel_num, const int32_t* point_num,
                                  float* max_point_features, float* max_voxel_features, dsvt_stream_t stream);

typedef struct dsvt_map2bev_params {
    int32_t batch;
    int32_t max_pillars_num;
    int32_t channel_num;
    int32_t grid_size_x, grid_size_y;
} dsvt_map2bev_params;
/* in : voxel_features [B,max_pillars_num,C] f32 ; coords [B,max_pillars_num,4] i32 (0,0,y,x) ; voxel_num [B]
 * out: map_features [B,grid_size_y,grid_size_x,C] f32, zero where no pillar (the reference declares the shape as
 *      [B,grid_size_x,grid_size_y,C] and indexes it y-major, map2bev.cu:141-145,:263) */
int dsvt_map2bev_launch(const dsvt_map2bev_params* p, const float* voxel_features, const int32_t* coords,
                        const int32_t* voxel_num, float* map_features, dsvt_stream_t stream);

/* ------------------------------------------------------------------------ *
 * (next #4, tail) CenterHead post-process graph         src/dsvt-ai-trt.cpp:1471-1691 (TensorRT layers: sigmoid, exp,
 *                                                       TopK x2, index arithmetic, gathers, atan) -- produces the eight
 *                                                       inputs of FilterBoxByScorePlugin
 * ------------------------------------------------------------------------ */
typedef struct dsvt_center_head_params {
    int32_t batch;
    int32_t num_classes;            /* 10 */
    int32_t height, width;          /* 468 x 468 (feature-map stride 1) */
    int32_t max_top_k;              /* HM_TOP_K 500 (<= 1024) */
} dsvt_center_head_params;
size_t dsvt_center_head_topk_workspace_size(const dsvt_center_head_params* p);
/*
 * in : heatmap [B,num_classes,H,W] f32 LOGITS (hm head output before the sigmoid) ; center [B,2,H,W] ; center_z [B,1,H,W] ;
 *      dim [B,3,H,W] (log sizes) ; rot [B,2,H,W] (channel 0 = cos, 1 = sin)
 * out: scores [B,K] f32 = sigmoid of the K largest logits over all classes and cells, descending (ties: ascending flat
 *      index class*H*W + y*W + x) ; classes, xs, ys [B,K] i32 ; center [B,1,K,2] ; center_z [B,1,K,1] ; angle [B,1,K,1] =
 *      atan(sin / cos) ; dim [B,1,K,3] = exp(dim)          -- the input tensors of dsvt_filter_box_launch, in its order.
 */
int dsvt_center_head_topk_launch(const dsvt_center_head_params* p, const float* heatmap, const float* center,
                                 const float* center_z, const float* dim, const float* rot, float* scores,
                                 int32_t* classes, int32_t* xs, int32_t* ys, float* center_g, float* center_z_g,
                                 float* angle, float* dim_g, void* workspace, size_t workspace_bytes, dsvt_stream_t stream);

/* ------------------------------------------------------------------------ *
 * (next #4, tail) rotated NMS                            include/helper.h:257-283 (nms_cpu; box_overlap :166-255), run by
 *                                                       the reference on the HOST after the engine (src/dsvt-ai-trt.cpp:1954)
 * ------------------------------------------------------------------------ */
typedef struct dsvt_nms_params {
    int32_t batch;
    int32_t max_boxes;              /* rows of `boxes` per frame: max_top_k 500 (<= 1024) */
    float nms_thresh;               /* NMS_THRESH 0.01 (params.h:334) */
    int32_t zero_tails;
} dsvt_nms_params;
size_t dsvt_rotated_nms_workspace_size(const dsvt_nms_params* p);
/*
 * in : boxes [B,max_boxes,9] f32 (x,y,z,dx,dy,dz,angle,class,score) and valid [B] i32 = FilterBoxByScorePlugin's outputs
 * out: out_boxes [B,max_boxes,9] = the surviving boxes in descending score order (ties: ascending input index -- the
 *      reference's std::sort leaves ties unspecified), out_num [B], keep_index [B,max_boxes] i32 (input row of every
 *      survivor; may be NULL).  A box suppresses every lower-scored box with rotated IoU >= nms_thresh.
 */
int dsvt_rotated_nms_launch(const dsvt_nms_params* p, const float* boxes, const int32_t* valid, float* out_boxes,
                            int32_t* out_num, int32_t* keep_index, void* workspace, size_t workspace_bytes,
                            dsvt_stream_t stream);

/* standalone forms of the two gather/scatter plugins (next #2), kept for graph compatibility */
int dsvt_get_value_by_index_launch(const dsvt_set_attention_params* p, const float* x, const float* pos,
                                   const int32_t* global_index_in_set, const int32_t* set_num,
                                   float* q, float* k, float* v, dsvt_stream_t stream);
int dsvt_map_set_feature2voxel_launch(const dsvt_set_attention_params* p, const float* set_features,
                                      const int32_t* global_index_in_set, const int32_t* set_num,
                                      float* voxel_features, dsvt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DSVT_B200_H */
