// Shared declarations of the set-attention implementations.
#pragma once
#include "common.cuh"

namespace dsvt {

// device-resident, pre-arranged weights of one attention layer
struct AttnWeightsDev {
    const float* w_in_t;    // [C][3C]  in_proj_weight transposed (k-major: coalesced over output columns)
    const float* b_in;      // [3C]
    const float* w_out_t;   // [C][C]
    const float* b_out;     // [C]
    const void* tc_blob;    // operand images for the tcgen05 path (attention_tc2.cu), or nullptr
};

int set_attention_fp32(const dsvt_set_attention_params* p, const AttnWeightsDev& w, bool fused,
                       const float* q, const float* k, const float* v, const float* pos, const int* idx,
                       const float* mask, const int* set_num, const int* voxel_num, float* out, cudaStream_t st);

// FP16 tensor-core path (attention_tc2.cu)
void* attention_tc_prepare(int C, int H, const float* w_in, const float* b_in, const float* w_out, const float* b_out);
int set_attention_tc2_fused(const dsvt_set_attention_params* p, const void* tc_blob,
                            const float* x, const float* pos, const int* idx, const float* mask,
                            const int* set_num, const int* voxel_num, float* out, cudaStream_t st);

// GEMM pipeline (attention_split.cu): DSVT_ATTN_FP32_TC (split = true) and DSVT_ATTN_FP16_GEMM (split = false)
void* attention_split_prepare(const float* w_in, const float* b_in, const float* w_out, const float* b_out,
                              float* out_mul);
size_t attention_split_workspace(const dsvt_set_attention_params* p);
size_t attention_split_plan_bytes(const dsvt_set_attention_params* p);
int attention_split_plan(const dsvt_set_attention_params* p, const int* idx, const float* mask, const int* set_num,
                         void* plan, size_t plan_bytes, cudaStream_t st);
// optional epilogue of the out-projection: out = LayerNorm(attention + residual) (norm1 of the encoder layer)
struct AttnNorm { const float* residual; const float* gamma; const float* beta; float eps; };
// position embedding given as a table over the cells of a window (pos = table[cy * win_x + cx]) instead of one row per voxel
struct AttnPosTable { const int* cell; int win_x; };
int set_attention_split_fused(const dsvt_set_attention_params* p, const void* split_blob, const float* out_mul,
                              bool split, const float* x, const float* pos, const int* idx, const float* mask,
                              const int* set_num, const int* voxel_num, float* out, const void* plan,
                              void* workspace, size_t workspace_bytes, cudaStream_t st, const AttnNorm* norm = nullptr,
                              int stages = 7, const AttnPosTable* pos_table = nullptr);

// the same pipeline for pre-gathered q / k / v [B, max_sets, S, 192] (the drop-in for multHeadAttention() itself)
size_t attention_split_plugin_workspace(const dsvt_set_attention_params* p);
int set_attention_split_plugin(const dsvt_set_attention_params* p, const void* split_blob, const float* out_mul, bool split,
                               const float* q, const float* k, const float* v, const float* mask, const int* set_num,
                               float* out, void* workspace, size_t workspace_bytes, cudaStream_t st);

}  // namespace dsvt

struct dsvt_attention_weights {
    int channel_num;
    int num_heads;
    int device;
    float* blob;            // one cudaMalloc: w_in_t | b_in | w_out_t | b_out
    void* tc_blob;
    void* split_blob;       // weight images + biases of the GEMM pipeline
    float split_out_mul[4]; // per role (Q, K, V, O): power of two undoing the weight pre-scaling
    dsvt::AttnWeightsDev dev;
};
