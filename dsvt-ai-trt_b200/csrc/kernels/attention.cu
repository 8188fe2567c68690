// a3 -- C ABI of the set attention: weight preparation and precision dispatch.
#include "attention_common.cuh"
#include <vector>
#include <cstdlib>

using namespace dsvt;

extern "C" dsvt_attention_weights* dsvt_attention_weights_create(int32_t C, int32_t heads,
                                                                 const float* in_proj_weight,
                                                                 const float* in_proj_bias,
                                                                 const float* out_proj_weight,
                                                                 const float* out_proj_bias)
{
    if (C != 192 || heads != 8 || !in_proj_weight || !in_proj_bias || !out_proj_weight || !out_proj_bias) {
        set_last_error("dsvt_attention_weights_create: only channel_num=192, num_heads=8 with non-NULL weights "
                       "is supported (got C=%d heads=%d)", C, heads);
        return nullptr;
    }
    // PyTorch layout [out][in] (FC weight layout of the reference) -> k-major [in][out]
    const size_t n_in = (size_t) C * 3 * C, n_out = (size_t) C * C;
    std::vector<float> host(n_in + 3 * C + n_out + C);
    float* w_in_t = host.data();
    float* b_in = w_in_t + n_in;
    float* w_out_t = b_in + 3 * C;
    float* b_out = w_out_t + n_out;
    for (int n = 0; n < 3 * C; ++n)
        for (int k = 0; k < C; ++k) w_in_t[(size_t) k * 3 * C + n] = in_proj_weight[(size_t) n * C + k];
    for (int n = 0; n < 3 * C; ++n) b_in[n] = in_proj_bias[n];
    for (int n = 0; n < C; ++n)
        for (int k = 0; k < C; ++k) w_out_t[(size_t) k * C + n] = out_proj_weight[(size_t) n * C + k];
    for (int n = 0; n < C; ++n) b_out[n] = out_proj_bias[n];

    auto* w = new (std::nothrow) dsvt_attention_weights();
    if (!w) return nullptr;
    w->channel_num = C;
    w->num_heads = heads;
    w->tc_blob = nullptr;
    w->split_blob = nullptr;
    if (cudaGetDevice(&w->device) != cudaSuccess ||
        cudaMalloc(&w->blob, host.size() * sizeof(float)) != cudaSuccess ||
        cudaMemcpy(w->blob, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_last_error("dsvt_attention_weights_create: CUDA allocation/copy failed: %s",
                       cudaGetErrorString(cudaGetLastError()));
        delete w;
        return nullptr;
    }
    w->dev.w_in_t = w->blob;
    w->dev.b_in = w->blob + n_in;
    w->dev.w_out_t = w->blob + n_in + 3 * C;
    w->dev.b_out = w->blob + n_in + 3 * C + n_out;
    w->tc_blob = attention_tc_prepare(C, heads, in_proj_weight, in_proj_bias, out_proj_weight, out_proj_bias);
    if (!w->tc_blob) {
        set_last_error("dsvt_attention_weights_create: preparing the tensor-core operand images failed");
        cudaFree(w->blob);
        delete w;
        return nullptr;
    }
    w->dev.tc_blob = w->tc_blob;
    w->split_blob = attention_split_prepare(in_proj_weight, in_proj_bias, out_proj_weight, out_proj_bias,
                                            w->split_out_mul);
    if (!w->split_blob) {
        set_last_error("dsvt_attention_weights_create: preparing the split-precision weight images failed");
        cudaFree(w->blob);
        cudaFree(w->tc_blob);
        delete w;
        return nullptr;
    }
    return w;
}

extern "C" void dsvt_attention_weights_destroy(dsvt_attention_weights* w) {
    if (!w) return;
    cudaFree(w->blob);
    if (w->tc_blob) cudaFree(w->tc_blob);
    if (w->split_blob) cudaFree(w->split_blob);
    delete w;
}

static int attn_check(const dsvt_set_attention_params* p, const dsvt_attention_weights* w) {
    DSVT_CHECK_ARG(p != nullptr && w != nullptr, "params / weights is NULL");
    DSVT_CHECK_ARG(p->batch >= 1 && p->max_set_num >= 1, "batch / max_set_num");
    DSVT_CHECK_ARG(p->channel_num == w->channel_num && p->num_heads == w->num_heads,
                   "channel_num / num_heads do not match the weights");
    DSVT_CHECK_ARG(p->channel_num == 192 && p->num_heads == 8, "only C=192, heads=8 is built");
    return DSVT_OK;
}

extern "C" size_t dsvt_set_attention_workspace_size(const dsvt_set_attention_params* p) {
    if (p && (p->precision == DSVT_ATTN_FP32_TC || p->precision == DSVT_ATTN_FP16_GEMM))
        // max_pillars_num == 0: the plugin-shaped (q, k, v) form, one row per set slot
        return p->max_pillars_num > 0 ? attention_split_workspace(p) : attention_split_plugin_workspace(p);
    return 0;   // single-kernel paths: everything between the token tile and the output row stays on chip
}

static int attn_dispatch(const dsvt_set_attention_params* p, const dsvt_attention_weights* w, bool fused,
                         const float* q, const float* k, const float* v, const float* pos, const int* idx,
                         const float* mask, const int* set_num, const int* voxel_num, float* out, const void* plan,
                         void* workspace, size_t workspace_bytes, cudaStream_t st, const AttnNorm* norm = nullptr,
                         int stages = 7, const AttnPosTable* pos_table = nullptr)
{
    if (pos_table && !(fused && p->precision == DSVT_ATTN_FP32_TC)) {
        set_last_error("set attention: the position-embedding table form exists for the fused DSVT_ATTN_FP32_TC entry only");
        return DSVT_ERR_UNSUPPORTED;
    }
    if (stages != 7 && p->precision != DSVT_ATTN_FP32_TC && p->precision != DSVT_ATTN_FP16_GEMM) {
        set_last_error("set attention: single-stage launches exist for the GEMM-pipeline precisions only");
        return DSVT_ERR_UNSUPPORTED;
    }
    if (norm && p->precision != DSVT_ATTN_FP32_TC && p->precision != DSVT_ATTN_FP16_GEMM) {
        // single-kernel precisions have no GEMM epilogue to carry the norm: attention into `out`, then the row-wise
        // LayerNorm kernel over out + residual in place (same arithmetic, one more launch)
        if (!fused) { set_last_error("set attention + norm: fused entry points only"); return DSVT_ERR_UNSUPPORTED; }
        int rc = attn_dispatch(p, w, fused, q, k, v, pos, idx, mask, set_num, voxel_num, out, plan, workspace, workspace_bytes, st,
                               nullptr, 7);
        if (rc != DSVT_OK) return rc;
        const dsvt_layer_norm_params lp{p->batch, p->max_pillars_num, p->channel_num, norm->eps, 1};
        return dsvt_layer_norm_launch(&lp, out, norm->residual, voxel_num, norm->gamma, norm->beta, out,
                                      reinterpret_cast<dsvt_stream_t>(st));
    }
    switch (p->precision) {
        case DSVT_ATTN_FP32_TC:
        case DSVT_ATTN_FP16_GEMM:
            if (!fused)
                return set_attention_split_plugin(p, w->split_blob, w->split_out_mul, p->precision == DSVT_ATTN_FP32_TC,
                                                  q, k, v, mask, set_num, out, workspace, workspace_bytes, st);
            return set_attention_split_fused(p, w->split_blob, w->split_out_mul, p->precision == DSVT_ATTN_FP32_TC,
                                             q, pos, idx, mask, set_num, voxel_num, out, plan, workspace, workspace_bytes, st, norm,
                                             stages, pos_table);
        case DSVT_ATTN_FP32:
            // fused form: these two kernels scatter set rows into `out`; a voxel that belongs to no set (a set dropped by a
            // capacity guard upstream) must read exactly 0 like the reference's zero-filled tensor (mapSetFeature2voxel.cu:312)
            if (fused) DSVT_CUDA(cudaMemsetAsync(out, 0, (size_t) p->batch * p->max_pillars_num * p->channel_num * sizeof(float), st));
            return set_attention_fp32(p, w->dev, fused, q, k, v, pos, idx, mask, set_num, voxel_num, out, st);
        case DSVT_ATTN_FP16:
            if (!fused) {
                set_last_error("set attention: the FP16 tensor-core path is built for the fused entry point "
                               "(dsvt_set_attention_fused_launch) only");
                return DSVT_ERR_UNSUPPORTED;
            }
            DSVT_CUDA(cudaMemsetAsync(out, 0, (size_t) p->batch * p->max_pillars_num * p->channel_num * sizeof(float), st));
            return set_attention_tc2_fused(p, w->tc_blob, q, pos, idx, mask, set_num, voxel_num, out, st);
        default:
            set_last_error("set attention: precision %d is not available in this build", p->precision);
            return DSVT_ERR_UNSUPPORTED;
    }
}

extern "C" int dsvt_set_attention_launch(const dsvt_set_attention_params* p, const dsvt_attention_weights* w,
                                         const float* q, const float* k, const float* v, const float* mask,
                                         const int32_t* set_num, float* out,
                                         void* workspace, size_t workspace_bytes, dsvt_stream_t stream)
{
    int rc = attn_check(p, w);
    if (rc != DSVT_OK) return rc;
    DSVT_CHECK_ARG(q && k && v && mask && out, "NULL tensor pointer");
    DSVT_CHECK_ARG(!(((uintptr_t) q | (uintptr_t) k | (uintptr_t) v | (uintptr_t) out) & 15), "16-B alignment");
    return attn_dispatch(p, w, false, q, k, v, nullptr, nullptr, mask, set_num, nullptr, out, nullptr,
                         workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

static int fused_common(const dsvt_set_attention_params* p, const dsvt_attention_weights* w,
                        const float* x, const float* pos, const int32_t* global_index_in_set, const float* mask,
                        const int32_t* set_num, const int32_t* voxel_num, float* out, const void* plan,
                        void* workspace, size_t workspace_bytes, dsvt_stream_t stream, const AttnNorm* norm = nullptr,
                        int stages = 7, const AttnPosTable* pos_table = nullptr)
{
    int rc = attn_check(p, w);
    if (rc != DSVT_OK) return rc;
    DSVT_CHECK_ARG(x && pos && global_index_in_set && mask && set_num && voxel_num && out, "NULL tensor pointer");
    DSVT_CHECK_ARG(p->max_pillars_num >= 1, "max_pillars_num");
    DSVT_CHECK_ARG(p->axis_id == 0 || p->axis_id == 1, "axis_id");
    DSVT_CHECK_ARG(!(((uintptr_t) x | (uintptr_t) pos | (uintptr_t) out) & 15), "16-B alignment");
    return attn_dispatch(p, w, true, x, nullptr, nullptr, pos, global_index_in_set, mask, set_num, voxel_num,
                         out, plan, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream), norm, stages, pos_table);
}

extern "C" int dsvt_set_attention_fused_launch(const dsvt_set_attention_params* p, const dsvt_attention_weights* w,
                                               const float* x, const float* pos,
                                               const int32_t* global_index_in_set, const float* mask,
                                               const int32_t* set_num, const int32_t* voxel_num, float* out,
                                               void* workspace, size_t workspace_bytes, dsvt_stream_t stream)
{
    return fused_common(p, w, x, pos, global_index_in_set, mask, set_num, voxel_num, out, nullptr, workspace,
                        workspace_bytes, stream);
}

extern "C" int dsvt_set_attention_fused_planned_launch(const dsvt_set_attention_params* p,
                                                       const dsvt_attention_weights* w, const float* x, const float* pos,
                                                       const int32_t* global_index_in_set, const float* mask,
                                                       const int32_t* set_num, const int32_t* voxel_num, float* out,
                                                       const void* plan, void* workspace, size_t workspace_bytes,
                                                       dsvt_stream_t stream)
{
    return fused_common(p, w, x, pos, global_index_in_set, mask, set_num, voxel_num, out, plan, workspace,
                        workspace_bytes, stream);
}

extern "C" int dsvt_set_attention_fused_norm_launch(const dsvt_set_attention_params* p, const dsvt_attention_weights* w,
                                                    const float* x, const float* pos, const int32_t* global_index_in_set,
                                                    const float* mask, const int32_t* set_num, const int32_t* voxel_num,
                                                    const float* residual, const float* gamma, const float* beta, float eps,
                                                    float* out, const void* plan, void* workspace, size_t workspace_bytes,
                                                    dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(residual && gamma && beta, "NULL residual / gamma / beta");
    DSVT_CHECK_ARG(!(((uintptr_t) residual | (uintptr_t) gamma | (uintptr_t) beta) & 15), "16-B alignment");
    const AttnNorm norm{residual, gamma, beta, eps};
    return fused_common(p, w, x, pos, global_index_in_set, mask, set_num, voxel_num, out, plan, workspace, workspace_bytes,
                        stream, &norm);
}

extern "C" int dsvt_set_attention_fused_stages_launch(const dsvt_set_attention_params* p, const dsvt_attention_weights* w,
                                                      const float* x, const float* pos, const int32_t* global_index_in_set,
                                                      const float* mask, const int32_t* set_num, const int32_t* voxel_num,
                                                      const float* residual, const float* gamma, const float* beta, float eps,
                                                      float* out, const void* plan, void* workspace, size_t workspace_bytes,
                                                      int32_t stages, dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(stages >= 1 && stages <= 7, "stages: bit 0 QKV GEMM, bit 1 per-set core, bit 2 out-projection");
    DSVT_CHECK_ARG(plan != nullptr, "a prebuilt plan is required (the stages share it)");
    const AttnNorm norm{residual, gamma, beta, eps};
    return fused_common(p, w, x, pos, global_index_in_set, mask, set_num, voxel_num, out, plan, workspace, workspace_bytes,
                        stream, residual ? &norm : nullptr, stages);
}

extern "C" int dsvt_set_attention_fused_table_launch(const dsvt_set_attention_params* p, const dsvt_attention_weights* w,
                                                     const float* x, const float* pos_table, const int32_t* coors_in_win_2d,
                                                     int32_t win_shape_x, const int32_t* global_index_in_set, const float* mask,
                                                     const int32_t* set_num, const int32_t* voxel_num, const float* residual,
                                                     const float* gamma, const float* beta, float eps, float* out,
                                                     const void* plan, void* workspace, size_t workspace_bytes,
                                                     int32_t stages, dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(stages >= 1 && stages <= 7, "stages: bit 0 QKV projection, bit 1 per-set core, bit 2 out-projection");
    DSVT_CHECK_ARG(plan != nullptr, "a prebuilt plan is required");
    DSVT_CHECK_ARG(coors_in_win_2d && win_shape_x >= 1, "coors_in_win_2d / win_shape_x");
    const AttnNorm norm{residual, gamma, beta, eps};
    const AttnPosTable pt{coors_in_win_2d, win_shape_x};
    return fused_common(p, w, x, pos_table, global_index_in_set, mask, set_num, voxel_num, out, plan, workspace, workspace_bytes,
                        stream, residual ? &norm : nullptr, stages, &pt);
}

extern "C" size_t dsvt_set_attention_plan_size(const dsvt_set_attention_params* p) {
    return p ? attention_split_plan_bytes(p) : 0;
}

extern "C" int dsvt_set_attention_plan_launch(const dsvt_set_attention_params* p, const int32_t* global_index_in_set,
                                              const float* mask, const int32_t* set_num, void* plan, size_t plan_bytes,
                                              dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(p && global_index_in_set && mask && set_num, "NULL argument");
    DSVT_CHECK_ARG(p->batch >= 1 && p->max_set_num >= 1 && p->max_pillars_num >= 1, "batch / capacities");
    DSVT_CHECK_ARG(p->axis_id == 0 || p->axis_id == 1, "axis_id");
    return attention_split_plan(p, global_index_in_set, mask, set_num, plan, plan_bytes,
                                reinterpret_cast<cudaStream_t>(stream));
}
