// TensorRT IPluginV2DynamicExt shells of the DSVT hot path.  Each class mirrors one reference
// plugin (same registered name, version "1", creator field names, output shapes / dtypes and
// serialised byte layout) and forwards enqueue() to the C ABI in include/dsvt_b200.h.
#include "plugin_base.h"

#include <cuda_runtime_api.h>
#include <new>

namespace dsvt_plugins {

static const DataType F = DataType::kFLOAT;
static const DataType I = DataType::kINT32;

static int batch_of(const PluginTensorDesc* in) { return in[0].dims.nbDims > 0 ? in[0].dims.d[0] : 1; }

// =================================================================================================
// Points2FeaturesPlugin -- reference plugins/include/points2Features.h:18-80, src/points2Features.cu
// =================================================================================================
class Points2FeaturesPlugin final : public PluginBase {
public:
    static constexpr const char* kName = "Points2FeaturesPlugin";
    static FieldList field_list() {   // points2Features.cu:1084-1092
        return {{"max_points_num", PluginFieldType::kINT32}, {"max_points_num_voxel_filter", PluginFieldType::kINT32},
                {"max_pillars_num", PluginFieldType::kINT32}, {"point_feature_num", PluginFieldType::kINT32},
                {"feature_num", PluginFieldType::kINT32}, {"max_num_points_per_voxel", PluginFieldType::kINT32},
                {"point_cloud_range", PluginFieldType::kFLOAT32}, {"voxel_size", PluginFieldType::kFLOAT32},
                {"grid_size", PluginFieldType::kINT32}};
    }
    explicit Points2FeaturesPlugin(const dsvt_points2features_params& p) : p_(p) {}
    static IPluginV2* from_fields(const PluginFieldCollection* fc) {
        dsvt_points2features_params p{};
        p.batch = 1;
        p.max_points_num = field_int(fc, "max_points_num");
        p.max_points_num_voxel_filter = field_int(fc, "max_points_num_voxel_filter");
        p.max_pillars_num = field_int(fc, "max_pillars_num");
        p.point_feature_num = field_int(fc, "point_feature_num");
        p.feature_num = field_int(fc, "feature_num");
        p.max_num_points_per_voxel = field_int(fc, "max_num_points_per_voxel");
        // point_cloud_range = (xmin,ymin,zmin,xmax,ymax,zmax)  (points2Features.cu:1161-1168)
        p.x_min = field_float(fc, "point_cloud_range", 0); p.y_min = field_float(fc, "point_cloud_range", 1);
        p.z_min = field_float(fc, "point_cloud_range", 2); p.x_max = field_float(fc, "point_cloud_range", 3);
        p.y_max = field_float(fc, "point_cloud_range", 4); p.z_max = field_float(fc, "point_cloud_range", 5);
        p.voxel_x = field_float(fc, "voxel_size", 0); p.voxel_y = field_float(fc, "voxel_size", 1);
        p.voxel_z = field_float(fc, "voxel_size", 2);
        p.grid_x = field_int(fc, "grid_size", 0); p.grid_y = field_int(fc, "grid_size", 1);
        p.grid_z = field_int(fc, "grid_size", 2);
        p.zero_tails = 1;
        return new (std::nothrow) Points2FeaturesPlugin(p);
    }
    static IPluginV2* from_bytes(const void* data, size_t len) {   // points2Features.cu:119-143
        Reader r(data, len);
        dsvt_points2features_params p{};
        p.batch = 1;
        p.max_points_num = r.get<int>(); p.max_points_num_voxel_filter = r.get<int>(); p.max_pillars_num = r.get<int>();
        p.point_feature_num = r.get<int>(); p.feature_num = r.get<int>(); p.max_num_points_per_voxel = r.get<int>();
        p.x_min = r.get<float>(); p.x_max = r.get<float>(); p.y_min = r.get<float>(); p.y_max = r.get<float>();
        p.z_min = r.get<float>(); p.z_max = r.get<float>();
        p.voxel_x = r.get<float>(); p.voxel_y = r.get<float>(); p.voxel_z = r.get<float>();
        p.grid_x = r.get<int>(); p.grid_y = r.get<int>(); p.grid_z = r.get<int>();
        p.zero_tails = 1;
        return r.ok() ? new (std::nothrow) Points2FeaturesPlugin(p) : nullptr;
    }
    size_t getSerializationSize() const noexcept override { return 9 * sizeof(int) + 9 * sizeof(float); }  // 72 B
    void serialize(void* buf) const noexcept override {            // points2Features.cu:1037-1060
        Writer w(buf);
        w.put(p_.max_points_num); w.put(p_.max_points_num_voxel_filter); w.put(p_.max_pillars_num);
        w.put(p_.point_feature_num); w.put(p_.feature_num); w.put(p_.max_num_points_per_voxel);
        w.put(p_.x_min); w.put(p_.x_max); w.put(p_.y_min); w.put(p_.y_max); w.put(p_.z_min); w.put(p_.z_max);
        w.put(p_.voxel_x); w.put(p_.voxel_y); w.put(p_.voxel_z);
        w.put(p_.grid_x); w.put(p_.grid_y); w.put(p_.grid_z);
    }
    IPluginV2DynamicExt* clone() const noexcept override {
        auto* c = new (std::nothrow) Points2FeaturesPlugin(p_);
        if (c) c->setPluginNamespace(ns_.c_str());
        return c;
    }
    const char* getPluginType() const noexcept override { return kName; }
    int32_t getNbOutputs() const noexcept override { return 6; }
    DimsExprs getOutputDimensions(int32_t i, const DimsExprs* in, int32_t, IExprBuilder& b) noexcept override {
        const IDimensionExpr* B = in[0].d[0];                     // points2Features.cu:161-213
        switch (i) {
            case 0: return dims(b, B, {p_.max_points_num_voxel_filter, p_.feature_num});
            case 1: return dims(b, B, {p_.max_pillars_num, p_.max_num_points_per_voxel});
            case 2: return dims(b, B, {p_.max_pillars_num, 4});
            case 3: return dims(b, B, {p_.max_pillars_num, 1});
            default: return dims(b, B, {});
        }
    }
    size_t getWorkspaceSize(const PluginTensorDesc* in, int32_t, const PluginTensorDesc*, int32_t) const noexcept override {
        dsvt_points2features_params p = p_;
        p.batch = batch_of(in);
        return dsvt_points2features_workspace_size(&p);
    }
    int32_t enqueue(const PluginTensorDesc* in, const PluginTensorDesc*, const void* const* inputs,
                    void* const* outputs, void* ws, cudaStream_t stream) noexcept override {
        dsvt_points2features_params p = p_;
        p.batch = batch_of(in);
        return report(dsvt_points2features_launch(
            &p, static_cast<const float*>(inputs[0]), static_cast<const int32_t*>(inputs[1]),
            static_cast<float*>(outputs[0]), static_cast<int32_t*>(outputs[1]), static_cast<int32_t*>(outputs[2]),
            static_cast<int32_t*>(outputs[3]), static_cast<int32_t*>(outputs[4]), static_cast<int32_t*>(outputs[5]),
            ws, dsvt_points2features_workspace_size(&p), stream), kName);
    }
protected:
    const std::vector<DataType>& io_types() const override {
        static const std::vector<DataType> t{F, I, F, I, I, I, I, I};
        return t;
    }
    size_t nb_inputs() const override { return 2; }
private:
    dsvt_points2features_params p_;
};

// =================================================================================================
// WindowPartitionPlugin -- reference plugins/src/windowPartition.cu (next #1)
// =================================================================================================
class WindowPartitionPlugin final : public PluginBase {
public:
    static constexpr const char* kName = "WindowPartitionPlugin";
    static FieldList field_list() {   // windowPartition.cu:549-553
        return {{"max_win_num", PluginFieldType::kINT32}, {"max_voxel_num_per_win", PluginFieldType::kINT32},
                {"sparse_shape", PluginFieldType::kINT32}, {"win_shape", PluginFieldType::kINT32},
                {"shift_list", PluginFieldType::kINT32}};
    }
    explicit WindowPartitionPlugin(const dsvt_window_partition_params& p) : p_(p) {}
    static IPluginV2* from_fields(const PluginFieldCollection* fc) {
        dsvt_window_partition_params p{};
        p.batch = 1;
        p.max_win_num = field_int(fc, "max_win_num");
        p.max_voxel_num_per_win = field_int(fc, "max_voxel_num_per_win");
        p.sparse_shape_x = field_int(fc, "sparse_shape", 0); p.sparse_shape_y = field_int(fc, "sparse_shape", 1);
        p.sparse_shape_z = field_int(fc, "sparse_shape", 2);
        p.win_shape_x = field_int(fc, "win_shape", 0); p.win_shape_y = field_int(fc, "win_shape", 1);
        p.win_shape_z = field_int(fc, "win_shape", 2);
        p.shift_x = field_int(fc, "shift_list", 0); p.shift_y = field_int(fc, "shift_list", 1);
        p.shift_z = field_int(fc, "shift_list", 2);
        p.zero_tails = 1;
        return new (std::nothrow) WindowPartitionPlugin(p);
    }
    static IPluginV2* from_bytes(const void* data, size_t len) {   // windowPartition.cu:115-128
        Reader r(data, len);
        dsvt_window_partition_params p{};
        p.batch = 1;
        p.sparse_shape_x = r.get<int>(); p.sparse_shape_y = r.get<int>(); p.sparse_shape_z = r.get<int>();
        p.win_shape_x = r.get<int>(); p.win_shape_y = r.get<int>(); p.win_shape_z = r.get<int>();
        p.shift_x = r.get<int>(); p.shift_y = r.get<int>(); p.shift_z = r.get<int>();
        p.max_win_num = r.get<int>(); p.max_voxel_num_per_win = r.get<int>();
        p.zero_tails = 1;
        return r.ok() ? new (std::nothrow) WindowPartitionPlugin(p) : nullptr;
    }
    size_t getSerializationSize() const noexcept override { return 11 * sizeof(int); }
    void serialize(void* buf) const noexcept override {            // windowPartition.cu:511-524
        Writer w(buf);
        w.put(p_.sparse_shape_x); w.put(p_.sparse_shape_y); w.put(p_.sparse_shape_z);
        w.put(p_.win_shape_x); w.put(p_.win_shape_y); w.put(p_.win_shape_z);
        w.put(p_.shift_x); w.put(p_.shift_y); w.put(p_.shift_z);
        w.put(p_.max_win_num); w.put(p_.max_voxel_num_per_win);
    }
    IPluginV2DynamicExt* clone() const noexcept override {
        auto* c = new (std::nothrow) WindowPartitionPlugin(p_);
        if (c) c->setPluginNamespace(ns_.c_str());
        return c;
    }
    const char* getPluginType() const noexcept override { return kName; }
    int32_t getNbOutputs() const noexcept override { return 6; }
    DimsExprs getOutputDimensions(int32_t i, const DimsExprs* in, int32_t, IExprBuilder& b) noexcept override {
        const IDimensionExpr* B = in[0].d[0];
        // the reference hard-codes MAX_PILLARS_NUM for outputs 4/5 (windowPartition.cu:181-198); it equals
        // the pillar capacity of the coords input, which is what we use
        const int maxv = in[0].nbDims > 1 && in[0].d[1]->isConstant() ? in[0].d[1]->getConstantValue() : 0;
        switch (i) {
            case 0: return dims(b, B, {p_.max_win_num, p_.max_voxel_num_per_win});
            case 1: return dims(b, B, {p_.max_win_num, p_.max_voxel_num_per_win, 3});
            case 2: return dims(b, B, {p_.max_win_num});
            case 3: return dims(b, B, {});
            case 4: return dims(b, B, {maxv, 3});
            default: return dims(b, B, {maxv, 2});
        }
    }
    size_t getWorkspaceSize(const PluginTensorDesc* in, int32_t, const PluginTensorDesc*, int32_t) const noexcept override {
        dsvt_window_partition_params p = with_shapes(in);
        return dsvt_window_partition_workspace_size(&p);
    }
    int32_t enqueue(const PluginTensorDesc* in, const PluginTensorDesc*, const void* const* inputs,
                    void* const* outputs, void* ws, cudaStream_t stream) noexcept override {
        dsvt_window_partition_params p = with_shapes(in);
        return report(dsvt_window_partition_launch(
            &p, static_cast<const int32_t*>(inputs[0]), static_cast<const int32_t*>(inputs[1]),
            static_cast<int32_t*>(outputs[0]), static_cast<int32_t*>(outputs[1]), static_cast<int32_t*>(outputs[2]),
            static_cast<int32_t*>(outputs[3]), static_cast<int32_t*>(outputs[4]), static_cast<float*>(outputs[5]),
            ws, dsvt_window_partition_workspace_size(&p), stream), kName);
    }
protected:
    const std::vector<DataType>& io_types() const override {
        static const std::vector<DataType> t{I, I, I, I, I, I, I, F};
        return t;
    }
    size_t nb_inputs() const override { return 2; }
private:
    dsvt_window_partition_params with_shapes(const PluginTensorDesc* in) const {
        dsvt_window_partition_params p = p_;
        p.batch = batch_of(in);
        p.max_pillars_num = in[0].dims.d[1];
        return p;
    }
    dsvt_window_partition_params p_;
};

// =================================================================================================
// GetSetPlugin -- reference plugins/include/getSet.h:20-67, src/getSet.cu
// =================================================================================================
class GetSetPlugin final : public PluginBase {
public:
    static constexpr const char* kName = "GetSetPlugin";
    static FieldList field_list() {   // getSet.cu:782-785
        return {{"max_win_num", PluginFieldType::kINT32}, {"max_voxel_num_per_win", PluginFieldType::kINT32},
                {"voxel_num_set", PluginFieldType::kINT32}, {"win_shape", PluginFieldType::kINT32}};
    }
    explicit GetSetPlugin(const dsvt_get_set_params& p) : p_(p) {}
    static IPluginV2* from_fields(const PluginFieldCollection* fc) {
        dsvt_get_set_params p{};
        p.batch = 1;
        p.max_win_num = field_int(fc, "max_win_num");
        p.max_voxel_num_per_win = field_int(fc, "max_voxel_num_per_win");
        p.voxel_num_set = field_int(fc, "voxel_num_set");
        p.win_shape_x = field_int(fc, "win_shape", 0); p.win_shape_y = field_int(fc, "win_shape", 1);
        p.win_shape_z = field_int(fc, "win_shape", 2);
        p.num_heads = 8;   // NUM_HEADS, params.h:73
        p.zero_tails = 1;
        return new (std::nothrow) GetSetPlugin(p);
    }
    static IPluginV2* from_bytes(const void* data, size_t len) {   // getSet.cu:113-122
        Reader r(data, len);
        dsvt_get_set_params p{};
        p.batch = 1;
        p.voxel_num_set = r.get<int>(); p.max_win_num = r.get<int>(); p.max_voxel_num_per_win = r.get<int>();
        p.win_shape_x = r.get<int>(); p.win_shape_y = r.get<int>(); p.win_shape_z = r.get<int>();
        p.num_heads = 8;
        p.zero_tails = 1;
        return r.ok() ? new (std::nothrow) GetSetPlugin(p) : nullptr;
    }
    size_t getSerializationSize() const noexcept override { return 6 * sizeof(int); }
    void serialize(void* buf) const noexcept override {            // getSet.cu:749-758
        Writer w(buf);
        w.put(p_.voxel_num_set); w.put(p_.max_win_num); w.put(p_.max_voxel_num_per_win);
        w.put(p_.win_shape_x); w.put(p_.win_shape_y); w.put(p_.win_shape_z);
    }
    IPluginV2DynamicExt* clone() const noexcept override {
        auto* c = new (std::nothrow) GetSetPlugin(p_);
        if (c) c->setPluginNamespace(ns_.c_str());
        return c;
    }
    const char* getPluginType() const noexcept override { return kName; }
    int32_t getNbOutputs() const noexcept override { return 5; }
    DimsExprs getOutputDimensions(int32_t i, const DimsExprs* in, int32_t, IExprBuilder& b) noexcept override {
        const IDimensionExpr* B = in[0].d[0];                     // getSet.cu:140-184
        switch (i) {
            case 0: case 1: return dims(b, B, {2, p_.max_win_num, p_.voxel_num_set});
            case 2: return dims(b, B, {});
            default: return dims(b, B, {p_.max_win_num, p_.num_heads, p_.voxel_num_set});
        }
    }
    int32_t enqueue(const PluginTensorDesc* in, const PluginTensorDesc*, const void* const* inputs,
                    void* const* outputs, void* ws, cudaStream_t stream) noexcept override {
        dsvt_get_set_params p = p_;
        p.batch = batch_of(in);
        return report(dsvt_get_set_launch(
            &p, static_cast<const int32_t*>(inputs[0]), static_cast<const int32_t*>(inputs[1]),
            static_cast<const int32_t*>(inputs[2]), static_cast<const int32_t*>(inputs[3]),
            static_cast<int32_t*>(outputs[0]), static_cast<float*>(outputs[1]), static_cast<int32_t*>(outputs[2]),
            static_cast<float*>(outputs[3]), static_cast<float*>(outputs[4]), ws, 0, stream), kName);
    }
protected:
    const std::vector<DataType>& io_types() const override {
        static const std::vector<DataType> t{I, I, I, I, I, F, I, F, F};
        return t;
    }
    size_t nb_inputs() const override { return 4; }
private:
    dsvt_get_set_params p_;
};

// =================================================================================================
// GeluPlugin -- reference plugins/include/gelu.h:17-62, src/gelu.cu
// =================================================================================================
class GeluPlugin final : public PluginBase {
public:
    static constexpr const char* kName = "GeluPlugin";
    static FieldList field_list() {   // gelu.cu:325-326
        return {{"max_pillars_num", PluginFieldType::kINT32}, {"channel_num", PluginFieldType::kINT32}};
    }
    explicit GeluPlugin(const dsvt_gelu_params& p) : p_(p) {}
    static IPluginV2* from_fields(const PluginFieldCollection* fc) {
        dsvt_gelu_params p{1, field_int(fc, "max_pillars_num"), field_int(fc, "channel_num"), 1};
        return new (std::nothrow) GeluPlugin(p);
    }
    static IPluginV2* from_bytes(const void* data, size_t len) {
        Reader r(data, len);
        dsvt_gelu_params p{};
        p.batch = 1; p.max_pillars_num = r.get<int>(); p.channel_num = r.get<int>(); p.zero_tails = 1;
        return r.ok() ? new (std::nothrow) GeluPlugin(p) : nullptr;
    }
    size_t getSerializationSize() const noexcept override { return 2 * sizeof(int); }
    void serialize(void* buf) const noexcept override {            // gelu.cu:296-302
        Writer w(buf);
        w.put(p_.max_pillars_num); w.put(p_.channel_num);
    }
    IPluginV2DynamicExt* clone() const noexcept override {
        auto* c = new (std::nothrow) GeluPlugin(p_);
        if (c) c->setPluginNamespace(ns_.c_str());
        return c;
    }
    const char* getPluginType() const noexcept override { return kName; }
    int32_t getNbOutputs() const noexcept override { return 1; }
    DimsExprs getOutputDimensions(int32_t, const DimsExprs* in, int32_t, IExprBuilder& b) noexcept override {
        return dims(b, in[0].d[0], {p_.max_pillars_num, p_.channel_num});   // gelu.cu:149-160
    }
    int32_t enqueue(const PluginTensorDesc* in, const PluginTensorDesc*, const void* const* inputs,
                    void* const* outputs, void*, cudaStream_t stream) noexcept override {
        dsvt_gelu_params p = p_;
        p.batch = batch_of(in);
        return report(dsvt_gelu_launch(&p, static_cast<const float*>(inputs[0]), static_cast<const int32_t*>(inputs[1]),
                                       static_cast<float*>(outputs[0]), stream), kName);
    }
protected:
    const std::vector<DataType>& io_types() const override {
        static const std::vector<DataType> t{F, I, F};
        return t;
    }
    size_t nb_inputs() const override { return 2; }
private:
    dsvt_gelu_params p_;
};

// =================================================================================================
// LayerNormPlugin -- reference plugins/include/layerNorm.h:17-78, src/layerNorm.cu
// Owns host copies of gamma/beta (serialised) and a device copy uploaded in initialize().
// =================================================================================================
class LayerNormPlugin final : public PluginBase {
public:
    static constexpr const char* kName = "LayerNormPlugin";
    static FieldList field_list() {   // layerNorm.cu:494-500 -- "pes" (sic) is the advertised name of eps
        return {{"max_pillars_num", PluginFieldType::kINT32}, {"channel_num", PluginFieldType::kINT32},
                {"weights_size", PluginFieldType::kINT32}, {"pes", PluginFieldType::kFLOAT32},
                {"weights", PluginFieldType::kFLOAT32}, {"bias", PluginFieldType::kFLOAT32}};
    }
    LayerNormPlugin(int max_pillars, int channels, int wsize, float eps, const float* gamma, const float* beta)
        : max_pillars_(max_pillars), channels_(channels), wsize_(wsize), eps_(eps),
          gamma_(gamma, gamma + wsize), beta_(beta, beta + wsize) {}
    ~LayerNormPlugin() override { release(); }
    static IPluginV2* from_fields(const PluginFieldCollection* fc) {
        const int wsize = field_int(fc, "weights_size");
        const PluginField* g = find_field(fc, "weights");
        const PluginField* b = find_field(fc, "bias");
        if (wsize <= 0 || !g || !b || !g->data || !b->data) return nullptr;
        // the kernels index gamma / beta by channel: a weights_size or field length that disagrees with channel_num would
        // be read out of bounds on the host (the reference copies weights_size floats unchecked, layerNorm.cu:126-158)
        if (wsize != field_int(fc, "channel_num") || g->length != wsize || b->length != wsize) return nullptr;
        // the creator parses "eps" (layerNorm.cu:558) although it advertises "pes": the reference helper
        // therefore never sends it and eps stays 0.0.  Same behaviour here.
        return new (std::nothrow) LayerNormPlugin(field_int(fc, "max_pillars_num"), field_int(fc, "channel_num"), wsize,
                                                  field_float(fc, "eps", 0, 0.0f),
                                                  static_cast<const float*>(g->data), static_cast<const float*>(b->data));
    }
    static IPluginV2* from_bytes(const void* data, size_t len) {   // layerNorm.cu:160-190
        Reader r(data, len);
        const int mp = r.get<int>(), ch = r.get<int>(), ws = r.get<int>();
        const float eps = r.get<float>();
        if (!r.ok() || ws <= 0 || r.left() < (size_t) ws * 2 * sizeof(float)) return nullptr;
        std::vector<float> g(ws), b(ws);
        r.get_array(g.data(), ws);
        r.get_array(b.data(), ws);
        return new (std::nothrow) LayerNormPlugin(mp, ch, ws, eps, g.data(), b.data());
    }
    size_t getSerializationSize() const noexcept override {       // layerNorm.cu:446-449 (with ws == channels)
        return 3 * sizeof(int) + sizeof(float) + (size_t) wsize_ * 2 * sizeof(float);
    }
    void serialize(void* buf) const noexcept override {            // layerNorm.cu:451-471
        Writer w(buf);
        w.put(max_pillars_); w.put(channels_); w.put(wsize_); w.put(eps_);
        w.put_array(gamma_.data(), gamma_.size());
        w.put_array(beta_.data(), beta_.size());
    }
    IPluginV2DynamicExt* clone() const noexcept override {
        auto* c = new (std::nothrow) LayerNormPlugin(max_pillars_, channels_, wsize_, eps_, gamma_.data(), beta_.data());
        if (c) c->setPluginNamespace(ns_.c_str());
        return c;
    }
    int32_t initialize() noexcept override { return upload(); }
    void terminate() noexcept override { release(); }
    const char* getPluginType() const noexcept override { return kName; }
    int32_t getNbOutputs() const noexcept override { return 1; }
    DimsExprs getOutputDimensions(int32_t, const DimsExprs* in, int32_t, IExprBuilder& b) noexcept override {
        return dims(b, in[0].d[0], {max_pillars_, channels_});     // layerNorm.cu:198-209
    }
    int32_t enqueue(const PluginTensorDesc* in, const PluginTensorDesc*, const void* const* inputs,
                    void* const* outputs, void*, cudaStream_t stream) noexcept override {
        if (upload() != 0) return DSVT_ERR_CUDA;
        dsvt_layer_norm_params p{batch_of(in), max_pillars_, channels_, eps_, 1};
        return report(dsvt_layer_norm_launch(&p, static_cast<const float*>(inputs[0]), nullptr,
                                             static_cast<const int32_t*>(inputs[1]), dev_, dev_ + wsize_,
                                             static_cast<float*>(outputs[0]), stream), kName);
    }
protected:
    const std::vector<DataType>& io_types() const override {
        static const std::vector<DataType> t{F, I, F};
        return t;
    }
    size_t nb_inputs() const override { return 2; }
private:
    int upload() {
        if (dev_) return 0;
        if (cudaMalloc(reinterpret_cast<void**>(&dev_), (size_t) wsize_ * 2 * sizeof(float)) != cudaSuccess) return 1;
        if (cudaMemcpy(dev_, gamma_.data(), wsize_ * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(dev_ + wsize_, beta_.data(), wsize_ * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
            release();
            return 1;
        }
        return 0;
    }
    void release() {
        if (dev_) { cudaFree(dev_); dev_ = nullptr; }
    }
    int max_pillars_, channels_, wsize_;
    float eps_;
    std::vector<float> gamma_, beta_;
    float* dev_ = nullptr;
};

// =================================================================================================
// FilterBoxByScorePlugin -- reference plugins/include/filterBoxByScore.h:17-70, src/filterBoxByScore.cu
// =================================================================================================
class FilterBoxByScorePlugin final : public PluginBase {
public:
    static constexpr const char* kName = "FilterBoxByScorePlugin";
    static FieldList field_list() {   // filterBoxByScore.cu:459-462
        return {{"max_top_k", PluginFieldType::kINT32}, {"point_cloud_range", PluginFieldType::kFLOAT32},
                {"voxel_size", PluginFieldType::kFLOAT32}, {"score_threshold", PluginFieldType::kFLOAT32}};
    }
    explicit FilterBoxByScorePlugin(const dsvt_filter_box_params& p) : p_(p) {}
    static IPluginV2* from_fields(const PluginFieldCollection* fc) {
        dsvt_filter_box_params p{};
        p.batch = 1;
        p.max_top_k = field_int(fc, "max_top_k");
        // NOTE the order differs from Points2Features: (xmin,xmax,ymin,ymax,zmin,zmax), filterBoxByScore.cu:504-512
        p.x_min = field_float(fc, "point_cloud_range", 0); p.x_max = field_float(fc, "point_cloud_range", 1);
        p.y_min = field_float(fc, "point_cloud_range", 2); p.y_max = field_float(fc, "point_cloud_range", 3);
        p.z_min = field_float(fc, "point_cloud_range", 4); p.z_max = field_float(fc, "point_cloud_range", 5);
        p.voxel_x = field_float(fc, "voxel_size", 0); p.voxel_y = field_float(fc, "voxel_size", 1);
        p.voxel_z = field_float(fc, "voxel_size", 2);
        p.score_threshold = field_float(fc, "score_threshold");
        p.zero_tails = 1;
        return new (std::nothrow) FilterBoxByScorePlugin(p);
    }
    static IPluginV2* from_bytes(const void* data, size_t len) {   // filterBoxByScore.cu:117-135
        Reader r(data, len);
        dsvt_filter_box_params p{};
        p.batch = 1;
        p.max_top_k = r.get<int>();
        p.x_min = r.get<float>(); p.x_max = r.get<float>(); p.y_min = r.get<float>(); p.y_max = r.get<float>();
        p.z_min = r.get<float>(); p.z_max = r.get<float>();
        p.voxel_x = r.get<float>(); p.voxel_y = r.get<float>(); p.voxel_z = r.get<float>();
        p.score_threshold = r.get<float>();
        p.zero_tails = 1;
        return r.ok() ? new (std::nothrow) FilterBoxByScorePlugin(p) : nullptr;
    }
    size_t getSerializationSize() const noexcept override { return sizeof(int) + 10 * sizeof(float); }   // 44 B
    void serialize(void* buf) const noexcept override {            // filterBoxByScore.cu:418-434
        Writer w(buf);
        w.put(p_.max_top_k);
        w.put(p_.x_min); w.put(p_.x_max); w.put(p_.y_min); w.put(p_.y_max); w.put(p_.z_min); w.put(p_.z_max);
        w.put(p_.voxel_x); w.put(p_.voxel_y); w.put(p_.voxel_z);
        w.put(p_.score_threshold);
    }
    IPluginV2DynamicExt* clone() const noexcept override {
        auto* c = new (std::nothrow) FilterBoxByScorePlugin(p_);
        if (c) c->setPluginNamespace(ns_.c_str());
        return c;
    }
    const char* getPluginType() const noexcept override { return kName; }
    int32_t getNbOutputs() const noexcept override { return 2; }
    DimsExprs getOutputDimensions(int32_t i, const DimsExprs* in, int32_t, IExprBuilder& b) noexcept override {
        if (i == 0) return dims(b, in[0].d[0], {p_.max_top_k, 9});   // filterBoxByScore.cu:144-195
        return dims(b, in[0].d[0], {});
    }
    int32_t enqueue(const PluginTensorDesc* in, const PluginTensorDesc*, const void* const* inputs,
                    void* const* outputs, void*, cudaStream_t stream) noexcept override {
        dsvt_filter_box_params p = p_;
        p.batch = batch_of(in);
        return report(dsvt_filter_box_launch(
            &p, static_cast<const float*>(inputs[0]), static_cast<const int32_t*>(inputs[1]),
            static_cast<const int32_t*>(inputs[2]), static_cast<const int32_t*>(inputs[3]),
            static_cast<const float*>(inputs[4]), static_cast<const float*>(inputs[5]),
            static_cast<const float*>(inputs[6]), static_cast<const float*>(inputs[7]),
            static_cast<float*>(outputs[0]), static_cast<int32_t*>(outputs[1]), stream), kName);
    }
protected:
    const std::vector<DataType>& io_types() const override {
        static const std::vector<DataType> t{F, I, I, I, F, F, F, F, F, I};
        return t;
    }
    size_t nb_inputs() const override { return 8; }
private:
    dsvt_filter_box_params p_;
};

// =================================================================================================
// GetValueByIndexPlugin / MapSetFeature2VoxelPlugin (next #2, standalone forms)
// =================================================================================================
class GetValueByIndexPlugin final : public PluginBase {
public:
    static constexpr const char* kName = "GetValueByIndexPlugin";
    static FieldList field_list() {   // getValueByIndex.cu:427-430
        return {{"max_win_num", PluginFieldType::kINT32}, {"voxel_num_set", PluginFieldType::kINT32},
                {"channel_num", PluginFieldType::kINT32}, {"axis_id", PluginFieldType::kINT32}};
    }
    GetValueByIndexPlugin(int S, int max_win, int C, int axis) : S_(S), max_win_(max_win), C_(C), axis_(axis) {}
    static IPluginV2* from_fields(const PluginFieldCollection* fc) {
        return new (std::nothrow) GetValueByIndexPlugin(field_int(fc, "voxel_num_set"), field_int(fc, "max_win_num"),
                                                        field_int(fc, "channel_num"), field_int(fc, "axis_id"));
    }
    static IPluginV2* from_bytes(const void* data, size_t len) {
        Reader r(data, len);
        const int S = r.get<int>(), mw = r.get<int>(), C = r.get<int>(), ax = r.get<int>();
        return r.ok() ? new (std::nothrow) GetValueByIndexPlugin(S, mw, C, ax) : nullptr;
    }
    size_t getSerializationSize() const noexcept override { return 4 * sizeof(int); }
    void serialize(void* buf) const noexcept override {            // getValueByIndex.cu:396-402
        Writer w(buf);
        w.put(S_); w.put(max_win_); w.put(C_); w.put(axis_);
    }
    IPluginV2DynamicExt* clone() const noexcept override {
        auto* c = new (std::nothrow) GetValueByIndexPlugin(S_, max_win_, C_, axis_);
        if (c) c->setPluginNamespace(ns_.c_str());
        return c;
    }
    const char* getPluginType() const noexcept override { return kName; }
    int32_t getNbOutputs() const noexcept override { return 3; }
    DimsExprs getOutputDimensions(int32_t, const DimsExprs* in, int32_t, IExprBuilder& b) noexcept override {
        return dims(b, in[0].d[0], {max_win_, S_, C_});
    }
    int32_t enqueue(const PluginTensorDesc* in, const PluginTensorDesc*, const void* const* inputs,
                    void* const* outputs, void*, cudaStream_t stream) noexcept override {
        dsvt_set_attention_params p{};
        p.batch = batch_of(in); p.max_set_num = max_win_; p.voxel_num_set = S_; p.channel_num = C_; p.num_heads = 8;
        p.max_pillars_num = in[0].dims.d[1]; p.axis_id = axis_; p.zero_tails = 1;
        return report(dsvt_get_value_by_index_launch(
            &p, static_cast<const float*>(inputs[0]), static_cast<const float*>(inputs[1]),
            static_cast<const int32_t*>(inputs[2]), static_cast<const int32_t*>(inputs[3]),
            static_cast<float*>(outputs[0]), static_cast<float*>(outputs[1]), static_cast<float*>(outputs[2]), stream),
            kName);
    }
protected:
    const std::vector<DataType>& io_types() const override {
        static const std::vector<DataType> t{F, F, I, I, F, F, F};
        return t;
    }
    size_t nb_inputs() const override { return 4; }
private:
    int S_, max_win_, C_, axis_;
};

class MapSetFeature2VoxelPlugin final : public PluginBase {
public:
    static constexpr const char* kName = "MapSetFeature2VoxelPlugin";
    static FieldList field_list() {   // mapSetFeature2voxel.cu:393-397
        return {{"max_win_num", PluginFieldType::kINT32}, {"voxel_num_set", PluginFieldType::kINT32},
                {"channel_num", PluginFieldType::kINT32}, {"axis_id", PluginFieldType::kINT32},
                {"max_pillars_num", PluginFieldType::kINT32}};
    }
    MapSetFeature2VoxelPlugin(int S, int max_win, int C, int max_pillars, int axis)
        : S_(S), max_win_(max_win), C_(C), max_pillars_(max_pillars), axis_(axis) {}
    static IPluginV2* from_fields(const PluginFieldCollection* fc) {
        return new (std::nothrow) MapSetFeature2VoxelPlugin(field_int(fc, "voxel_num_set"), field_int(fc, "max_win_num"),
                                                            field_int(fc, "channel_num"), field_int(fc, "max_pillars_num"),
                                                            field_int(fc, "axis_id"));
    }
    static IPluginV2* from_bytes(const void* data, size_t len) {
        Reader r(data, len);
        const int S = r.get<int>(), mw = r.get<int>(), C = r.get<int>(), mp = r.get<int>(), ax = r.get<int>();
        return r.ok() ? new (std::nothrow) MapSetFeature2VoxelPlugin(S, mw, C, mp, ax) : nullptr;
    }
    size_t getSerializationSize() const noexcept override { return 5 * sizeof(int); }
    void serialize(void* buf) const noexcept override {            // mapSetFeature2voxel.cu:361-368
        Writer w(buf);
        w.put(S_); w.put(max_win_); w.put(C_); w.put(max_pillars_); w.put(axis_);
    }
    IPluginV2DynamicExt* clone() const noexcept override {
        auto* c = new (std::nothrow) MapSetFeature2VoxelPlugin(S_, max_win_, C_, max_pillars_, axis_);
        if (c) c->setPluginNamespace(ns_.c_str());
        return c;
    }
    const char* getPluginType() const noexcept override { return kName; }
    int32_t getNbOutputs() const noexcept override { return 1; }
    DimsExprs getOutputDimensions(int32_t, const DimsExprs* in, int32_t, IExprBuilder& b) noexcept override {
        return dims(b, in[0].d[0], {max_pillars_, C_});
    }
    int32_t enqueue(const PluginTensorDesc* in, const PluginTensorDesc*, const void* const* inputs,
                    void* const* outputs, void*, cudaStream_t stream) noexcept override {
        dsvt_set_attention_params p{};
        p.batch = batch_of(in); p.max_set_num = max_win_; p.voxel_num_set = S_; p.channel_num = C_; p.num_heads = 8;
        p.max_pillars_num = max_pillars_; p.axis_id = axis_; p.zero_tails = 1;
        return report(dsvt_map_set_feature2voxel_launch(
            &p, static_cast<const float*>(inputs[0]), static_cast<const int32_t*>(inputs[1]),
            static_cast<const int32_t*>(inputs[2]), static_cast<float*>(outputs[0]), stream), kName);
    }
protected:
    const std::vector<DataType>& io_types() const override {
        static const std::vector<DataType> t{F, I, I, F};
        return t;
    }
    size_t nb_inputs() const override { return 3; }
private:
    int S_, max_win_, C_, max_pillars_, axis_;
};

// =================================================================================================
// SetAttentionPlugin -- NEW: replaces the ~25-layer TensorRT sub-graph built by multHeadAttention()
// (src/dsvt-ai-trt.cpp:288-458).  Inputs q,k,v [B,max_sets,S,C], mask [B,max_sets,heads,S] and
// (optional 5th input) set_num [B]; output [B,max_sets,S,C].  Weights travel as creator fields in
// the PyTorch layouts the reference's loader produces (include/helper.h:367-433).
// Serialised: max_set_num, voxel_num_set, channel_num, num_heads, precision (5 x i32) then
// in_proj_weight[3C*C], in_proj_bias[3C], out_proj_weight[C*C], out_proj_bias[C] (f32).
// =================================================================================================
class SetAttentionPlugin final : public PluginBase {
public:
    static constexpr const char* kName = "SetAttentionPlugin";
    static FieldList field_list() {
        return {{"max_win_num", PluginFieldType::kINT32}, {"voxel_num_set", PluginFieldType::kINT32},
                {"channel_num", PluginFieldType::kINT32}, {"num_heads", PluginFieldType::kINT32},
                {"precision", PluginFieldType::kINT32},
                {"in_proj_weight", PluginFieldType::kFLOAT32}, {"in_proj_bias", PluginFieldType::kFLOAT32},
                {"out_proj_weight", PluginFieldType::kFLOAT32}, {"out_proj_bias", PluginFieldType::kFLOAT32}};
    }
    SetAttentionPlugin(int max_sets, int S, int C, int heads, int precision, const float* w_in, const float* b_in,
                       const float* w_out, const float* b_out)
        : max_sets_(max_sets), S_(S), C_(C), heads_(heads), precision_(precision),
          w_in_(w_in, w_in + (size_t) 3 * C * C), b_in_(b_in, b_in + 3 * C),
          w_out_(w_out, w_out + (size_t) C * C), b_out_(b_out, b_out + C) {}
    ~SetAttentionPlugin() override { release(); }
    static IPluginV2* from_fields(const PluginFieldCollection* fc) {
        const PluginField* wi = find_field(fc, "in_proj_weight");
        const PluginField* bi = find_field(fc, "in_proj_bias");
        const PluginField* wo = find_field(fc, "out_proj_weight");
        const PluginField* bo = find_field(fc, "out_proj_bias");
        const int C = field_int(fc, "channel_num");
        if (C <= 0 || C > 4096 || !wi || !bi || !wo || !bo || !wi->data || !bi->data || !wo->data || !bo->data) return nullptr;
        if (wi->length != 3 * C * C || bi->length != 3 * C || wo->length != C * C || bo->length != C) return nullptr;
        return new (std::nothrow) SetAttentionPlugin(
            field_int(fc, "max_win_num"), field_int(fc, "voxel_num_set"), C, field_int(fc, "num_heads", 0, 8),
            field_int(fc, "precision", 0, DSVT_ATTN_FP32), static_cast<const float*>(wi->data),
            static_cast<const float*>(bi->data), static_cast<const float*>(wo->data), static_cast<const float*>(bo->data));
    }
    static IPluginV2* from_bytes(const void* data, size_t len) {
        Reader r(data, len);
        const int ms = r.get<int>(), S = r.get<int>(), C = r.get<int>(), H = r.get<int>(), prec = r.get<int>();
        const int nb_in = r.get<int>();          // 4 or 5: whether the optional set_num input is connected
        if (!r.ok() || C <= 0 || C > 4096 || r.left() < ((size_t) 4 * C * C + 4 * C) * sizeof(float)) return nullptr;
        std::vector<float> wi((size_t) 3 * C * C), bi(3 * C), wo((size_t) C * C), bo(C);
        r.get_array(wi.data(), wi.size()); r.get_array(bi.data(), bi.size());
        r.get_array(wo.data(), wo.size()); r.get_array(bo.data(), bo.size());
        auto* pl = new (std::nothrow) SetAttentionPlugin(ms, S, C, H, prec, wi.data(), bi.data(), wo.data(), bo.data());
        if (pl) pl->nb_inputs_seen_ = nb_in == 5 ? 5 : 4;
        return pl;
    }
    size_t getSerializationSize() const noexcept override {
        return 6 * sizeof(int) + ((size_t) 4 * C_ * C_ + 4 * C_) * sizeof(float);
    }
    void serialize(void* buf) const noexcept override {
        Writer w(buf);
        w.put(max_sets_); w.put(S_); w.put(C_); w.put(heads_); w.put(precision_); w.put(nb_inputs_seen_);
        w.put_array(w_in_.data(), w_in_.size()); w.put_array(b_in_.data(), b_in_.size());
        w.put_array(w_out_.data(), w_out_.size()); w.put_array(b_out_.data(), b_out_.size());
    }
    IPluginV2DynamicExt* clone() const noexcept override {
        auto* c = new (std::nothrow) SetAttentionPlugin(max_sets_, S_, C_, heads_, precision_, w_in_.data(),
                                                        b_in_.data(), w_out_.data(), b_out_.data());
        if (c) { c->setPluginNamespace(ns_.c_str()); c->nb_inputs_seen_ = nb_inputs_seen_; }
        return c;
    }
    int32_t initialize() noexcept override { return upload(); }
    void terminate() noexcept override { release(); }
    const char* getPluginType() const noexcept override { return kName; }
    int32_t getNbOutputs() const noexcept override { return 1; }
    DimsExprs getOutputDimensions(int32_t, const DimsExprs* in, int32_t, IExprBuilder& b) noexcept override {
        return dims(b, in[0].d[0], {max_sets_, S_, C_});
    }
    bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* io, int32_t nbIn, int32_t nbOut) noexcept override {
        if (pos < 0 || pos >= nbIn + nbOut || io[pos].format != TensorFormat::kLINEAR) return false;
        if (nbIn == 5 && pos == 4) return io[pos].type == I;      // optional set_num
        return io[pos].type == F;
    }
    DataType getOutputDataType(int32_t, const DataType*, int32_t) const noexcept override { return F; }
    int32_t enqueue(const PluginTensorDesc* in, const PluginTensorDesc*, const void* const* inputs,
                    void* const* outputs, void* ws, cudaStream_t stream) noexcept override {
        if (upload() != 0) return DSVT_ERR_CUDA;
        dsvt_set_attention_params p{};
        p.batch = batch_of(in); p.max_set_num = max_sets_; p.voxel_num_set = S_; p.channel_num = C_;
        p.num_heads = heads_; p.precision = precision_; p.zero_tails = 1;
        // a 5th input (set_num) is detected by the mask descriptor being followed by a rank-1 int tensor;
        // TensorRT passes nbInputs only to the build-time calls, so the shell records it in configurePlugin
        const int32_t* set_num = nb_inputs_seen_ == 5 ? static_cast<const int32_t*>(inputs[4]) : nullptr;
        return report(dsvt_set_attention_launch(&p, dev_, static_cast<const float*>(inputs[0]),
                                                static_cast<const float*>(inputs[1]), static_cast<const float*>(inputs[2]),
                                                static_cast<const float*>(inputs[3]), set_num,
                                                static_cast<float*>(outputs[0]), ws, dsvt_set_attention_workspace_size(&p),
                                                stream), kName);
    }
    // tensor-core precisions (DSVT_ATTN_FP32_TC / DSVT_ATTN_FP16_GEMM): projections of the set slots + per-set core need
    // FP32 intermediates; 0 bytes for the CUDA-core precision
    size_t getWorkspaceSize(const PluginTensorDesc* in, int32_t, const PluginTensorDesc*, int32_t) const noexcept override {
        dsvt_set_attention_params p{};
        p.batch = batch_of(in); p.max_set_num = max_sets_; p.voxel_num_set = S_; p.channel_num = C_;
        p.num_heads = heads_; p.precision = precision_; p.zero_tails = 1;
        return dsvt_set_attention_workspace_size(&p);
    }
    void configurePlugin(const DynamicPluginTensorDesc*, int32_t nbInputs, const DynamicPluginTensorDesc*,
                         int32_t) noexcept override { nb_inputs_seen_ = nbInputs; }
    void set_nb_inputs(int n) { nb_inputs_seen_ = n; }
protected:
    const std::vector<DataType>& io_types() const override {
        static const std::vector<DataType> t{F, F, F, F, F};
        return t;
    }
    size_t nb_inputs() const override { return 4; }
private:
    int upload() {
        if (dev_) return 0;
        dev_ = dsvt_attention_weights_create(C_, heads_, w_in_.data(), b_in_.data(), w_out_.data(), b_out_.data());
        return dev_ ? 0 : 1;
    }
    void release() {
        if (dev_) { dsvt_attention_weights_destroy(dev_); dev_ = nullptr; }
    }
    int max_sets_, S_, C_, heads_, precision_;
    std::vector<float> w_in_, b_in_, w_out_, b_out_;
    dsvt_attention_weights* dev_ = nullptr;
    int nb_inputs_seen_ = 4;
};

// SetAttentionFusedPlugin -- one node for GetValueByIndexPlugin -> multHeadAttention() -> MapSetFeature2VoxelPlugin
// (src/dsvt-ai-trt.cpp:653-668 and the 7 sibling call sites).  This is the form every tensor-core
// precision is built for.
// Inputs : x [B,max_pillars,C] f32, pos [B,max_pillars,C] f32, global_index_in_set [B,2,max_sets,S] i32,
//          mask [B,max_sets,heads,S] f32, set_num [B] i32, voxel_num [B] i32  (+ optional 7th input: the attention
//          plan of this (partition, axis) from SetAttentionPlanPlugin -- without it the plan is rebuilt on every
//          enqueue).   Output: [B,max_pillars,C] f32.
// Fields : SetAttentionPlugin's + max_pillars_num, axis_id.  Serialised: 7 x i32 then the four weight arrays.
class SetAttentionFusedPlugin final : public PluginBase {
public:
    static constexpr const char* kName = "SetAttentionFusedPlugin";
    static FieldList field_list() {
        return {{"max_win_num", PluginFieldType::kINT32}, {"voxel_num_set", PluginFieldType::kINT32},
                {"channel_num", PluginFieldType::kINT32}, {"num_heads", PluginFieldType::kINT32},
                {"precision", PluginFieldType::kINT32}, {"max_pillars_num", PluginFieldType::kINT32},
                {"axis_id", PluginFieldType::kINT32},
                {"in_proj_weight", PluginFieldType::kFLOAT32}, {"in_proj_bias", PluginFieldType::kFLOAT32},
                {"out_proj_weight", PluginFieldType::kFLOAT32}, {"out_proj_bias", PluginFieldType::kFLOAT32},
                // optional: the encoder layer's norm1 -- output = LayerNorm(attention + x) (src/dsvt-ai-trt.cpp:669-676)
                {"norm_weight", PluginFieldType::kFLOAT32}, {"norm_bias", PluginFieldType::kFLOAT32},
                {"norm_eps", PluginFieldType::kFLOAT32}};
    }
    SetAttentionFusedPlugin(int max_sets, int S, int C, int heads, int precision, int max_pillars, int axis,
                            const float* w_in, const float* b_in, const float* w_out, const float* b_out,
                            const float* norm_w = nullptr, const float* norm_b = nullptr, float norm_eps = 0.f)
        : max_sets_(max_sets), S_(S), C_(C), heads_(heads), precision_(precision), max_pillars_(max_pillars), axis_(axis),
          w_in_(w_in, w_in + (size_t) 3 * C * C), b_in_(b_in, b_in + 3 * C),
          w_out_(w_out, w_out + (size_t) C * C), b_out_(b_out, b_out + C), norm_eps_(norm_eps) {
        if (norm_w && norm_b) { norm_w_.assign(norm_w, norm_w + C); norm_b_.assign(norm_b, norm_b + C); }
    }
    ~SetAttentionFusedPlugin() override { release(); }
    static IPluginV2* from_fields(const PluginFieldCollection* fc) {
        const PluginField* wi = find_field(fc, "in_proj_weight");
        const PluginField* bi = find_field(fc, "in_proj_bias");
        const PluginField* wo = find_field(fc, "out_proj_weight");
        const PluginField* bo = find_field(fc, "out_proj_bias");
        const int C = field_int(fc, "channel_num");
        if (C <= 0 || C > 4096 || !wi || !bi || !wo || !bo || !wi->data || !bi->data || !wo->data || !bo->data) return nullptr;
        if (wi->length != 3 * C * C || bi->length != 3 * C || wo->length != C * C || bo->length != C) return nullptr;
        const PluginField* nw = find_field(fc, "norm_weight");
        const PluginField* nb = find_field(fc, "norm_bias");
        const bool norm = nw && nb && nw->data && nb->data;
        if (norm && (nw->length != C || nb->length != C)) return nullptr;
        return new (std::nothrow) SetAttentionFusedPlugin(
            field_int(fc, "max_win_num"), field_int(fc, "voxel_num_set"), C, field_int(fc, "num_heads", 0, 8),
            field_int(fc, "precision", 0, DSVT_ATTN_FP32_TC), field_int(fc, "max_pillars_num"), field_int(fc, "axis_id"),
            static_cast<const float*>(wi->data), static_cast<const float*>(bi->data),
            static_cast<const float*>(wo->data), static_cast<const float*>(bo->data),
            norm ? static_cast<const float*>(nw->data) : nullptr, norm ? static_cast<const float*>(nb->data) : nullptr,
            field_float(fc, "norm_eps", 0, 0.0f));
    }
    static IPluginV2* from_bytes(const void* data, size_t len) {
        Reader r(data, len);
        const int ms = r.get<int>(), S = r.get<int>(), C = r.get<int>(), H = r.get<int>(), prec = r.get<int>();
        const int mp = r.get<int>(), axis = r.get<int>();
        if (!r.ok() || C <= 0 || C > 4096 || r.left() < ((size_t) 4 * C * C + 4 * C) * sizeof(float)) return nullptr;
        std::vector<float> wi((size_t) 3 * C * C), bi(3 * C), wo((size_t) C * C), bo(C);
        r.get_array(wi.data(), wi.size()); r.get_array(bi.data(), bi.size());
        r.get_array(wo.data(), wo.size()); r.get_array(bo.data(), bo.size());
        // trailer (absent in blobs written before the norm epilogue existed): has_norm, eps, gamma[C], beta[C]
        std::vector<float> nw, nb;
        float eps = 0.f;
        if (r.left() >= sizeof(int) + sizeof(float)) {
            const int has = r.get<int>();
            eps = r.get<float>();
            if (has) {
                if (r.left() < (size_t) 2 * C * sizeof(float)) return nullptr;
                nw.resize(C); nb.resize(C);
                r.get_array(nw.data(), C); r.get_array(nb.data(), C);
            }
        }
        return new (std::nothrow) SetAttentionFusedPlugin(ms, S, C, H, prec, mp, axis, wi.data(), bi.data(), wo.data(), bo.data(),
                                                          nw.empty() ? nullptr : nw.data(), nb.empty() ? nullptr : nb.data(), eps);
    }
    size_t getSerializationSize() const noexcept override {
        return 7 * sizeof(int) + ((size_t) 4 * C_ * C_ + 4 * C_) * sizeof(float) + sizeof(int) + sizeof(float) +
               (norm_w_.size() + norm_b_.size()) * sizeof(float);
    }
    void serialize(void* buf) const noexcept override {
        Writer w(buf);
        w.put(max_sets_); w.put(S_); w.put(C_); w.put(heads_); w.put(precision_); w.put(max_pillars_); w.put(axis_);
        w.put_array(w_in_.data(), w_in_.size()); w.put_array(b_in_.data(), b_in_.size());
        w.put_array(w_out_.data(), w_out_.size()); w.put_array(b_out_.data(), b_out_.size());
        const int has = norm_w_.empty() ? 0 : 1;
        w.put(has); w.put(norm_eps_);
        if (has) { w.put_array(norm_w_.data(), norm_w_.size()); w.put_array(norm_b_.data(), norm_b_.size()); }
    }
    IPluginV2DynamicExt* clone() const noexcept override {
        auto* c = new (std::nothrow) SetAttentionFusedPlugin(max_sets_, S_, C_, heads_, precision_, max_pillars_, axis_,
                                                             w_in_.data(), b_in_.data(), w_out_.data(), b_out_.data(),
                                                             norm_w_.empty() ? nullptr : norm_w_.data(),
                                                             norm_b_.empty() ? nullptr : norm_b_.data(), norm_eps_);
        if (c) { c->setPluginNamespace(ns_.c_str()); c->nb_inputs_seen_ = nb_inputs_seen_; }
        return c;
    }
    int32_t initialize() noexcept override { return upload(); }
    void terminate() noexcept override { release(); }
    const char* getPluginType() const noexcept override { return kName; }
    int32_t getNbOutputs() const noexcept override { return 1; }
    DimsExprs getOutputDimensions(int32_t, const DimsExprs* in, int32_t, IExprBuilder& b) noexcept override {
        return dims(b, in[0].d[0], {max_pillars_, C_});
    }
    bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* io, int32_t nbIn, int32_t nbOut) noexcept override {
        if (pos < 0 || pos >= nbIn + nbOut || io[pos].format != TensorFormat::kLINEAR) return false;
        const bool is_int = pos == 2 || pos == 4 || pos == 5 || (nbIn == 7 && pos == 6);
        return io[pos].type == (is_int ? I : F);
    }
    void configurePlugin(const DynamicPluginTensorDesc*, int32_t nbInputs, const DynamicPluginTensorDesc*,
                         int32_t) noexcept override { nb_inputs_seen_ = nbInputs; }
    DataType getOutputDataType(int32_t, const DataType*, int32_t) const noexcept override { return F; }
    size_t getWorkspaceSize(const PluginTensorDesc* in, int32_t, const PluginTensorDesc*, int32_t) const noexcept override {
        const dsvt_set_attention_params p = params(batch_of(in));
        return dsvt_set_attention_workspace_size(&p);
    }
    int32_t enqueue(const PluginTensorDesc* in, const PluginTensorDesc*, const void* const* inputs,
                    void* const* outputs, void* ws, cudaStream_t stream) noexcept override {
        if (upload() != 0) return DSVT_ERR_CUDA;
        const dsvt_set_attention_params p = params(batch_of(in));
        const void* plan = nb_inputs_seen_ == 7 ? inputs[6] : nullptr;
        if (norm_dev_)          // norm1 folded into the out-projection's epilogue: residual = x (input 0)
            return report(dsvt_set_attention_fused_norm_launch(
                              &p, dev_, static_cast<const float*>(inputs[0]), static_cast<const float*>(inputs[1]),
                              static_cast<const int32_t*>(inputs[2]), static_cast<const float*>(inputs[3]),
                              static_cast<const int32_t*>(inputs[4]), static_cast<const int32_t*>(inputs[5]),
                              static_cast<const float*>(inputs[0]), norm_dev_, norm_dev_ + C_, norm_eps_,
                              static_cast<float*>(outputs[0]), plan, ws, dsvt_set_attention_workspace_size(&p), stream), kName);
        return report(dsvt_set_attention_fused_planned_launch(
                          &p, dev_, static_cast<const float*>(inputs[0]), static_cast<const float*>(inputs[1]),
                          static_cast<const int32_t*>(inputs[2]), static_cast<const float*>(inputs[3]),
                          static_cast<const int32_t*>(inputs[4]), static_cast<const int32_t*>(inputs[5]),
                          static_cast<float*>(outputs[0]), plan, ws, dsvt_set_attention_workspace_size(&p), stream), kName);
    }
protected:
    const std::vector<DataType>& io_types() const override {
        static const std::vector<DataType> t{F, F, I, F, I, I, F};
        return t;
    }
    size_t nb_inputs() const override { return 6; }
private:
    dsvt_set_attention_params params(int batch) const {
        dsvt_set_attention_params p{};
        p.batch = batch; p.max_set_num = max_sets_; p.voxel_num_set = S_; p.channel_num = C_; p.num_heads = heads_;
        p.max_pillars_num = max_pillars_; p.axis_id = axis_; p.precision = precision_; p.zero_tails = 1;
        return p;
    }
    int upload() {
        if (!dev_) dev_ = dsvt_attention_weights_create(C_, heads_, w_in_.data(), b_in_.data(), w_out_.data(), b_out_.data());
        if (!dev_) return 1;
        if (!norm_w_.empty() && !norm_dev_) {
            if (cudaMalloc(reinterpret_cast<void**>(&norm_dev_), (size_t) 2 * C_ * sizeof(float)) != cudaSuccess) return 1;
            if (cudaMemcpy(norm_dev_, norm_w_.data(), C_ * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
                cudaMemcpy(norm_dev_ + C_, norm_b_.data(), C_ * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return 1;
        }
        return 0;
    }
    void release() {
        if (dev_) { dsvt_attention_weights_destroy(dev_); dev_ = nullptr; }
        if (norm_dev_) { cudaFree(norm_dev_); norm_dev_ = nullptr; }
    }
    int max_sets_, S_, C_, heads_, precision_, max_pillars_, axis_;
    std::vector<float> w_in_, b_in_, w_out_, b_out_, norm_w_, norm_b_;
    float norm_eps_ = 0.f;
    dsvt_attention_weights* dev_ = nullptr;
    float* norm_dev_ = nullptr;
    int nb_inputs_seen_ = 6;
};

// SetAttentionPlanPlugin -- the set partition of one (window partition, axis) in token order, built ONCE from the
// GetSetPlugin outputs and fed to the SetAttentionFusedPlugin nodes of the layers that share the partition (2 of the 8
// in the reference graph each).  Inputs: global_index_in_set [B,2,max_sets,S] i32, mask [B,max_sets,heads,S] f32,
// set_num [B] i32.  Output: plan [B, plan_ints] i32 (opaque).  Fields / serialised: max_win_num, voxel_num_set,
// num_heads, max_pillars_num, axis_id (5 x i32).
class SetAttentionPlanPlugin final : public PluginBase {
public:
    static constexpr const char* kName = "SetAttentionPlanPlugin";
    static FieldList field_list() {
        return {{"max_win_num", PluginFieldType::kINT32}, {"voxel_num_set", PluginFieldType::kINT32},
                {"num_heads", PluginFieldType::kINT32}, {"max_pillars_num", PluginFieldType::kINT32},
                {"axis_id", PluginFieldType::kINT32}};
    }
    SetAttentionPlanPlugin(int max_sets, int S, int heads, int max_pillars, int axis)
        : max_sets_(max_sets), S_(S), heads_(heads), max_pillars_(max_pillars), axis_(axis) {}
    static IPluginV2* from_fields(const PluginFieldCollection* fc) {
        return new (std::nothrow) SetAttentionPlanPlugin(field_int(fc, "max_win_num"), field_int(fc, "voxel_num_set"),
                                                         field_int(fc, "num_heads", 0, 8), field_int(fc, "max_pillars_num"),
                                                         field_int(fc, "axis_id"));
    }
    static IPluginV2* from_bytes(const void* data, size_t len) {
        Reader r(data, len);
        const int a = r.get<int>(), b = r.get<int>(), c = r.get<int>(), d = r.get<int>(), e = r.get<int>();
        return r.ok() ? new (std::nothrow) SetAttentionPlanPlugin(a, b, c, d, e) : nullptr;
    }
    size_t getSerializationSize() const noexcept override { return 5 * sizeof(int); }
    void serialize(void* buf) const noexcept override {
        Writer w(buf);
        w.put(max_sets_); w.put(S_); w.put(heads_); w.put(max_pillars_); w.put(axis_);
    }
    IPluginV2DynamicExt* clone() const noexcept override {
        auto* c = new (std::nothrow) SetAttentionPlanPlugin(max_sets_, S_, heads_, max_pillars_, axis_);
        if (c) c->setPluginNamespace(ns_.c_str());
        return c;
    }
    const char* getPluginType() const noexcept override { return kName; }
    int32_t getNbOutputs() const noexcept override { return 1; }
    DimsExprs getOutputDimensions(int32_t, const DimsExprs* in, int32_t, IExprBuilder& b) noexcept override {
        const dsvt_set_attention_params p = params(1);
        return dims(b, in[0].d[0], {(int) (dsvt_set_attention_plan_size(&p) / sizeof(int32_t))});
    }
    int32_t enqueue(const PluginTensorDesc* in, const PluginTensorDesc*, const void* const* inputs,
                    void* const* outputs, void*, cudaStream_t stream) noexcept override {
        const dsvt_set_attention_params p = params(batch_of(in));
        return report(dsvt_set_attention_plan_launch(&p, static_cast<const int32_t*>(inputs[0]),
                                                     static_cast<const float*>(inputs[1]), static_cast<const int32_t*>(inputs[2]),
                                                     outputs[0], dsvt_set_attention_plan_size(&p), stream), kName);
    }
protected:
    const std::vector<DataType>& io_types() const override {
        static const std::vector<DataType> t{I, F, I, I};
        return t;
    }
    size_t nb_inputs() const override { return 3; }
private:
    dsvt_set_attention_params params(int batch) const {
        dsvt_set_attention_params p{};
        p.batch = batch; p.max_set_num = max_sets_; p.voxel_num_set = S_; p.channel_num = 192; p.num_heads = heads_;
        p.max_pillars_num = max_pillars_; p.axis_id = axis_; p.precision = DSVT_ATTN_FP32_TC; p.zero_tails = 1;
        return p;
    }
    int max_sets_, S_, heads_, max_pillars_, axis_;
};

// =================================================================================================
// TorchScatterMaxPlugin  (reference plugins/src/torchScatterMax.cu; fields :388-390, serialised 3 x i32 :346-352)
// Inputs : point_features [B,max_points,F] f32, point_index_in_voxel [B,max_pillars,npv] i32,
//          point_num_in_voxel [B,max_pillars] i32, voxel_num [B] i32  (+ optional 5th input: point_num [B] i32, the
//          voxeliser's row count -- lets the plugin zero-fill only the unused rows instead of the whole tensor)
// Outputs: max_point_features [B,max_points,F], max_voxel_features [B,max_pillars,F]  (:116-141)
class TorchScatterMaxPlugin final : public PluginBase {
public:
    static constexpr const char* kName = "TorchScatterMaxPlugin";
    static FieldList field_list() {
        return {{"max_points_num", PluginFieldType::kINT32}, {"max_pillars_num", PluginFieldType::kINT32},
                {"feature_num", PluginFieldType::kINT32}};
    }
    TorchScatterMaxPlugin(int max_points, int max_pillars, int feature_num)
        : max_points_(max_points), max_pillars_(max_pillars), feature_num_(feature_num) {}
    static IPluginV2* from_fields(const PluginFieldCollection* fc) {
        return new (std::nothrow) TorchScatterMaxPlugin(field_int(fc, "max_points_num"), field_int(fc, "max_pillars_num"),
                                                        field_int(fc, "feature_num"));
    }
    static IPluginV2* from_bytes(const void* data, size_t len) {
        Reader r(data, len);
        const int a = r.get<int>(), b = r.get<int>(), c = r.get<int>();
        return r.ok() ? new (std::nothrow) TorchScatterMaxPlugin(a, b, c) : nullptr;
    }
    size_t getSerializationSize() const noexcept override { return 3 * sizeof(int); }
    void serialize(void* buf) const noexcept override {
        Writer w(buf);
        w.put(max_points_); w.put(max_pillars_); w.put(feature_num_);
    }
    IPluginV2DynamicExt* clone() const noexcept override {
        auto* c = new (std::nothrow) TorchScatterMaxPlugin(max_points_, max_pillars_, feature_num_);
        if (c) { c->setPluginNamespace(ns_.c_str()); c->nb_inputs_seen_ = nb_inputs_seen_; }
        return c;
    }
    const char* getPluginType() const noexcept override { return kName; }
    int32_t getNbOutputs() const noexcept override { return 2; }
    DimsExprs getOutputDimensions(int32_t index, const DimsExprs* in, int32_t, IExprBuilder& b) noexcept override {
        return dims(b, in[0].d[0], {index == 0 ? max_points_ : max_pillars_, feature_num_});
    }
    bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* io, int32_t nbIn, int32_t nbOut) noexcept override {
        if (pos < 0 || pos >= nbIn + nbOut || io[pos].format != TensorFormat::kLINEAR) return false;
        return io[pos].type == ((pos >= 1 && pos < nbIn) ? I : F);
    }
    DataType getOutputDataType(int32_t, const DataType*, int32_t) const noexcept override { return F; }
    void configurePlugin(const DynamicPluginTensorDesc*, int32_t nbInputs, const DynamicPluginTensorDesc*,
                         int32_t) noexcept override { nb_inputs_seen_ = nbInputs; }
    int32_t enqueue(const PluginTensorDesc* in, const PluginTensorDesc*, const void* const* inputs,
                    void* const* outputs, void*, cudaStream_t stream) noexcept override {
        dsvt_torch_scatter_max_params p{};
        p.batch = batch_of(in); p.max_points_num = max_points_; p.max_pillars_num = max_pillars_;
        p.feature_num = feature_num_; p.zero_tails = 1;
        p.max_num_points_per_voxel = in[1].dims.nbDims >= 3 ? in[1].dims.d[2] : 48;    // POINTS_NUM_PER_VOXEL
        const int32_t* point_num = nb_inputs_seen_ == 5 ? static_cast<const int32_t*>(inputs[4]) : nullptr;
        return report(dsvt_torch_scatter_max_launch(&p, static_cast<const float*>(inputs[0]),
                                                    static_cast<const int32_t*>(inputs[1]), static_cast<const int32_t*>(inputs[2]),
                                                    static_cast<const int32_t*>(inputs[3]), point_num,
                                                    static_cast<float*>(outputs[0]), static_cast<float*>(outputs[1]), stream), kName);
    }
protected:
    const std::vector<DataType>& io_types() const override {
        static const std::vector<DataType> t{F, I, I, I, F, F};
        return t;
    }
    size_t nb_inputs() const override { return 4; }
private:
    int max_points_, max_pillars_, feature_num_;
    int nb_inputs_seen_ = 4;
};

// =================================================================================================
// Map2BevPlugin  (reference plugins/src/map2bev.cu; fields :383-386, serialised 4 x i32 :352-359)
// Inputs : voxel_features [B,max_pillars,C] f32, coords [B,max_pillars,4] i32, valid_voxel_num [B] i32
// Output : map [B,grid_size_x,grid_size_y,C] f32 as the reference declares it (:141-145); indexed y-major (:263)
class Map2BevPlugin final : public PluginBase {
public:
    static constexpr const char* kName = "Map2BevPlugin";
    static FieldList field_list() {
        return {{"max_pillars_num", PluginFieldType::kINT32}, {"channel_num", PluginFieldType::kINT32},
                {"grid_size_x", PluginFieldType::kINT32}, {"grid_size_y", PluginFieldType::kINT32}};
    }
    explicit Map2BevPlugin(const dsvt_map2bev_params& p) : p_(p) {}
    static IPluginV2* from_fields(const PluginFieldCollection* fc) {
        dsvt_map2bev_params p{1, field_int(fc, "max_pillars_num"), field_int(fc, "channel_num"),
                              field_int(fc, "grid_size_x"), field_int(fc, "grid_size_y")};
        return new (std::nothrow) Map2BevPlugin(p);
    }
    static IPluginV2* from_bytes(const void* data, size_t len) {
        Reader r(data, len);
        dsvt_map2bev_params p{};
        p.batch = 1; p.max_pillars_num = r.get<int>(); p.channel_num = r.get<int>();
        p.grid_size_x = r.get<int>(); p.grid_size_y = r.get<int>();
        return r.ok() ? new (std::nothrow) Map2BevPlugin(p) : nullptr;
    }
    size_t getSerializationSize() const noexcept override { return 4 * sizeof(int); }
    void serialize(void* buf) const noexcept override {
        Writer w(buf);
        w.put(p_.max_pillars_num); w.put(p_.channel_num); w.put(p_.grid_size_x); w.put(p_.grid_size_y);
    }
    IPluginV2DynamicExt* clone() const noexcept override {
        auto* c = new (std::nothrow) Map2BevPlugin(p_);
        if (c) c->setPluginNamespace(ns_.c_str());
        return c;
    }
    const char* getPluginType() const noexcept override { return kName; }
    int32_t getNbOutputs() const noexcept override { return 1; }
    DimsExprs getOutputDimensions(int32_t, const DimsExprs* in, int32_t, IExprBuilder& b) noexcept override {
        return dims(b, in[0].d[0], {p_.grid_size_x, p_.grid_size_y, p_.channel_num});
    }
    int32_t enqueue(const PluginTensorDesc* in, const PluginTensorDesc*, const void* const* inputs,
                    void* const* outputs, void*, cudaStream_t stream) noexcept override {
        dsvt_map2bev_params p = p_;
        p.batch = batch_of(in);
        return report(dsvt_map2bev_launch(&p, static_cast<const float*>(inputs[0]), static_cast<const int32_t*>(inputs[1]),
                                          static_cast<const int32_t*>(inputs[2]), static_cast<float*>(outputs[0]), stream), kName);
    }
protected:
    const std::vector<DataType>& io_types() const override {
        static const std::vector<DataType> t{F, I, I, F};
        return t;
    }
    size_t nb_inputs() const override { return 3; }
private:
    dsvt_map2bev_params p_;
};

// =================================================================================================
// LayerNormChainPlugin (new) -- n = 1..3 consecutive (addElementWise(kSUM) + LayerNormPlugin) pairs of the reference graph
// (src/dsvt-ai-trt.cpp:685-697, :750-756) as ONE node: y = LN_n(... LN_1(x + r_1) ... + r_n).
// Inputs : x [B,max_pillars,C] f32, voxel_num [B] i32, r_1 .. r_n [B,max_pillars,C] f32.  Output [B,max_pillars,C] f32.
// Fields : max_pillars_num, channel_num, n_stages (i32), eps (f32), weights f32[n*C], bias f32[n*C] (stage-major).
// Serialised: 3 x i32, f32, then gamma[n*C], beta[n*C].  Same arithmetic as n LayerNormPlugin nodes (bit-identical).
// =================================================================================================
class LayerNormChainPlugin final : public PluginBase {
public:
    static constexpr const char* kName = "LayerNormChainPlugin";
    static FieldList field_list() {
        return {{"max_pillars_num", PluginFieldType::kINT32}, {"channel_num", PluginFieldType::kINT32},
                {"n_stages", PluginFieldType::kINT32}, {"eps", PluginFieldType::kFLOAT32},
                {"weights", PluginFieldType::kFLOAT32}, {"bias", PluginFieldType::kFLOAT32}};
    }
    LayerNormChainPlugin(int max_pillars, int channels, int n, float eps, const float* gamma, const float* beta)
        : max_pillars_(max_pillars), channels_(channels), n_(n), eps_(eps),
          gamma_(gamma, gamma + (size_t) n * channels), beta_(beta, beta + (size_t) n * channels) {}
    ~LayerNormChainPlugin() override { release(); }
    static IPluginV2* from_fields(const PluginFieldCollection* fc) {
        const int C = field_int(fc, "channel_num"), n = field_int(fc, "n_stages");
        const PluginField* g = find_field(fc, "weights");
        const PluginField* b = find_field(fc, "bias");
        if (C <= 0 || C > 65536 || n < 1 || n > 3 || !g || !b || !g->data || !b->data) return nullptr;
        if (g->length != n * C || b->length != n * C) return nullptr;
        return new (std::nothrow) LayerNormChainPlugin(field_int(fc, "max_pillars_num"), C, n, field_float(fc, "eps", 0, 0.0f),
                                                       static_cast<const float*>(g->data), static_cast<const float*>(b->data));
    }
    static IPluginV2* from_bytes(const void* data, size_t len) {
        Reader r(data, len);
        const int mp = r.get<int>(), C = r.get<int>(), n = r.get<int>();
        const float eps = r.get<float>();
        if (!r.ok() || C <= 0 || C > 65536 || n < 1 || n > 3 || r.left() < (size_t) 2 * n * C * sizeof(float)) return nullptr;
        std::vector<float> g((size_t) n * C), b((size_t) n * C);
        r.get_array(g.data(), g.size());
        r.get_array(b.data(), b.size());
        return new (std::nothrow) LayerNormChainPlugin(mp, C, n, eps, g.data(), b.data());
    }
    size_t getSerializationSize() const noexcept override { return 3 * sizeof(int) + sizeof(float) + (gamma_.size() + beta_.size()) * sizeof(float); }
    void serialize(void* buf) const noexcept override {
        Writer w(buf);
        w.put(max_pillars_); w.put(channels_); w.put(n_); w.put(eps_);
        w.put_array(gamma_.data(), gamma_.size());
        w.put_array(beta_.data(), beta_.size());
    }
    IPluginV2DynamicExt* clone() const noexcept override {
        auto* c = new (std::nothrow) LayerNormChainPlugin(max_pillars_, channels_, n_, eps_, gamma_.data(), beta_.data());
        if (c) c->setPluginNamespace(ns_.c_str());
        return c;
    }
    int32_t initialize() noexcept override { return upload(); }
    void terminate() noexcept override { release(); }
    const char* getPluginType() const noexcept override { return kName; }
    int32_t getNbOutputs() const noexcept override { return 1; }
    DimsExprs getOutputDimensions(int32_t, const DimsExprs* in, int32_t, IExprBuilder& b) noexcept override {
        return dims(b, in[0].d[0], {max_pillars_, channels_});
    }
    bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* io, int32_t nbIn, int32_t nbOut) noexcept override {
        if (pos < 0 || pos >= nbIn + nbOut || nbIn != 2 + n_ || io[pos].format != TensorFormat::kLINEAR) return false;
        return io[pos].type == (pos == 1 ? I : F);
    }
    DataType getOutputDataType(int32_t, const DataType*, int32_t) const noexcept override { return F; }
    int32_t enqueue(const PluginTensorDesc* in, const PluginTensorDesc*, const void* const* inputs,
                    void* const* outputs, void*, cudaStream_t stream) noexcept override {
        if (upload() != 0) return DSVT_ERR_CUDA;
        dsvt_layer_norm_params p{batch_of(in), max_pillars_, channels_, eps_, 1};
        dsvt_ln_stage st[3];
        for (int s = 0; s < n_; ++s)
            st[s] = dsvt_ln_stage{static_cast<const float*>(inputs[2 + s]), dev_ + (size_t) s * channels_,
                                  dev_ + (size_t) (n_ + s) * channels_};
        return report(dsvt_layer_norm_chain_launch(&p, static_cast<const float*>(inputs[0]), static_cast<const int32_t*>(inputs[1]),
                                                   st, n_, static_cast<float*>(outputs[0]), stream), kName);
    }
protected:
    const std::vector<DataType>& io_types() const override {
        static const std::vector<DataType> t{F, I, F, F, F, F};
        return t;
    }
    size_t nb_inputs() const override { return 2 + (size_t) n_; }
private:
    int upload() {
        if (dev_) return 0;
        const size_t n = gamma_.size();
        if (cudaMalloc(reinterpret_cast<void**>(&dev_), 2 * n * sizeof(float)) != cudaSuccess) return 1;
        if (cudaMemcpy(dev_, gamma_.data(), n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(dev_ + n, beta_.data(), n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) { release(); return 1; }
        return 0;
    }
    void release() { if (dev_) { cudaFree(dev_); dev_ = nullptr; } }
    int max_pillars_, channels_, n_;
    float eps_;
    std::vector<float> gamma_, beta_;
    float* dev_ = nullptr;
};

// =================================================================================================
// LinearPlugin (new) -- one TensorRT FullyConnected layer of the 3-D backbone with what follows it in the reference graph
// folded into the same node (fullyConnected_gelu_fullyConnected src/dsvt-ai-trt.cpp:494-529, the kSUM + LayerNormPlugin
// pairs :685-697 / :750-756):  y = act(x W^T + b), then optionally n_stages x (add residual_s, LayerNorm_s).
// Inputs : x [B,max_rows,K] f32, rows [B] i32 (valid row count, e.g. Points2Features output 4), residual_1..n [B,max_rows,N].
// Output : [B,max_rows,N] f32, rows beyond the count zero.
// Fields : max_rows, in_features (K), out_features (N), activation (0 none, 1 GELU, 2 ReLU), precision (DSVT_ATTN_FP32_TC or
//          DSVT_ATTN_FP16_GEMM), weight f32[N*K], bias f32[N], n_stages (0..3; needs N == 192, activation 0), ln_eps,
//          ln_weights f32[n*N], ln_bias f32[n*N].
// =================================================================================================
class LinearPlugin final : public PluginBase {
public:
    static constexpr const char* kName = "LinearPlugin";
    static FieldList field_list() {
        return {{"max_rows", PluginFieldType::kINT32}, {"in_features", PluginFieldType::kINT32},
                {"out_features", PluginFieldType::kINT32}, {"activation", PluginFieldType::kINT32},
                {"precision", PluginFieldType::kINT32}, {"weight", PluginFieldType::kFLOAT32},
                {"bias", PluginFieldType::kFLOAT32}, {"n_stages", PluginFieldType::kINT32},
                {"ln_eps", PluginFieldType::kFLOAT32}, {"ln_weights", PluginFieldType::kFLOAT32},
                {"ln_bias", PluginFieldType::kFLOAT32}};
    }
    LinearPlugin(int max_rows, int K, int N, int act, int precision, const float* W, const float* b, int n_ln, float eps,
                 const float* gamma, const float* beta)
        : max_rows_(max_rows), K_(K), N_(N), act_(act), precision_(precision), n_ln_(n_ln), eps_(eps),
          w_(W, W + (size_t) N * K), b_(b, b + N) {
        if (n_ln > 0) { gamma_.assign(gamma, gamma + (size_t) n_ln * N); beta_.assign(beta, beta + (size_t) n_ln * N); }
    }
    ~LinearPlugin() override { release(); }
    static bool valid(int K, int N, int act, int n_ln) {
        return K > 0 && N > 0 && K % 192 == 0 && N % 192 == 0 && K <= 768 && N <= 768 && act >= 0 && act <= 2 && n_ln >= 0 &&
               n_ln <= 3 && (n_ln == 0 || (N == 192 && act == 0 && K <= 384));
    }
    static IPluginV2* from_fields(const PluginFieldCollection* fc) {
        const int K = field_int(fc, "in_features"), N = field_int(fc, "out_features"), act = field_int(fc, "activation");
        const int n_ln = field_int(fc, "n_stages");
        const PluginField* w = find_field(fc, "weight");
        const PluginField* b = find_field(fc, "bias");
        const PluginField* g = find_field(fc, "ln_weights");
        const PluginField* be = find_field(fc, "ln_bias");
        if (!valid(K, N, act, n_ln) || !w || !b || !w->data || !b->data || w->length != N * K || b->length != N) return nullptr;
        if (n_ln > 0 && (!g || !be || !g->data || !be->data || g->length != n_ln * N || be->length != n_ln * N)) return nullptr;
        return new (std::nothrow) LinearPlugin(field_int(fc, "max_rows"), K, N, act, field_int(fc, "precision", 0, DSVT_ATTN_FP32_TC),
                                               static_cast<const float*>(w->data), static_cast<const float*>(b->data), n_ln,
                                               field_float(fc, "ln_eps", 0, 0.0f), n_ln ? static_cast<const float*>(g->data) : nullptr,
                                               n_ln ? static_cast<const float*>(be->data) : nullptr);
    }
    static IPluginV2* from_bytes(const void* data, size_t len) {
        Reader r(data, len);
        const int mr = r.get<int>(), K = r.get<int>(), N = r.get<int>(), act = r.get<int>(), prec = r.get<int>(), n_ln = r.get<int>();
        const float eps = r.get<float>();
        if (!r.ok() || !valid(K, N, act, n_ln) || r.left() < ((size_t) N * K + N + (size_t) 2 * n_ln * N) * sizeof(float)) return nullptr;
        std::vector<float> w((size_t) N * K), b(N), g((size_t) n_ln * N), be((size_t) n_ln * N);
        r.get_array(w.data(), w.size()); r.get_array(b.data(), b.size());
        r.get_array(g.data(), g.size()); r.get_array(be.data(), be.size());
        return new (std::nothrow) LinearPlugin(mr, K, N, act, prec, w.data(), b.data(), n_ln, eps, g.data(), be.data());
    }
    size_t getSerializationSize() const noexcept override {
        return 6 * sizeof(int) + sizeof(float) + (w_.size() + b_.size() + gamma_.size() + beta_.size()) * sizeof(float);
    }
    void serialize(void* buf) const noexcept override {
        Writer w(buf);
        w.put(max_rows_); w.put(K_); w.put(N_); w.put(act_); w.put(precision_); w.put(n_ln_); w.put(eps_);
        w.put_array(w_.data(), w_.size()); w.put_array(b_.data(), b_.size());
        w.put_array(gamma_.data(), gamma_.size()); w.put_array(beta_.data(), beta_.size());
    }
    IPluginV2DynamicExt* clone() const noexcept override {
        auto* c = new (std::nothrow) LinearPlugin(max_rows_, K_, N_, act_, precision_, w_.data(), b_.data(), n_ln_, eps_,
                                                  gamma_.data(), beta_.data());
        if (c) c->setPluginNamespace(ns_.c_str());
        return c;
    }
    int32_t initialize() noexcept override { return upload(); }
    void terminate() noexcept override { release(); }
    const char* getPluginType() const noexcept override { return kName; }
    int32_t getNbOutputs() const noexcept override { return 1; }
    DimsExprs getOutputDimensions(int32_t, const DimsExprs* in, int32_t, IExprBuilder& b) noexcept override {
        return dims(b, in[0].d[0], {max_rows_, N_});
    }
    bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* io, int32_t nbIn, int32_t nbOut) noexcept override {
        if (pos < 0 || pos >= nbIn + nbOut || nbIn != 2 + n_ln_ || io[pos].format != TensorFormat::kLINEAR) return false;
        return io[pos].type == (pos == 1 ? I : F);
    }
    DataType getOutputDataType(int32_t, const DataType*, int32_t) const noexcept override { return F; }
    int32_t enqueue(const PluginTensorDesc* in, const PluginTensorDesc*, const void* const* inputs,
                    void* const* outputs, void*, cudaStream_t stream) noexcept override {
        if (upload() != 0) return DSVT_ERR_CUDA;
        if (batch_of(in) != 1) { std::fprintf(stderr, "[dsvt_b200] LinearPlugin: batch 1 only\n"); return DSVT_ERR_UNSUPPORTED; }
        const float* x = static_cast<const float*>(inputs[0]);
        const int32_t* rows = static_cast<const int32_t*>(inputs[1]);
        float* y = static_cast<float*>(outputs[0]);
        if (n_ln_ == 0) return report(dsvt_linear_rows_launch(dev_, x, rows, max_rows_, act_, y, 1, stream), kName);
        dsvt_ln_stage st[3];
        for (int s = 0; s < n_ln_; ++s)
            st[s] = dsvt_ln_stage{static_cast<const float*>(inputs[2 + s]), ln_dev_ + (size_t) s * N_, ln_dev_ + (size_t) (n_ln_ + s) * N_};
        return report(dsvt_linear_rows_norm_launch(dev_, x, rows, max_rows_, st, n_ln_, eps_, y, 1, stream), kName);
    }
protected:
    const std::vector<DataType>& io_types() const override {
        static const std::vector<DataType> t{F, I, F, F, F, F};
        return t;
    }
    size_t nb_inputs() const override { return 2 + (size_t) n_ln_; }
private:
    int upload() {
        if (!dev_) dev_ = dsvt_linear_weights_create(N_, K_, w_.data(), b_.data(), precision_);
        if (!dev_) return 1;
        if (n_ln_ > 0 && !ln_dev_) {
            const size_t n = gamma_.size();
            if (cudaMalloc(reinterpret_cast<void**>(&ln_dev_), 2 * n * sizeof(float)) != cudaSuccess) return 1;
            if (cudaMemcpy(ln_dev_, gamma_.data(), n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
                cudaMemcpy(ln_dev_ + n, beta_.data(), n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return 1;
        }
        return 0;
    }
    void release() {
        if (dev_) { dsvt_linear_weights_destroy(dev_); dev_ = nullptr; }
        if (ln_dev_) { cudaFree(ln_dev_); ln_dev_ = nullptr; }
    }
    int max_rows_, K_, N_, act_, precision_, n_ln_;
    float eps_;
    std::vector<float> w_, b_, gamma_, beta_;
    dsvt_linear_weights* dev_ = nullptr;
    float* ln_dev_ = nullptr;
};

// =================================================================================================
// FfnFusedPlugin (new) -- fullyConnected_gelu_fullyConnected (src/dsvt-ai-trt.cpp:494-529: FullyConnected 192 -> 384, GeluPlugin,
// FullyConnected 384 -> 192) and the n_stages (kSUM + LayerNormPlugin) pairs behind it (:685-697, :750-756) as ONE node / ONE
// kernel (dsvt_ffn_fused_launch): the 384-wide hidden rows stay in tensor memory.
// Inputs : x [B,max_rows,192] f32, rows [B] i32, residual_1..n [B,max_rows,192].   Output: [B,max_rows,192], rows beyond the count zero.
// Fields : max_rows, weight1 f32[384*192], bias1 f32[384], weight2 f32[192*384], bias2 f32[192], n_stages (1..3), ln_eps,
//          ln_weights f32[n*192], ln_bias f32[n*192].
// =================================================================================================
class FfnFusedPlugin final : public PluginBase {
public:
    static constexpr const char* kName = "FfnFusedPlugin";
    static constexpr int kCin = 192, kHid = 384;
    static FieldList field_list() {
        return {{"max_rows", PluginFieldType::kINT32}, {"weight1", PluginFieldType::kFLOAT32}, {"bias1", PluginFieldType::kFLOAT32},
                {"weight2", PluginFieldType::kFLOAT32}, {"bias2", PluginFieldType::kFLOAT32}, {"n_stages", PluginFieldType::kINT32},
                {"ln_eps", PluginFieldType::kFLOAT32}, {"ln_weights", PluginFieldType::kFLOAT32}, {"ln_bias", PluginFieldType::kFLOAT32}};
    }
    FfnFusedPlugin(int max_rows, const float* w1, const float* b1, const float* w2, const float* b2, int n_ln, float eps,
                   const float* gamma, const float* beta)
        : max_rows_(max_rows), n_ln_(n_ln), eps_(eps), w1_(w1, w1 + kHid * kCin), b1_(b1, b1 + kHid), w2_(w2, w2 + kCin * kHid),
          b2_(b2, b2 + kCin), gamma_(gamma, gamma + (size_t) n_ln * kCin), beta_(beta, beta + (size_t) n_ln * kCin) {}
    ~FfnFusedPlugin() override { release(); }
    static IPluginV2* from_fields(const PluginFieldCollection* fc) {
        const int n_ln = field_int(fc, "n_stages");
        const PluginField *w1 = find_field(fc, "weight1"), *b1 = find_field(fc, "bias1"), *w2 = find_field(fc, "weight2"),
                          *b2 = find_field(fc, "bias2"), *g = find_field(fc, "ln_weights"), *be = find_field(fc, "ln_bias");
        if (n_ln < 1 || n_ln > 3 || !w1 || !b1 || !w2 || !b2 || !g || !be || !w1->data || !b1->data || !w2->data || !b2->data ||
            !g->data || !be->data || w1->length != kHid * kCin || b1->length != kHid || w2->length != kCin * kHid ||
            b2->length != kCin || g->length != n_ln * kCin || be->length != n_ln * kCin)
            return nullptr;
        return new (std::nothrow) FfnFusedPlugin(field_int(fc, "max_rows"), static_cast<const float*>(w1->data),
                                                 static_cast<const float*>(b1->data), static_cast<const float*>(w2->data),
                                                 static_cast<const float*>(b2->data), n_ln, field_float(fc, "ln_eps", 0, 0.0f),
                                                 static_cast<const float*>(g->data), static_cast<const float*>(be->data));
    }
    static IPluginV2* from_bytes(const void* data, size_t len) {
        Reader r(data, len);
        const int mr = r.get<int>(), n_ln = r.get<int>();
        const float eps = r.get<float>();
        const size_t nw = (size_t) kHid * kCin;
        if (!r.ok() || n_ln < 1 || n_ln > 3 || r.left() < (2 * nw + kHid + kCin + (size_t) 2 * n_ln * kCin) * sizeof(float)) return nullptr;
        std::vector<float> w1(nw), b1(kHid), w2(nw), b2(kCin), g((size_t) n_ln * kCin), be((size_t) n_ln * kCin);
        r.get_array(w1.data(), w1.size()); r.get_array(b1.data(), b1.size()); r.get_array(w2.data(), w2.size());
        r.get_array(b2.data(), b2.size()); r.get_array(g.data(), g.size()); r.get_array(be.data(), be.size());
        return new (std::nothrow) FfnFusedPlugin(mr, w1.data(), b1.data(), w2.data(), b2.data(), n_ln, eps, g.data(), be.data());
    }
    size_t getSerializationSize() const noexcept override {
        return 2 * sizeof(int) + sizeof(float) +
               (w1_.size() + b1_.size() + w2_.size() + b2_.size() + gamma_.size() + beta_.size()) * sizeof(float);
    }
    void serialize(void* buf) const noexcept override {
        Writer w(buf);
        w.put(max_rows_); w.put(n_ln_); w.put(eps_);
        w.put_array(w1_.data(), w1_.size()); w.put_array(b1_.data(), b1_.size()); w.put_array(w2_.data(), w2_.size());
        w.put_array(b2_.data(), b2_.size()); w.put_array(gamma_.data(), gamma_.size()); w.put_array(beta_.data(), beta_.size());
    }
    IPluginV2DynamicExt* clone() const noexcept override {
        auto* c = new (std::nothrow) FfnFusedPlugin(max_rows_, w1_.data(), b1_.data(), w2_.data(), b2_.data(), n_ln_, eps_,
                                                    gamma_.data(), beta_.data());
        if (c) c->setPluginNamespace(ns_.c_str());
        return c;
    }
    int32_t initialize() noexcept override { return upload(); }
    void terminate() noexcept override { release(); }
    const char* getPluginType() const noexcept override { return kName; }
    int32_t getNbOutputs() const noexcept override { return 1; }
    DimsExprs getOutputDimensions(int32_t, const DimsExprs* in, int32_t, IExprBuilder& b) noexcept override {
        return dims(b, in[0].d[0], {max_rows_, kCin});
    }
    bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* io, int32_t nbIn, int32_t nbOut) noexcept override {
        if (pos < 0 || pos >= nbIn + nbOut || nbIn != 2 + n_ln_ || io[pos].format != TensorFormat::kLINEAR) return false;
        return io[pos].type == (pos == 1 ? I : F);
    }
    DataType getOutputDataType(int32_t, const DataType*, int32_t) const noexcept override { return F; }
    int32_t enqueue(const PluginTensorDesc* in, const PluginTensorDesc*, const void* const* inputs,
                    void* const* outputs, void*, cudaStream_t stream) noexcept override {
        if (upload() != 0) return DSVT_ERR_CUDA;
        if (batch_of(in) != 1) { std::fprintf(stderr, "[dsvt_b200] FfnFusedPlugin: batch 1 only\n"); return DSVT_ERR_UNSUPPORTED; }
        dsvt_ln_stage st[3];
        for (int s = 0; s < n_ln_; ++s)
            st[s] = dsvt_ln_stage{static_cast<const float*>(inputs[2 + s]), ln_dev_ + (size_t) s * kCin, ln_dev_ + (size_t) (n_ln_ + s) * kCin};
        return report(dsvt_ffn_fused_launch(fc1_, fc2_, static_cast<const float*>(inputs[0]), static_cast<const int32_t*>(inputs[1]),
                                            max_rows_, st, n_ln_, eps_, static_cast<float*>(outputs[0]), 1, stream), kName);
    }
protected:
    const std::vector<DataType>& io_types() const override {
        static const std::vector<DataType> t{F, I, F, F, F, F};
        return t;
    }
    size_t nb_inputs() const override { return 2 + (size_t) n_ln_; }
private:
    int upload() {
        if (!fc1_) fc1_ = dsvt_linear_weights_create(kHid, kCin, w1_.data(), b1_.data(), DSVT_ATTN_FP32_TC);
        if (!fc2_) fc2_ = dsvt_linear_weights_create(kCin, kHid, w2_.data(), b2_.data(), DSVT_ATTN_FP32_TC);
        if (!fc1_ || !fc2_) return 1;
        if (!ln_dev_) {
            const size_t n = gamma_.size();
            if (cudaMalloc(reinterpret_cast<void**>(&ln_dev_), 2 * n * sizeof(float)) != cudaSuccess) return 1;
            if (cudaMemcpy(ln_dev_, gamma_.data(), n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
                cudaMemcpy(ln_dev_ + n, beta_.data(), n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return 1;
        }
        return 0;
    }
    void release() {
        if (fc1_) { dsvt_linear_weights_destroy(fc1_); fc1_ = nullptr; }
        if (fc2_) { dsvt_linear_weights_destroy(fc2_); fc2_ = nullptr; }
        if (ln_dev_) { cudaFree(ln_dev_); ln_dev_ = nullptr; }
    }
    int max_rows_, n_ln_;
    float eps_;
    std::vector<float> w1_, b1_, w2_, b2_, gamma_, beta_;
    dsvt_linear_weights *fc1_ = nullptr, *fc2_ = nullptr;
    float* ln_dev_ = nullptr;
};

// =================================================================================================
// VfeFusedPlugin (new) -- the pillar feature net of src/dsvt-ai-trt.cpp:571-590 as ONE node (dsvt_vfe_fused_launch): PFN layer 0
// (FullyConnected 10 -> 96 + Scale + ReLU), TorchScatterMaxPlugin, concatenation, PFN layer 1 (FullyConnected 192 -> 192 + Scale +
// ReLU, BatchNorm folded into weight1 / bias1 by the builder) and the second TorchScatterMaxPlugin's output 1.
// Inputs : Points2FeaturesPlugin outputs 0 (point rows [B,max_points,10]), 1 (point_index_in_voxel [B,max_pillars,npv] i32),
//          4 (pillar count [B] i32), 5 (row count [B] i32).       Output: voxel features [B,max_pillars,192] f32.
// Fields : max_points_num, max_pillars_num, max_num_points_per_voxel, weight0 f32[96*10], scale0 f32[96], shift0 f32[96],
//          weight1 f32[192*192], bias1 f32[192].
// =================================================================================================
class VfeFusedPlugin final : public PluginBase {
public:
    static constexpr const char* kName = "VfeFusedPlugin";
    static constexpr int kIn = 10, kMid = 96, kOut = 192;
    static FieldList field_list() {
        return {{"max_points_num", PluginFieldType::kINT32}, {"max_pillars_num", PluginFieldType::kINT32},
                {"max_num_points_per_voxel", PluginFieldType::kINT32}, {"weight0", PluginFieldType::kFLOAT32},
                {"scale0", PluginFieldType::kFLOAT32}, {"shift0", PluginFieldType::kFLOAT32}, {"weight1", PluginFieldType::kFLOAT32},
                {"bias1", PluginFieldType::kFLOAT32}};
    }
    VfeFusedPlugin(int max_points, int max_pillars, int npv, const float* w0, const float* sc0, const float* sh0, const float* w1,
                   const float* b1)
        : max_points_(max_points), max_pillars_(max_pillars), npv_(npv), w0_(w0, w0 + kMid * kIn), sc0_(sc0, sc0 + kMid),
          sh0_(sh0, sh0 + kMid), w1_(w1, w1 + kOut * kOut), b1_(b1, b1 + kOut) {}
    ~VfeFusedPlugin() override { release(); }
    static bool valid(int mp, int mv, int npv) { return mp >= 1 && mv >= 1 && npv >= 1 && npv <= 64; }
    static IPluginV2* from_fields(const PluginFieldCollection* fc) {
        const int mp = field_int(fc, "max_points_num"), mv = field_int(fc, "max_pillars_num"), npv = field_int(fc, "max_num_points_per_voxel");
        const PluginField *w0 = find_field(fc, "weight0"), *sc = find_field(fc, "scale0"), *sh = find_field(fc, "shift0"),
                          *w1 = find_field(fc, "weight1"), *b1 = find_field(fc, "bias1");
        if (!valid(mp, mv, npv) || !w0 || !sc || !sh || !w1 || !b1 || !w0->data || !sc->data || !sh->data || !w1->data || !b1->data ||
            w0->length != kMid * kIn || sc->length != kMid || sh->length != kMid || w1->length != kOut * kOut || b1->length != kOut)
            return nullptr;
        return new (std::nothrow) VfeFusedPlugin(mp, mv, npv, static_cast<const float*>(w0->data), static_cast<const float*>(sc->data),
                                                 static_cast<const float*>(sh->data), static_cast<const float*>(w1->data),
                                                 static_cast<const float*>(b1->data));
    }
    static IPluginV2* from_bytes(const void* data, size_t len) {
        Reader r(data, len);
        const int mp = r.get<int>(), mv = r.get<int>(), npv = r.get<int>();
        if (!r.ok() || !valid(mp, mv, npv) || r.left() < ((size_t) kMid * kIn + 2 * kMid + (size_t) kOut * kOut + kOut) * sizeof(float)) return nullptr;
        std::vector<float> w0((size_t) kMid * kIn), sc(kMid), sh(kMid), w1((size_t) kOut * kOut), b1(kOut);
        r.get_array(w0.data(), w0.size()); r.get_array(sc.data(), sc.size()); r.get_array(sh.data(), sh.size());
        r.get_array(w1.data(), w1.size()); r.get_array(b1.data(), b1.size());
        return new (std::nothrow) VfeFusedPlugin(mp, mv, npv, w0.data(), sc.data(), sh.data(), w1.data(), b1.data());
    }
    size_t getSerializationSize() const noexcept override {
        return 3 * sizeof(int) + (w0_.size() + sc0_.size() + sh0_.size() + w1_.size() + b1_.size()) * sizeof(float);
    }
    void serialize(void* buf) const noexcept override {
        Writer w(buf);
        w.put(max_points_); w.put(max_pillars_); w.put(npv_);
        w.put_array(w0_.data(), w0_.size()); w.put_array(sc0_.data(), sc0_.size()); w.put_array(sh0_.data(), sh0_.size());
        w.put_array(w1_.data(), w1_.size()); w.put_array(b1_.data(), b1_.size());
    }
    IPluginV2DynamicExt* clone() const noexcept override {
        auto* c = new (std::nothrow) VfeFusedPlugin(max_points_, max_pillars_, npv_, w0_.data(), sc0_.data(), sh0_.data(), w1_.data(), b1_.data());
        if (c) c->setPluginNamespace(ns_.c_str());
        return c;
    }
    int32_t initialize() noexcept override { return upload(); }
    void terminate() noexcept override { release(); }
    const char* getPluginType() const noexcept override { return kName; }
    int32_t getNbOutputs() const noexcept override { return 1; }
    DimsExprs getOutputDimensions(int32_t, const DimsExprs* in, int32_t, IExprBuilder& b) noexcept override {
        return dims(b, in[0].d[0], {max_pillars_, kOut});
    }
    size_t getWorkspaceSize(const PluginTensorDesc*, int32_t, const PluginTensorDesc*, int32_t) const noexcept override {
        return dsvt_vfe_fused_workspace_size(max_points_, npv_);
    }
    int32_t enqueue(const PluginTensorDesc* in, const PluginTensorDesc*, const void* const* inputs,
                    void* const* outputs, void* ws, cudaStream_t stream) noexcept override {
        if (upload() != 0) return DSVT_ERR_CUDA;
        if (batch_of(in) != 1) { std::fprintf(stderr, "[dsvt_b200] VfeFusedPlugin: batch 1 only\n"); return DSVT_ERR_UNSUPPORTED; }
        return report(dsvt_vfe_fused_launch(pfn0_, pfn1_, static_cast<const float*>(inputs[0]), static_cast<const int32_t*>(inputs[1]),
                                            static_cast<const int32_t*>(inputs[2]), static_cast<const int32_t*>(inputs[3]), max_points_,
                                            max_pillars_, npv_, static_cast<float*>(outputs[0]), ws,
                                            dsvt_vfe_fused_workspace_size(max_points_, npv_), 1, stream), kName);
    }
protected:
    const std::vector<DataType>& io_types() const override {
        static const std::vector<DataType> t{F, I, I, I, F};
        return t;
    }
    size_t nb_inputs() const override { return 4; }
private:
    int upload() {
        if (!pfn0_) pfn0_ = dsvt_small_linear_create(kMid, kIn, w0_.data(), sc0_.data(), sh0_.data());
        if (!pfn1_) pfn1_ = dsvt_linear_weights_create(kOut, kOut, w1_.data(), b1_.data(), DSVT_ATTN_FP32_TC);
        return pfn0_ && pfn1_ ? 0 : 1;
    }
    void release() {
        if (pfn0_) { dsvt_small_linear_destroy(pfn0_); pfn0_ = nullptr; }
        if (pfn1_) { dsvt_linear_weights_destroy(pfn1_); pfn1_ = nullptr; }
    }
    int max_points_, max_pillars_, npv_;
    std::vector<float> w0_, sc0_, sh0_, w1_, b1_;
    dsvt_small_linear* pfn0_ = nullptr;
    dsvt_linear_weights* pfn1_ = nullptr;
};

// registration: from the library only (the reference registers from two images)
using Points2FeaturesPluginCreator = CreatorBase<Points2FeaturesPlugin>;
using WindowPartitionPluginCreator = CreatorBase<WindowPartitionPlugin>;
using GetSetPluginCreator = CreatorBase<GetSetPlugin>;
using GeluPluginCreator = CreatorBase<GeluPlugin>;
using LayerNormPluginCreator = CreatorBase<LayerNormPlugin>;
using FilterBoxByScorePluginCreator = CreatorBase<FilterBoxByScorePlugin>;
using GetValueByIndexPluginCreator = CreatorBase<GetValueByIndexPlugin>;
using MapSetFeature2VoxelPluginCreator = CreatorBase<MapSetFeature2VoxelPlugin>;
using SetAttentionPluginCreator = CreatorBase<SetAttentionPlugin>;
using SetAttentionFusedPluginCreator = CreatorBase<SetAttentionFusedPlugin>;
using TorchScatterMaxPluginCreator = CreatorBase<TorchScatterMaxPlugin>;
using Map2BevPluginCreator = CreatorBase<Map2BevPlugin>;
using SetAttentionPlanPluginCreator = CreatorBase<SetAttentionPlanPlugin>;
using LayerNormChainPluginCreator = CreatorBase<LayerNormChainPlugin>;
using LinearPluginCreator = CreatorBase<LinearPlugin>;
using FfnFusedPluginCreator = CreatorBase<FfnFusedPlugin>;
using VfeFusedPluginCreator = CreatorBase<VfeFusedPlugin>;

REGISTER_TENSORRT_PLUGIN(Points2FeaturesPluginCreator);
REGISTER_TENSORRT_PLUGIN(WindowPartitionPluginCreator);
REGISTER_TENSORRT_PLUGIN(GetSetPluginCreator);
REGISTER_TENSORRT_PLUGIN(GeluPluginCreator);
REGISTER_TENSORRT_PLUGIN(LayerNormPluginCreator);
REGISTER_TENSORRT_PLUGIN(FilterBoxByScorePluginCreator);
REGISTER_TENSORRT_PLUGIN(GetValueByIndexPluginCreator);
REGISTER_TENSORRT_PLUGIN(MapSetFeature2VoxelPluginCreator);
REGISTER_TENSORRT_PLUGIN(SetAttentionPluginCreator);
REGISTER_TENSORRT_PLUGIN(SetAttentionFusedPluginCreator);
REGISTER_TENSORRT_PLUGIN(TorchScatterMaxPluginCreator);
REGISTER_TENSORRT_PLUGIN(Map2BevPluginCreator);
REGISTER_TENSORRT_PLUGIN(SetAttentionPlanPluginCreator);
REGISTER_TENSORRT_PLUGIN(LayerNormChainPluginCreator);
REGISTER_TENSORRT_PLUGIN(LinearPluginCreator);
REGISTER_TENSORRT_PLUGIN(FfnFusedPluginCreator);
REGISTER_TENSORRT_PLUGIN(VfeFusedPluginCreator);

}  // namespace dsvt_plugins
