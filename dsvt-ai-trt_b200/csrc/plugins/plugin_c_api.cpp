// C harness over nvinfer1::IPluginCreator / IPluginV2DynamicExt (see include/dsvt_b200_plugin_c.h).
// Compiled twice: into libdsvt_b200_plugins.so next to our plugin shells, and (with
// -DDSVT_HARNESS_FOR_REFERENCE) next to each unmodified reference plugin source for oracle/_ref.
#include "NvInferPlugin.h"
#include "dsvt_b200_plugin_c.h"

#include <cstring>
#include <memory>
#include <new>
#include <vector>

#define DSVT_EXPORT extern "C" __attribute__((visibility("default")))

using namespace nvinfer1;

struct dsvt_plugin {
    IPluginV2DynamicExt* impl;
};

namespace {

// constant-only IExprBuilder: enough for plugins whose output shapes are static capacities
struct ConstExpr final : IDimensionExpr {
    int32_t v;
    explicit ConstExpr(int32_t x) : v(x) {}
    bool isConstant() const noexcept override { return true; }
    int32_t getConstantValue() const noexcept override { return v; }
};

struct ConstBuilder final : IExprBuilder {
    std::vector<std::unique_ptr<ConstExpr>> pool;
    const IDimensionExpr* constant(int32_t value) noexcept override {
        pool.emplace_back(new ConstExpr(value));
        return pool.back().get();
    }
    const IDimensionExpr* operation(DimensionOperation op, const IDimensionExpr& a,
                                    const IDimensionExpr& b) noexcept override {
        const int32_t x = a.getConstantValue(), y = b.getConstantValue();
        int32_t r = 0;
        switch (op) {
            case DimensionOperation::kSUM: r = x + y; break;
            case DimensionOperation::kPROD: r = x * y; break;
            case DimensionOperation::kMAX: r = x > y ? x : y; break;
            case DimensionOperation::kMIN: r = x < y ? x : y; break;
            case DimensionOperation::kSUB: r = x - y; break;
            case DimensionOperation::kEQUAL: r = x == y; break;
            case DimensionOperation::kLESS: r = x < y; break;
            case DimensionOperation::kFLOOR_DIV: r = y ? x / y : 0; break;
            case DimensionOperation::kCEIL_DIV: r = y ? (x + y - 1) / y : 0; break;
        }
        return constant(r);
    }
};

PluginTensorDesc to_trt(const dsvt_tensor_desc& d) {
    PluginTensorDesc t{};
    t.dims.nbDims = d.nb_dims;
    for (int i = 0; i < 8; ++i) t.dims.d[i] = i < d.nb_dims ? d.dims[i] : 0;
    t.type = static_cast<DataType>(d.dtype);
    t.format = TensorFormat::kLINEAR;
    t.scale = 1.0f;
    return t;
}

std::vector<PluginTensorDesc> to_trt(const dsvt_tensor_desc* d, int n) {
    std::vector<PluginTensorDesc> v;
    for (int i = 0; i < n; ++i) v.push_back(to_trt(d[i]));
    return v;
}

dsvt_plugin* wrap(IPluginV2* raw) {
    if (!raw) return nullptr;
    auto* dyn = static_cast<IPluginV2DynamicExt*>(raw);   // every plugin here derives from IPluginV2DynamicExt
    if (dyn->initialize() != 0) { dyn->destroy(); return nullptr; }
    auto* h = new (std::nothrow) dsvt_plugin{dyn};
    if (!h) dyn->destroy();
    return h;
}

IPluginCreator* find_creator(const char* name, const char* version) {
    if (!name) return nullptr;
    return getPluginRegistry()->getPluginCreator(name, version ? version : "1", "");
}

}  // namespace

DSVT_EXPORT int dsvt_plugin_registry_size(void) {
    int32_t n = 0;
    getPluginRegistry()->getPluginCreatorList(&n);
    return n;
}

DSVT_EXPORT const char* dsvt_plugin_registry_name(int i) {
    int32_t n = 0;
    IPluginCreator* const* list = getPluginRegistry()->getPluginCreatorList(&n);
    return (i >= 0 && i < n) ? list[i]->getPluginName() : nullptr;
}

DSVT_EXPORT int dsvt_plugin_field_names(const char* plugin_name, const char* version, const char** names, int max) {
    IPluginCreator* c = find_creator(plugin_name, version);
    if (!c) return -1;
    const PluginFieldCollection* fc = c->getFieldNames();
    for (int i = 0; i < fc->nbFields && i < max; ++i) names[i] = fc->fields[i].name;
    return fc->nbFields;
}

DSVT_EXPORT dsvt_plugin* dsvt_plugin_create(const char* plugin_name, const char* version,
                                            const dsvt_plugin_field* fields, int n_fields) {
    IPluginCreator* c = find_creator(plugin_name, version);
    if (!c) return nullptr;
    std::vector<PluginField> f;
    for (int i = 0; i < n_fields; ++i)
        f.emplace_back(fields[i].name, fields[i].data, static_cast<PluginFieldType>(fields[i].type), fields[i].length);
    PluginFieldCollection fc{static_cast<int32_t>(f.size()), f.data()};
    return wrap(c->createPlugin(plugin_name, &fc));
}

DSVT_EXPORT dsvt_plugin* dsvt_plugin_deserialize(const char* plugin_name, const char* version, const void* data,
                                                 size_t len) {
    IPluginCreator* c = find_creator(plugin_name, version);
    if (!c || !data) return nullptr;
    IPluginV2* raw = c->deserializePlugin(plugin_name, data, len);
    if (!raw) return nullptr;
    // TensorRT clones a deserialised plugin before use; the reference's LayerNorm relies on that to
    // upload its weights (layerNorm.cu:160-190 vs :126-158), so do the same here.
    auto* dyn = static_cast<IPluginV2DynamicExt*>(raw);
    IPluginV2DynamicExt* cl = dyn->clone();
    dsvt_plugin* h = wrap(cl);
#ifndef DSVT_HARNESS_FOR_REFERENCE
    dyn->destroy();
#else
    (void) dyn;   // the reference's LayerNorm dtor cudaFree()s uninitialised pointers on this path; leak it instead
#endif
    return h;
}

DSVT_EXPORT dsvt_plugin* dsvt_plugin_clone(const dsvt_plugin* p) {
    return p ? wrap(p->impl->clone()) : nullptr;
}

DSVT_EXPORT void dsvt_plugin_destroy(dsvt_plugin* p) {
    if (!p) return;
    p->impl->terminate();
    p->impl->destroy();
#ifdef DSVT_HARNESS_FOR_REFERENCE
    // the reference's LayerNorm frees its device weights in terminate() AND again in its destructor
    // (layerNorm.cu:432-444); swallow the resulting cudaErrorInvalidValue so it does not leak into the caller
    (void) cudaGetLastError();
#endif
    delete p;
}

DSVT_EXPORT const char* dsvt_plugin_type(const dsvt_plugin* p) { return p ? p->impl->getPluginType() : nullptr; }
DSVT_EXPORT const char* dsvt_plugin_version(const dsvt_plugin* p) { return p ? p->impl->getPluginVersion() : nullptr; }
DSVT_EXPORT int dsvt_plugin_nb_outputs(const dsvt_plugin* p) { return p ? p->impl->getNbOutputs() : -1; }
DSVT_EXPORT size_t dsvt_plugin_serialization_size(const dsvt_plugin* p) {
    return p ? p->impl->getSerializationSize() : 0;
}
DSVT_EXPORT void dsvt_plugin_serialize(const dsvt_plugin* p, void* buffer) {
    if (p && buffer) p->impl->serialize(buffer);
}

DSVT_EXPORT int dsvt_plugin_output_desc(dsvt_plugin* p, int output_index, const dsvt_tensor_desc* inputs,
                                        int nb_inputs, dsvt_tensor_desc* out) {
    if (!p || !out || output_index < 0 || output_index >= p->impl->getNbOutputs()) return 1;
    ConstBuilder b;
    std::vector<DimsExprs> in(nb_inputs);
    std::vector<DataType> types(nb_inputs);
    for (int i = 0; i < nb_inputs; ++i) {
        in[i].nbDims = inputs[i].nb_dims;
        for (int k = 0; k < inputs[i].nb_dims; ++k) in[i].d[k] = b.constant(inputs[i].dims[k]);
        types[i] = static_cast<DataType>(inputs[i].dtype);
    }
    const DimsExprs o = p->impl->getOutputDimensions(output_index, in.data(), nb_inputs, b);
    out->nb_dims = o.nbDims;
    for (int k = 0; k < 8; ++k) out->dims[k] = k < o.nbDims ? o.d[k]->getConstantValue() : 0;
    out->dtype = static_cast<int32_t>(p->impl->getOutputDataType(output_index, types.data(), nb_inputs));
    return 0;
}

DSVT_EXPORT int dsvt_plugin_supports_format(dsvt_plugin* p, int pos, const dsvt_tensor_desc* in_out, int nb_inputs,
                                            int nb_outputs) {
    if (!p) return 0;
    auto v = to_trt(in_out, nb_inputs + nb_outputs);
    return p->impl->supportsFormatCombination(pos, v.data(), nb_inputs, nb_outputs) ? 1 : 0;
}

DSVT_EXPORT size_t dsvt_plugin_workspace_size(dsvt_plugin* p, const dsvt_tensor_desc* inputs, int nb_inputs,
                                              const dsvt_tensor_desc* outputs, int nb_outputs) {
    if (!p) return 0;
    auto in = to_trt(inputs, nb_inputs), out = to_trt(outputs, nb_outputs);
    return p->impl->getWorkspaceSize(in.data(), nb_inputs, out.data(), nb_outputs);
}

DSVT_EXPORT int dsvt_plugin_enqueue(dsvt_plugin* p, const dsvt_tensor_desc* inputs, int nb_inputs,
                                    const dsvt_tensor_desc* outputs, int nb_outputs,
                                    const void* const* input_ptrs, void* const* output_ptrs, void* workspace,
                                    dsvt_stream_t stream) {
    if (!p) return -1;
    auto in = to_trt(inputs, nb_inputs), out = to_trt(outputs, nb_outputs);
    // TensorRT calls configurePlugin() with the concrete shapes before the first enqueue
    std::vector<DynamicPluginTensorDesc> din(nb_inputs), dout(nb_outputs);
    for (int i = 0; i < nb_inputs; ++i) { din[i].desc = in[i]; din[i].min = in[i].dims; din[i].max = in[i].dims; }
    for (int i = 0; i < nb_outputs; ++i) { dout[i].desc = out[i]; dout[i].min = out[i].dims; dout[i].max = out[i].dims; }
    p->impl->configurePlugin(din.data(), nb_inputs, dout.data(), nb_outputs);
    return p->impl->enqueue(in.data(), out.data(), input_ptrs, output_ptrs, workspace,
                            reinterpret_cast<cudaStream_t>(stream));
}
