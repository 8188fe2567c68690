// Common scaffolding of the TensorRT plugin shells.
//
// The reference repeats ~350 lines of IPluginV2DynamicExt / IPluginCreator boilerplate in each of
// its ten plugins.  Here a plugin is described once -- name, creator field
// list, serialised POD layout, I/O signature -- and the shell only forwards enqueue() to the C ABI
// of include/dsvt_b200.h.  Names, versions ("1"), field names, output shapes/dtypes and the byte
// layout of serialize() are the reference's, so plans and graph-building code keep working.
#pragma once
#include "NvInferPlugin.h"
#include "dsvt_b200.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace dsvt_plugins {

using namespace nvinfer1;

// ---- little-endian POD (de)serialisation, same as the reference's writeToBuffer/readFromBuffer --
class Writer {
public:
    explicit Writer(void* p) : p_(static_cast<char*>(p)) {}
    template <typename T> void put(const T& v) { std::memcpy(p_, &v, sizeof(T)); p_ += sizeof(T); }
    void put_array(const float* v, size_t n) { std::memcpy(p_, v, n * sizeof(float)); p_ += n * sizeof(float); }
private:
    char* p_;
};

class Reader {
public:
    Reader(const void* p, size_t n) : p_(static_cast<const char*>(p)), left_(n) {}
    template <typename T> T get() {
        T v{};
        if (left_ >= sizeof(T)) { std::memcpy(&v, p_, sizeof(T)); p_ += sizeof(T); left_ -= sizeof(T); }
        else ok_ = false;
        return v;
    }
    void get_array(float* dst, size_t n) {
        if (left_ >= n * sizeof(float)) { std::memcpy(dst, p_, n * sizeof(float)); p_ += n * sizeof(float); left_ -= n * sizeof(float); }
        else ok_ = false;
    }
    bool ok() const { return ok_; }
    size_t left() const { return left_; }
private:
    const char* p_;
    size_t left_;
    bool ok_ = true;
};

// ---- field lookup ------------------------------------------------------------------------------
inline const PluginField* find_field(const PluginFieldCollection* fc, const char* name) {
    if (!fc) return nullptr;
    for (int i = 0; i < fc->nbFields; ++i)
        if (fc->fields[i].name && !std::strcmp(fc->fields[i].name, name)) return &fc->fields[i];
    return nullptr;
}
inline int field_int(const PluginFieldCollection* fc, const char* name, int idx = 0, int dflt = 0) {
    const PluginField* f = find_field(fc, name);
    return (f && f->data) ? static_cast<const int*>(f->data)[idx] : dflt;
}
inline float field_float(const PluginFieldCollection* fc, const char* name, int idx = 0, float dflt = 0.f) {
    const PluginField* f = find_field(fc, name);
    return (f && f->data) ? static_cast<const float*>(f->data)[idx] : dflt;
}

// ---- plugin base: everything that is identical across plugins ----------------------------------
class PluginBase : public IPluginV2DynamicExt {
public:
    const char* getPluginVersion() const noexcept override { return "1"; }
    int32_t initialize() noexcept override { return 0; }
    void terminate() noexcept override {}
    void destroy() noexcept override { delete this; }
    void setPluginNamespace(const char* ns) noexcept override { ns_ = ns ? ns : ""; }
    const char* getPluginNamespace() const noexcept override { return ns_.c_str(); }
    void configurePlugin(const DynamicPluginTensorDesc*, int32_t, const DynamicPluginTensorDesc*,
                         int32_t) noexcept override {}
    size_t getWorkspaceSize(const PluginTensorDesc*, int32_t, const PluginTensorDesc*,
                            int32_t) const noexcept override { return 0; }

    // I/O signature: dtype of every input then every output (all kLINEAR)
    bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* inOut, int32_t nbInputs,
                                   int32_t nbOutputs) noexcept override {
        const std::vector<DataType>& sig = io_types();
        if (pos < 0 || pos >= nbInputs + nbOutputs || pos >= (int32_t) sig.size()) return false;
        return inOut[pos].type == sig[pos] && inOut[pos].format == TensorFormat::kLINEAR;
    }
    DataType getOutputDataType(int32_t index, const DataType*, int32_t) const noexcept override {
        const std::vector<DataType>& sig = io_types();
        const size_t i = nb_inputs() + (size_t) index;
        return i < sig.size() ? sig[i] : DataType::kFLOAT;
    }
protected:
    virtual const std::vector<DataType>& io_types() const = 0;
    virtual size_t nb_inputs() const = 0;
    static DimsExprs dims(IExprBuilder& b, const IDimensionExpr* batch, std::initializer_list<int> rest) {
        DimsExprs d{};
        d.nbDims = 1 + (int32_t) rest.size();
        d.d[0] = batch;
        int i = 1;
        for (int v : rest) d.d[i++] = b.constant(v);
        return d;
    }
    // the reference ABORTS on CUDA errors; we log and return the non-zero code (SURVEY.md 8b "Errors")
    static int report(int rc, const char* who) {
        if (rc != DSVT_OK) std::fprintf(stderr, "[dsvt_b200] %s: enqueue failed (%d): %s\n", who, rc, dsvt_last_error());
        return rc;
    }
    std::string ns_;
};

template <typename PluginT>
class CreatorBase : public IPluginCreator {
public:
    const char* getPluginName() const noexcept override { return PluginT::kName; }
    const char* getPluginVersion() const noexcept override { return "1"; }
    const PluginFieldCollection* getFieldNames() noexcept override {
        if (attrs_.empty()) {
            for (const auto& f : PluginT::field_list()) attrs_.emplace_back(f.first, nullptr, f.second, 1);
            fc_.nbFields = (int32_t) attrs_.size();
            fc_.fields = attrs_.data();
        }
        return &fc_;
    }
    IPluginV2* createPlugin(const char*, const PluginFieldCollection* fc) noexcept override {
        return PluginT::from_fields(fc);
    }
    IPluginV2* deserializePlugin(const char*, const void* data, size_t len) noexcept override {
        return PluginT::from_bytes(data, len);
    }
    void setPluginNamespace(const char* ns) noexcept override { ns_ = ns ? ns : ""; }
    const char* getPluginNamespace() const noexcept override { return ns_.c_str(); }
private:
    std::vector<PluginField> attrs_;
    PluginFieldCollection fc_{};
    std::string ns_;
};

using FieldList = std::vector<std::pair<const char*, PluginFieldType>>;

}  // namespace dsvt_plugins
