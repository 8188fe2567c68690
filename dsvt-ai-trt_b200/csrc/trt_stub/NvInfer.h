// Minimal, SOURCE-compatible scaffold of the TensorRT 8.x plugin interface.
//
// TensorRT is not installed in the build image, so the plugin
// shells in ../plugins/ and the reference's own plugin sources (oracle/_ref) are
// compiled against this header instead.  It declares only what
// IPluginV2DynamicExt / IPluginCreator plugins touch, plus a tiny in-process
// plugin registry so that REGISTER_TENSORRT_PLUGIN + getPluginRegistry()
// ->getPluginCreator(name, version) work the way include/plugin_helper.h of the
// reference uses them (plugin_helper.h:44, :262, :498, :566, :618).
//
// ABI WARNING: the vtable layout here is NOT the one of a real libnvinfer.
// Objects built against this header must never be handed to a real TensorRT.
// For deployment, compile ../plugins/*.cpp against the genuine NvInfer.h of the
// target TensorRT (the shells only use the methods declared below, which exist
// with identical signatures in TensorRT 8.2 - 8.6).  All kernel logic sits
// behind the extern "C" API of include/dsvt_b200.h and does not depend on this.
#ifndef DSVT_B200_TRT_STUB_NVINFER_H
#define DSVT_B200_TRT_STUB_NVINFER_H

#define DSVT_B200_TRT_STUB 1
#define NV_TENSORRT_MAJOR 8
#define NV_TENSORRT_MINOR 2

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <cuda_runtime_api.h>

namespace nvinfer1 {

using AsciiChar = char;

enum class DataType : int32_t { kFLOAT = 0, kHALF = 1, kINT8 = 2, kINT32 = 3, kBOOL = 4 };
enum class TensorFormat : int32_t { kLINEAR = 0, kCHW2 = 1, kHWC8 = 2, kCHW4 = 3, kCHW16 = 4, kCHW32 = 5 };
using PluginFormat = TensorFormat;

// gelu.cu / layerNorm.cu of the reference pull in include/logging.h, which
// only needs the logger base class to exist.
class ILogger {
public:
    enum class Severity : int32_t { kINTERNAL_ERROR = 0, kERROR = 1, kWARNING = 2, kINFO = 3, kVERBOSE = 4 };
    virtual void log(Severity severity, const AsciiChar* msg) noexcept = 0;
    ILogger() = default;
    virtual ~ILogger() = default;
};

class Dims {
public:
    static constexpr int32_t MAX_DIMS = 8;
    int32_t nbDims;
    int32_t d[MAX_DIMS];
};

class Weights {
public:
    DataType type;
    const void* values;
    int64_t count;
};

// ---- dimension expressions (build-time shape arithmetic) -----------------
class IDimensionExpr {
public:
    virtual bool isConstant() const noexcept = 0;
    virtual int32_t getConstantValue() const noexcept = 0;
protected:
    virtual ~IDimensionExpr() noexcept = default;
};

enum class DimensionOperation : int32_t {
    kSUM = 0, kPROD = 1, kMAX = 2, kMIN = 3, kSUB = 4, kEQUAL = 5, kLESS = 6, kFLOOR_DIV = 7, kCEIL_DIV = 8
};

class IExprBuilder {
public:
    virtual const IDimensionExpr* constant(int32_t value) noexcept = 0;
    virtual const IDimensionExpr* operation(DimensionOperation op, const IDimensionExpr& first,
                                            const IDimensionExpr& second) noexcept = 0;
protected:
    virtual ~IExprBuilder() noexcept = default;
};

class DimsExprs {
public:
    int32_t nbDims;
    const IDimensionExpr* d[Dims::MAX_DIMS];
};

struct PluginTensorDesc {
    Dims dims;
    DataType type;
    TensorFormat format;
    float scale;
};

struct DynamicPluginTensorDesc {
    PluginTensorDesc desc;
    Dims min;
    Dims max;
};

// ---- plugin fields ---------------------------------------------------------
enum class PluginFieldType : int32_t {
    kFLOAT16 = 0, kFLOAT32 = 1, kFLOAT64 = 2, kINT8 = 3, kINT16 = 4, kINT32 = 5, kCHAR = 6, kDIMS = 7, kUNKNOWN = 8
};

class PluginField {
public:
    const AsciiChar* name;
    const void* data;
    PluginFieldType type;
    int32_t length;
    PluginField(const AsciiChar* name_ = nullptr, const void* data_ = nullptr,
                PluginFieldType type_ = PluginFieldType::kUNKNOWN, int32_t length_ = 0) noexcept
        : name(name_), data(data_), type(type_), length(length_) {}
};

struct PluginFieldCollection {
    int32_t nbFields;
    const PluginField* fields;
};

// ---- plugin interfaces -----------------------------------------------------
class IPluginV2 {
public:
    virtual const AsciiChar* getPluginType() const noexcept = 0;
    virtual const AsciiChar* getPluginVersion() const noexcept = 0;
    virtual int32_t getNbOutputs() const noexcept = 0;
    virtual int32_t initialize() noexcept = 0;
    virtual void terminate() noexcept = 0;
    virtual size_t getSerializationSize() const noexcept = 0;
    virtual void serialize(void* buffer) const noexcept = 0;
    virtual void destroy() noexcept = 0;
    virtual void setPluginNamespace(const AsciiChar* pluginNamespace) noexcept = 0;
    virtual const AsciiChar* getPluginNamespace() const noexcept = 0;
    virtual ~IPluginV2() noexcept = default;
};

class IPluginV2Ext : public IPluginV2 {
public:
    virtual DataType getOutputDataType(int32_t index, const DataType* inputTypes,
                                       int32_t nbInputs) const noexcept = 0;
};

class IPluginV2DynamicExt : public IPluginV2Ext {
public:
    virtual IPluginV2DynamicExt* clone() const noexcept = 0;
    virtual DimsExprs getOutputDimensions(int32_t outputIndex, const DimsExprs* inputs, int32_t nbInputs,
                                          IExprBuilder& exprBuilder) noexcept = 0;
    virtual bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* inOut, int32_t nbInputs,
                                           int32_t nbOutputs) noexcept = 0;
    virtual void configurePlugin(const DynamicPluginTensorDesc* in, int32_t nbInputs,
                                 const DynamicPluginTensorDesc* out, int32_t nbOutputs) noexcept = 0;
    virtual size_t getWorkspaceSize(const PluginTensorDesc* inputs, int32_t nbInputs,
                                    const PluginTensorDesc* outputs, int32_t nbOutputs) const noexcept = 0;
    virtual int32_t enqueue(const PluginTensorDesc* inputDesc, const PluginTensorDesc* outputDesc,
                            const void* const* inputs, void* const* outputs, void* workspace,
                            cudaStream_t stream) noexcept = 0;
};

class IPluginCreator {
public:
    virtual const AsciiChar* getPluginName() const noexcept = 0;
    virtual const AsciiChar* getPluginVersion() const noexcept = 0;
    virtual const PluginFieldCollection* getFieldNames() noexcept = 0;
    virtual IPluginV2* createPlugin(const AsciiChar* name, const PluginFieldCollection* fc) noexcept = 0;
    virtual IPluginV2* deserializePlugin(const AsciiChar* name, const void* serialData,
                                         size_t serialLength) noexcept = 0;
    virtual void setPluginNamespace(const AsciiChar* pluginNamespace) noexcept = 0;
    virtual const AsciiChar* getPluginNamespace() const noexcept = 0;
    virtual ~IPluginCreator() = default;
};

// ---- registry (process-local; one instance per shared object) -------------
class IPluginRegistry {
public:
    virtual bool registerCreator(IPluginCreator& creator, const AsciiChar* pluginNamespace) noexcept = 0;
    virtual IPluginCreator* const* getPluginCreatorList(int32_t* numCreators) const noexcept = 0;
    virtual IPluginCreator* getPluginCreator(const AsciiChar* pluginName, const AsciiChar* pluginVersion,
                                             const AsciiChar* pluginNamespace = "") noexcept = 0;
protected:
    virtual ~IPluginRegistry() noexcept = default;
};

namespace stub_detail {
class Registry final : public IPluginRegistry {
public:
    static constexpr int32_t kMax = 64;
    bool registerCreator(IPluginCreator& creator, const AsciiChar* ns) noexcept override {
        for (int32_t i = 0; i < n_; ++i) {
            if (!std::strcmp(list_[i]->getPluginName(), creator.getPluginName()) &&
                !std::strcmp(list_[i]->getPluginVersion(), creator.getPluginVersion())) {
                return false;  // already registered (the reference registers from two images, SURVEY A-11)
            }
        }
        if (n_ >= kMax) return false;
        creator.setPluginNamespace(ns ? ns : "");
        list_[n_++] = &creator;
        return true;
    }
    IPluginCreator* const* getPluginCreatorList(int32_t* numCreators) const noexcept override {
        if (numCreators) *numCreators = n_;
        return list_;
    }
    IPluginCreator* getPluginCreator(const AsciiChar* name, const AsciiChar* version,
                                     const AsciiChar* ns = "") noexcept override {
        (void) ns;
        for (int32_t i = 0; i < n_; ++i) {
            if (!std::strcmp(list_[i]->getPluginName(), name) &&
                !std::strcmp(list_[i]->getPluginVersion(), version)) {
                return list_[i];
            }
        }
        return nullptr;
    }
private:
    IPluginCreator* list_[kMax] = {};
    int32_t n_ = 0;
};
// hidden visibility: one registry per shared object, never unified across dlopen()ed libraries
__attribute__((visibility("hidden"))) inline Registry& registry() {
    static Registry r;
    return r;
}
}  // namespace stub_detail

template <typename T>
class PluginRegistrar {
public:
    PluginRegistrar() { stub_detail::registry().registerCreator(instance, ""); }
private:
    T instance{};
};

}  // namespace nvinfer1

__attribute__((visibility("hidden"))) inline nvinfer1::IPluginRegistry* getPluginRegistry() noexcept { return &nvinfer1::stub_detail::registry(); }

#define REGISTER_TENSORRT_PLUGIN(name) \
    static nvinfer1::PluginRegistrar<name> pluginRegistrar##name {}

#endif  // DSVT_B200_TRT_STUB_NVINFER_H
