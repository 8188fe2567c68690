/*
 * dsvt_b200_plugin_c.h -- C harness over the TensorRT plugin surface.
 *
 * The reference binds its plugins to TensorRT through nvinfer1::IPluginCreator /
 * IPluginV2DynamicExt (C++ vtables; the headers under plugins/include/ and include/plugin_helper.h).  This header
 * exposes exactly those calls -- look a creator up by name/version in the plugin registry, build a
 * PluginFieldCollection, createPlugin / deserializePlugin, then getOutputDimensions,
 * getWorkspaceSize, enqueue, serialize, clone, destroy -- as plain C functions, so that tests,
 * ctypes and non-C++ hosts can drive a plugin the way TensorRT would.  The same harness source
 * (dsvt-ai-trt_b200/csrc/plugins/plugin_c_api.cpp) is compiled against the reference's unmodified
 * plugin sources to form oracle/_ref/libref_<plugin>.so, so both sides are driven identically.
 */
#ifndef DSVT_B200_PLUGIN_C_H
#define DSVT_B200_PLUGIN_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* dsvt_stream_t;
typedef struct dsvt_plugin dsvt_plugin;     /* wraps one nvinfer1::IPluginV2DynamicExt* */

/* nvinfer1::PluginFieldType values */
enum { DSVT_FIELD_FLOAT32 = 1, DSVT_FIELD_INT32 = 5 };
/* nvinfer1::DataType values */
enum { DSVT_DTYPE_FLOAT = 0, DSVT_DTYPE_HALF = 1, DSVT_DTYPE_INT32 = 3 };

typedef struct dsvt_plugin_field {
    const char* name;
    const void* data;
    int32_t type;     /* DSVT_FIELD_* */
    int32_t length;   /* number of elements */
} dsvt_plugin_field;

/* tensor descriptor == nvinfer1::PluginTensorDesc with kLINEAR format */
typedef struct dsvt_tensor_desc {
    int32_t nb_dims;
    int32_t dims[8];
    int32_t dtype;    /* DSVT_DTYPE_* */
} dsvt_tensor_desc;

/* registry: number of registered creators / their names (getPluginRegistry()->getPluginCreatorList) */
int dsvt_plugin_registry_size(void);
const char* dsvt_plugin_registry_name(int i);
/* creator->getFieldNames(): returns the count, fills up to `max` names */
int dsvt_plugin_field_names(const char* plugin_name, const char* version, const char** names, int max);

/* creator->createPlugin(name, fc) / creator->deserializePlugin(name, data, len); NULL on failure */
dsvt_plugin* dsvt_plugin_create(const char* plugin_name, const char* version,
                                const dsvt_plugin_field* fields, int n_fields);
dsvt_plugin* dsvt_plugin_deserialize(const char* plugin_name, const char* version, const void* data, size_t len);
dsvt_plugin* dsvt_plugin_clone(const dsvt_plugin* p);
void dsvt_plugin_destroy(dsvt_plugin* p);

const char* dsvt_plugin_type(const dsvt_plugin* p);
const char* dsvt_plugin_version(const dsvt_plugin* p);
int dsvt_plugin_nb_outputs(const dsvt_plugin* p);
size_t dsvt_plugin_serialization_size(const dsvt_plugin* p);
void dsvt_plugin_serialize(const dsvt_plugin* p, void* buffer);

/* getOutputDimensions / getOutputDataType with constant input shapes */
int dsvt_plugin_output_desc(dsvt_plugin* p, int output_index, const dsvt_tensor_desc* inputs, int nb_inputs,
                            dsvt_tensor_desc* out);
/* supportsFormatCombination(pos, inOut, nbInputs, nbOutputs) with kLINEAR tensors */
int dsvt_plugin_supports_format(dsvt_plugin* p, int pos, const dsvt_tensor_desc* in_out, int nb_inputs,
                                int nb_outputs);
size_t dsvt_plugin_workspace_size(dsvt_plugin* p, const dsvt_tensor_desc* inputs, int nb_inputs,
                                  const dsvt_tensor_desc* outputs, int nb_outputs);
/* initialize() is called by create/deserialize/clone; returns enqueue()'s status (0 = success) */
int dsvt_plugin_enqueue(dsvt_plugin* p, const dsvt_tensor_desc* inputs, int nb_inputs,
                        const dsvt_tensor_desc* outputs, int nb_outputs,
                        const void* const* input_ptrs, void* const* output_ptrs, void* workspace,
                        dsvt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
