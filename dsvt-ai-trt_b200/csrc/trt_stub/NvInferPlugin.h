// Scaffold counterpart of TensorRT's NvInferPlugin.h: the reference plugin
// headers include only this file (e.g. plugins/include/getSet.h:4).
#ifndef DSVT_B200_TRT_STUB_NVINFERPLUGIN_H
#define DSVT_B200_TRT_STUB_NVINFERPLUGIN_H
#include "NvInfer.h"
#endif
