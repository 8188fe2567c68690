// C entry point around the reference's OWN host-side rotated NMS (include/helper.h:92-283), compiled unmodified from
// /root/reference by oracle/build.py into oracle/_ref/libref_nms.so.  TEST INFRASTRUCTURE ONLY.
// helper.h relies on what src/dsvt-ai-trt.cpp sets up before including it (:1-30: <cassert>, `using namespace std`)
#include <cassert>
#include <cmath>
#include <string>
#include <vector>
using namespace std;
#include "helper.h"

extern "C" int ref_nms_cpu(const float* boxes, int n, float nms_thresh, int* keep)
{
    std::vector<Bndbox> in, out;
    for (int i = 0; i < n; ++i)      // id carries the input index (the reference stores the class there; NMS never reads it)
        in.emplace_back(boxes[i * 9], boxes[i * 9 + 1], boxes[i * 9 + 2], boxes[i * 9 + 3], boxes[i * 9 + 4], boxes[i * 9 + 5],
                        boxes[i * 9 + 6], i, boxes[i * 9 + 8]);
    nms_cpu(in, nms_thresh, out);
    for (size_t k = 0; k < out.size(); ++k) keep[k] = out[k].id;
    return (int) out.size();
}
