// b2t_decode.cuh -- the arithmetic of Detect.forward's inference branch (models/yolo.py:44-55), shared by the
// kernel that materialises `pred` (b2t_detect.cu) and by the fused decode + candidate filter (b2t_nms.cu) so that
// both produce the same floats:
//     y = x[i].sigmoid()
//     xy = (y[..., 0:2] * 2. - 0.5 + grid) * stride          wh = (y[..., 2:4] * 2) ** 2 * anchor_grid
#pragma once
#include "b2t_platform.cuh"

namespace b2t {

B2T_DEV float det_sigmoid(float r) { return 1.0f / (1.0f + expf(-r)); }
B2T_DEV float det_xy(float s, float g, float stride) { return (s * 2.0f - 0.5f + g) * stride; }
B2T_DEV float det_wh(float s, float anchor) { return (s * 2.0f) * (s * 2.0f) * anchor; }

}  // namespace b2t
