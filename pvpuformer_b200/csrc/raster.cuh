#pragma once
#include "common.cuh"

namespace vpu {

// Box / scribble outline planes (raster.cu): what cv2.rectangle / cv2.polylines (thickness 3) draw in the reference
struct RasterArgs {
    int type = 0;                       // 1 box, 2 scribble
    const int32_t* boxes = nullptr;     // [B,5] (x_c, y_c, w, h, slot)
    const int32_t* scribbles = nullptr; // [B,S,2] (x, y)
    int S = 0;                          // points per scribble
    int n = 0;                          // click slots per half: slot < n -> plane 0 (positive), else plane 1
    int size = 0;                       // image side
    uint8_t* planes = nullptr;          // [B,2,size,size], cleared by the launch
};
int raster_prompts_launch(const RasterArgs& a, int B, cudaStream_t stream);

}  // namespace vpu
