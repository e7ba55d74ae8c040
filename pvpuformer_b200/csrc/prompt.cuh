#pragma once
#include <climits>

#include "common.cuh"

namespace vpu {

struct PpueArgs {
    const double* points = nullptr;     // [B, 2n, 3] (coord0, coord1, order); order == -1 => padding
    int n = 0;                          // clicks per half actually present (<= num_max_points)
    int num_max_points = 24;
    int size = 448;                     // image side; row length = 2*size + 3
    int type = 0;                       // 0 clicks, 1 box, 2 scribble
    const int* boxes = nullptr;         // [B, 5] (x_c, y_c, w, h, slot)            (type 1)
    const int* scrib_sel = nullptr;     // [B, 2, size] offsets or INT_MIN           (type 2)
    const int* scrib_slot = nullptr;    // [B] slot of the scribble row or -1         (type 2)
    float* out = nullptr;               // [B, 2*num_max_points, 2*size+3] fp32
    __nv_bfloat16* out_bf16 = nullptr;  // optional [B*2*num_max_points, ld_bf16], zero padded
    int ld_bf16 = 0;
    int click_radius = 9;
    float click_table[32];              // 2*click_radius+1 taps (host: the reference's float32 formula)
};
int ppue_launch(const PpueArgs& a, int B, cudaStream_t stream);

struct CoordArgs {
    const float* image4 = nullptr;        // [B, 4, H, W] fp32 (RGB + previous mask)
    const double* points = nullptr;       // [B, 2n, 3] (row, col, order)
    const uint8_t* extra_mask = nullptr;  // optional [B, 2, H, W]: box/scribble raster OR-ed into the disks
    int n = 0, H = 0, W = 0;
    float radius = 5.f;
};
int coord_features_launch(const CoordArgs& a, int B, float* out, cudaStream_t stream);
int patch_operand_launch(const CoordArgs& a, int B, __nv_bfloat16* A, int patch, int lda, cudaStream_t stream);

}  // namespace vpu
