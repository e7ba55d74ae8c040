#pragma once
#include "common.cuh"

namespace vpu {

// Next oracle click + IoU counts for S click sessions on [S, H, W] masks (noc.cu).
//   gt int8 (1 object, 0 background, -1 ignore), pred uint8 (0/1), not_clicked uint8 (updated in place),
//   clicks int32 [S][4] = (is_positive, row, col, squared distance), iou_counts int64 [S][2] = (intersection, union)
size_t noc_workspace_bytes(int S, int H, int W);
int noc_next_clicks_launch(const int8_t* gt, const uint8_t* pred, uint8_t* not_clicked, int S, int H, int W, int32_t* clicks,
                           long long* iou_counts, void* workspace, cudaStream_t stream);

}  // namespace vpu
