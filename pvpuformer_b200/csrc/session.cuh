#pragma once
#include "common.cuh"

namespace vpu {

// Device-resident state of S click sessions of the NoBRS predictor (session.cu).  Mirrors vpu_session_state of the C ABI.
struct SessionState {
    int32_t S, H, W, T, max_clicks, n_half;
    const float* images;      // [S,3,H,W] fp32 in [0,1]
    float* prev_probs;        // [S,H,W] last full-size probabilities (predictor.prev_prediction == ZoomIn._prev_probs)
    uint8_t* pred;            // [S,H,W] prev_probs > pred_thr (input of the device clicker)
    int32_t* clicks;          // [S,max_clicks,3] (is_positive, row, col); the click order is the index
    int32_t* nclicks;         // [S]
    int32_t* roi;             // [S,4] rmin,rmax,cmin,cmax (inclusive); roi[0] < 0: none yet
    int32_t* fgbox;           // [S,5] bbox of prev_probs > zoom_thr + state (-1 no prediction yet, 0 empty, 1 non-empty)
    float pred_thr, zoom_thr;
    double expansion_ratio, recompute_thresh_iou;
    int32_t min_crop_size;
};

int session_prepare_launch(const SessionState& st, const int32_t* active, int A, const int32_t* new_clicks, float* net_image,
                           double* net_points, cudaStream_t stream);
int image_from_u8_launch(const uint8_t* rgb_nhwc, const float* prev, float* image4, int B, int H, int W, cudaStream_t stream);
int session_finish_launch(const SessionState& st, const int32_t* active, int A, const float* logits, cudaStream_t stream);

}  // namespace vpu
