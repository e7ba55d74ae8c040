// Device-resident NoBRS click sessions: the predictor transforms either side of the forward.
// Reference: isegm/inference/predictors/base.py:106-151,195-213 (input assembly, get_points_nd), transforms/zoom_in.py:30-112,
// 171-189 (ZoomIn.transform / inv_transform / _transform_clicks), transforms/flip.py:9-28 (AddHorizontalFlip),
// transforms/base.py:29-38 (SigmoidForPred), utils/misc.py:36-79 (bbox helpers), in the configuration
// scripts/evaluate_vpumodel.py:187-192 builds for the VPU models (skip_clicks = -1, fixed target size, flip TTA).
//
// One click of S sessions is two launches before the forward and one after it, with no host round trip:
//   session_roi_points_kernel  one block per active session: append the clicker's new click, recompute the zoom-in region
//                              (bbox of prev_probs > 0.5 joined with the positive clicks, expanded x1.4, >= min_crop_size,
//                              clamped; kept unless a positive click left the old region or bbox-IoU < 0.5), write the
//                              network's point rows for the crop and for its mirror image.  All of it is the reference's
//                              float64 / integer arithmetic, evaluated with explicitly rounded operations (no contraction),
//                              Python's round() = rint (half to even).  The point rows stay float64: the reference's clicker
//                              takes its coordinates from np.where (numpy int64), so the rescaled coordinates are numpy
//                              float64 scalars and torch.tensor() keeps that dtype (clicker.py:55-69, base.py:195-213).
//   session_crop_kernel        network input [2A,4,T,T]: RGB + previous probabilities of the region, bilinear
//                              align_corners=True to T x T, plus the horizontally flipped copy.
//   session_finish_kernel      logits [2A,1,T,T] -> 0.5 * (l + flip(l_mirror)) -> sigmoid -> bilinear align_corners=True to the
//                              region size -> pasted into a zero full-size map; also the thresholded mask for the clicker
//                              and the bbox of the > 0.5 pixels for the next click's region.
// Floating-point parity: the interpolation weights follow torch's formula (scale = (in-1)/(out-1) in fp32, index = scale * i,
// lambda = index - floor); an identity-sized region reproduces the input bit for bit; otherwise values agree with
// F.interpolate to ~1 ulp (tests gate 2e-6 absolute on [0,1] data).
#include "session.cuh"

#include <climits>

namespace vpu {

namespace {

constexpr int MAX_CLICKS_CAP = 64;

__device__ __forceinline__ int py_round(double x) { return (int)rint(x); }

__device__ __forceinline__ double seg_iou(int a, int b, int c, int d) {
    const int num = max(0, min(b, d) - max(a, c) + 1);
    const double den = fmax(1e-6, (double)(max(b, d) - min(a, c) + 1));
    return __ddiv_rn((double)num, den);
}

__global__ void __launch_bounds__(64) session_roi_points_kernel(SessionState st, const int32_t* __restrict__ active,
                                                               const int32_t* __restrict__ new_clicks, int A,
                                                               double* __restrict__ net_points) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ int s_roi[4];
    __shared__ int s_n;
    __shared__ int s_pos[MAX_CLICKS_CAP], s_neg[MAX_CLICKS_CAP];
    __shared__ int s_npos, s_nneg;
    const int a = blockIdx.x, s = active[a];
    int32_t* ck = st.clicks + (size_t)s * st.max_clicks * 3;
    if (threadIdx.x == 0) {
        int nc = st.nclicks[s];
        if (new_clicks != nullptr && nc < st.max_clicks) {
            ck[3 * nc] = new_clicks[4 * s];
            ck[3 * nc + 1] = new_clicks[4 * s + 1];
            ck[3 * nc + 2] = new_clicks[4 * s + 2];
            st.nclicks[s] = ++nc;
        }
        int np = 0, nn = 0;
        for (int i = 0; i < nc; ++i) {
            if (ck[3 * i]) s_pos[np++] = i; else s_neg[nn++] = i;
        }
        s_npos = np; s_nneg = nn; s_n = nc;
        // ---- ZoomIn.transform (zoom_in.py:30-66) ----
        int32_t* fb = st.fgbox + 5 * s;
        int r0, r1, c0, c1;
        if (fb[4] == 1) {                                                  // prev_probs exist and (prev_probs > 0.5).sum() > 0
            int bmin_r = fb[0], bmax_r = fb[1], bmin_c = fb[2], bmax_c = fb[3];
            for (int i = 0; i < np; ++i) {                                 // get_object_roi: positive clicks join the mask
                const int r = ck[3 * s_pos[i] + 1], c = ck[3 * s_pos[i] + 2];
                bmin_r = min(bmin_r, r); bmax_r = max(bmax_r, r); bmin_c = min(bmin_c, c); bmax_c = max(bmax_c, c);
            }
            const double rc = __dmul_rn(0.5, (double)(bmin_r + bmax_r)), cc = __dmul_rn(0.5, (double)(bmin_c + bmax_c));
            double hh = __dmul_rn(st.expansion_ratio, (double)(bmax_r - bmin_r + 1));
            double ww = __dmul_rn(st.expansion_ratio, (double)(bmax_c - bmin_c + 1));
            if (st.min_crop_size >= 0) { hh = fmax(hh, (double)st.min_crop_size); ww = fmax(ww, (double)st.min_crop_size); }
            r0 = max(0, py_round(__dsub_rn(rc, __dmul_rn(0.5, hh))));
            r1 = min(st.H - 1, py_round(__dadd_rn(rc, __dmul_rn(0.5, hh))));
            c0 = max(0, py_round(__dsub_rn(cc, __dmul_rn(0.5, ww))));
            c1 = min(st.W - 1, py_round(__dadd_rn(cc, __dmul_rn(0.5, ww))));
        } else {                                                           // skip_clicks < 0: the whole image
            r0 = 0; r1 = st.H - 1; c0 = 0; c1 = st.W - 1;
        }
        int32_t* roi = st.roi + 4 * s;
        bool replace = roi[0] < 0;
        if (!replace) {
            for (int i = 0; i < np && !replace; ++i) {                     // check_object_roi (zoom_in.py:171-183)
                const int r = ck[3 * s_pos[i] + 1], c = ck[3 * s_pos[i] + 2];
                if (!(roi[0] <= r && r < roi[1]) || !(roi[2] <= c && c < roi[3])) replace = true;
            }
            if (!replace)
                replace = __dmul_rn(seg_iou(r0, r1, roi[0], roi[1]), seg_iou(c0, c1, roi[2], roi[3])) < st.recompute_thresh_iou;
        }
        if (replace) { roi[0] = r0; roi[1] = r1; roi[2] = c0; roi[3] = c1; }
        s_roi[0] = roi[0]; s_roi[1] = roi[1]; s_roi[2] = roi[2]; s_roi[3] = roi[3];
        fb[0] = INT_MAX; fb[1] = -1; fb[2] = INT_MAX; fb[3] = -1; fb[4] = 0;   // re-armed for session_finish_kernel
    }
    __syncthreads();
    // ---- _transform_clicks (zoom_in.py:102-112), flip (flip.py:15-19), get_points_nd (base.py:195-213) ----
    const int nh = st.n_half, T = st.T;
    const double rh = (double)(s_roi[1] - s_roi[0] + 1), rw = (double)(s_roi[3] - s_roi[2] + 1);
    double* p0 = net_points + (size_t)a * 2 * nh * 3;
    double* p1 = net_points + (size_t)(A + a) * 2 * nh * 3;
    for (int j = threadIdx.x; j < 2 * nh; j += blockDim.x) {
        const bool posrow = j < nh;
        const int k = posrow ? j : j - nh;
        const int cnt = posrow ? s_npos : s_nneg;
        double y = -1.0, x = -1.0, xf = -1.0, id = -1.0;
        if (k < cnt) {
            const int i = posrow ? s_pos[k] : s_neg[k];
            const double yy = __ddiv_rn((double)(T * (ck[3 * i + 1] - s_roi[0])), rh);
            const double xx = __ddiv_rn((double)(T * (ck[3 * i + 2] - s_roi[2])), rw);
            y = yy;
            x = xx;
            xf = __dsub_rn(__dsub_rn((double)T, xx), 1.0);
            id = (double)i;
        }
        p0[3 * j] = y; p0[3 * j + 1] = x; p0[3 * j + 2] = id;
        p1[3 * j] = y; p1[3 * j + 1] = xf; p1[3 * j + 2] = id;
    }
}

// torch's bilinear source index for align_corners=True: scale = (in-1)/(out-1) (fp32), src = scale * i
struct Tap {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ Tap make_tap(int i, int in_size, int out_size) {
    const float scale = out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
    const float src = scale * (float)i;
    Tap t;
    t.i0 = min((int)src, in_size - 1);
    t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
    t.l1 = fminf(fmaxf(src - (float)t.i0, 0.f), 1.f);
    t.l0 = 1.f - t.l1;
    return t;
}
__device__ __forceinline__ float bilerp(float a, float b, float c, float d, const Tap& ty, const Tap& tx) {
    return ty.l0 * (tx.l0 * a + tx.l1 * b) + ty.l1 * (tx.l0 * c + tx.l1 * d);
}

__global__ void __launch_bounds__(256) session_crop_kernel(SessionState st, const int32_t* __restrict__ active, int A,
                                                           float* __restrict__ net_image) {
    pdl_launch_dependents();
    pdl_wait();
    const int a = blockIdx.y, s = active[a], T = st.T;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * T) return;
    const int y = idx / T, x = idx - y * T;
    const int32_t* roi = st.roi + 4 * s;
    const int rmin = roi[0], cmin = roi[2], rh = roi[1] - rmin + 1, rw = roi[3] - cmin + 1;
    const Tap ty = make_tap(y, rh, T), tx = make_tap(x, rw, T);
    const size_t plane = (size_t)st.H * st.W;
    const size_t o00 = (size_t)(rmin + ty.i0) * st.W + cmin + tx.i0, o01 = (size_t)(rmin + ty.i0) * st.W + cmin + tx.i1;
    const size_t o10 = (size_t)(rmin + ty.i1) * st.W + cmin + tx.i0, o11 = (size_t)(rmin + ty.i1) * st.W + cmin + tx.i1;
    const size_t tp = (size_t)T * T;
    float* d0 = net_image + (size_t)a * 4 * tp + (size_t)y * T + x;
    float* d1 = net_image + (size_t)(A + a) * 4 * tp + (size_t)y * T + (T - 1 - x);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float* src = c < 3 ? st.images + ((size_t)s * 3 + c) * plane : st.prev_probs + (size_t)s * plane;
        const float v = bilerp(__ldg(src + o00), __ldg(src + o01), __ldg(src + o10), __ldg(src + o11), ty, tx);
        d0[c * tp] = v;
        d1[c * tp] = v;
    }
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

constexpr int FINISH_PIX = 4;      // pixels per thread: 1024 per block, one set of bbox atomics per block

__global__ void __launch_bounds__(256) session_finish_kernel(SessionState st, const int32_t* __restrict__ active, int A,
                                                             const float* __restrict__ logits) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ int s_box[4][8];
    const int a = blockIdx.y, s = active[a], T = st.T;
    const int HW = st.H * st.W;
    const int32_t* roi = st.roi + 4 * s;
    const int rmin = roi[0], rmax = roi[1], cmin = roi[2], cmax = roi[3];
    const float* l0 = logits + (size_t)a * T * T;
    const float* l1 = logits + (size_t)(A + a) * T * T;
    int fy0 = INT_MAX, fy1 = -1, fx0 = INT_MAX, fx1 = -1;
#pragma unroll
    for (int u = 0; u < FINISH_PIX; ++u) {
        const int idx = (blockIdx.x * FINISH_PIX + u) * 256 + threadIdx.x;
        if (idx >= HW) break;
        const int Y = idx / st.W, X = idx - Y * st.W;
        float v = 0.f;
        if (Y >= rmin && Y <= rmax && X >= cmin && X <= cmax) {
            const Tap ty = make_tap(Y - rmin, T, rmax - rmin + 1), tx = make_tap(X - cmin, T, cmax - cmin + 1);
            // flip.py:21-28 averages the LOGITS of the image and of its mirror, base.py:147-151 applies the sigmoid next
            auto prob = [&](int yy, int xx) {
                return sigmoid_f(0.5f * (__ldg(l0 + (size_t)yy * T + xx) + __ldg(l1 + (size_t)yy * T + (T - 1 - xx))));
            };
            const float p00 = prob(ty.i0, tx.i0);
            const float p01 = tx.i1 != tx.i0 ? prob(ty.i0, tx.i1) : p00;
            const float p10 = ty.i1 != ty.i0 ? prob(ty.i1, tx.i0) : p00;
            const float p11 = ty.i1 != ty.i0 ? (tx.i1 != tx.i0 ? prob(ty.i1, tx.i1) : p10) : p01;
            v = bilerp(p00, p01, p10, p11, ty, tx);
        }
        st.prev_probs[(size_t)s * HW + idx] = v;
        st.pred[(size_t)s * HW + idx] = v > st.pred_thr ? 1 : 0;
        if (v > st.zoom_thr) { fy0 = min(fy0, Y); fy1 = max(fy1, Y); fx0 = min(fx0, X); fx1 = max(fx1, X); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        fy0 = min(fy0, __shfl_xor_sync(0xffffffffu, fy0, o));
        fy1 = max(fy1, __shfl_xor_sync(0xffffffffu, fy1, o));
        fx0 = min(fx0, __shfl_xor_sync(0xffffffffu, fx0, o));
        fx1 = max(fx1, __shfl_xor_sync(0xffffffffu, fx1, o));
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_box[0][w] = fy0; s_box[1][w] = fy1; s_box[2][w] = fx0; s_box[3][w] = fx1; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 1; i < 8; ++i) {
            fy0 = min(fy0, s_box[0][i]); fy1 = max(fy1, s_box[1][i]); fx0 = min(fx0, s_box[2][i]); fx1 = max(fx1, s_box[3][i]);
        }
        if (fy1 >= 0) {
            int32_t* fb = st.fgbox + 5 * s;
            atomicMin(fb + 0, fy0); atomicMax(fb + 1, fy1); atomicMin(fb + 2, fx0); atomicMax(fb + 3, fx1); atomicMax(fb + 4, 1);
        }
    }
}

// ToTensor of the predictor (isegm/inference/predictors/base.py:30,45: transforms.ToTensor() -> HWC uint8 / 255 as CHW fp32) for a
// whole batch on the device, so that a batch uploads 3 bytes per pixel instead of 12: image4[b, c] = u8[b, y, x, c] / 255 (IEEE
// division, the host's result bit for bit), image4[b, 3] = prev (or 0).  One thread per 4 pixels of a row: 12 bytes in, 4 x 16 out.
__global__ void __launch_bounds__(256) image_from_u8_kernel(const uint8_t* __restrict__ rgb, const float* __restrict__ prev,
                                                            float* __restrict__ image4, size_t HW4, int B) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // quad index inside one image
    const int b = blockIdx.y;
    if (i >= HW4) return;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(rgb + ((size_t)b * HW4 + i) * 12);
    const uint32_t w0 = src[0], w1 = src[1], w2 = src[2];
    const uint8_t px[12] = {(uint8_t)w0, (uint8_t)(w0 >> 8), (uint8_t)(w0 >> 16), (uint8_t)(w0 >> 24),
                            (uint8_t)w1, (uint8_t)(w1 >> 8), (uint8_t)(w1 >> 16), (uint8_t)(w1 >> 24),
                            (uint8_t)w2, (uint8_t)(w2 >> 8), (uint8_t)(w2 >> 16), (uint8_t)(w2 >> 24)};
    float* dst = image4 + (size_t)b * 16 * HW4 + 4 * i;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float4 v;
        v.x = __fdiv_rn((float)px[c], 255.f); v.y = __fdiv_rn((float)px[3 + c], 255.f);
        v.z = __fdiv_rn((float)px[6 + c], 255.f); v.w = __fdiv_rn((float)px[9 + c], 255.f);
        *reinterpret_cast<float4*>(dst + (size_t)c * 4 * HW4) = v;
    }
    float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (prev) pv = *reinterpret_cast<const float4*>(prev + (size_t)b * 4 * HW4 + 4 * i);
    *reinterpret_cast<float4*>(dst + (size_t)12 * HW4) = pv;
}

int check_state(const SessionState& st, const int32_t* active, int A) {
    VPU_REQUIRE(st.S > 0 && st.H > 0 && st.W > 0 && st.T > 1 && st.H <= 8192 && st.W <= 8192, "session: bad sizes S=%d H=%d W=%d T=%d",
                st.S, st.H, st.W, st.T);
    VPU_REQUIRE(st.max_clicks >= 1 && st.max_clicks <= MAX_CLICKS_CAP, "session: max_clicks %d outside [1, %d]", st.max_clicks, MAX_CLICKS_CAP);
    VPU_REQUIRE(st.n_half >= 1 && st.n_half >= st.max_clicks, "session: n_half %d must hold max_clicks %d clicks of one kind", st.n_half,
                st.max_clicks);
    VPU_REQUIRE(st.images && st.prev_probs && st.pred && st.clicks && st.nclicks && st.roi && st.fgbox, "session: null state pointer");
    VPU_REQUIRE(active && A >= 1 && A <= st.S, "session: active list of %d sessions (S = %d)", A, st.S);
    return 0;
}

}  // namespace

int session_prepare_launch(const SessionState& st, const int32_t* active, int A, const int32_t* new_clicks, float* net_image,
                           double* net_points, cudaStream_t stream) {
    if (int rc = check_state(st, active, A)) return rc;
    VPU_REQUIRE(net_image && net_points, "session_prepare: null output");
    VPU_CHECK_CUDA(launch_pdl(session_roi_points_kernel, dim3(A), dim3(64), 0, stream, st, active, new_clicks, A, net_points));
    VPU_CHECK_CUDA(launch_pdl(session_crop_kernel, dim3((st.T * st.T + 255) / 256, A), dim3(256), 0, stream, st, active, A, net_image));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch(2);
    return 0;
}

int image_from_u8_launch(const uint8_t* rgb, const float* prev, float* image4, int B, int H, int W, cudaStream_t stream) {
    VPU_REQUIRE(rgb && image4 && B > 0 && H > 0 && W > 0, "image_from_u8: null argument");
    VPU_REQUIRE(((size_t)H * W) % 4 == 0, "image_from_u8: H*W must be a multiple of 4");
    VPU_REQUIRE((reinterpret_cast<uintptr_t>(rgb) & 3) == 0 && (reinterpret_cast<uintptr_t>(image4) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(prev) & 15) == 0, "image_from_u8: unaligned pointer");
    const size_t HW4 = (size_t)H * W / 4;
    VPU_CHECK_CUDA(launch_pdl(image_from_u8_kernel, dim3((unsigned)((HW4 + 255) / 256), B), dim3(256), 0, stream, rgb, prev, image4, HW4, B));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch(1);
    return 0;
}

int session_finish_launch(const SessionState& st, const int32_t* active, int A, const float* logits, cudaStream_t stream) {
    if (int rc = check_state(st, active, A)) return rc;
    VPU_REQUIRE(logits, "session_finish: null logits");
    VPU_CHECK_CUDA(launch_pdl(session_finish_kernel, dim3((st.H * st.W + 256 * FINISH_PIX - 1) / (256 * FINISH_PIX), A), dim3(256), 0, stream, st, active, A, logits));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch(1);
    return 0;
}

}  // namespace vpu
