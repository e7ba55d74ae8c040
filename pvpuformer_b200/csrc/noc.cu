// Device-side oracle clicker + IoU tally of the NoC evaluation protocol.
// Reference: isegm/inference/clicker.py:29-69 (Clicker._get_next_click) and isegm/inference/utils.py:80-87 (get_iou).
//
// Per click session: false-negative / false-positive masks of the current prediction against the ground truth, each
// zero-padded by one pixel and distance-transformed (cv2.distanceTransform(DIST_L2, mask 0) = exact Euclidean distance to
// the nearest zero pixel); already-clicked pixels are excluded; the region with the larger maximum distance wins
// (positive click iff FN strictly larger) and ties resolve to the first maximum in row-major order.
// cv2 returns float32 sqrt of the exact integer squared distance; sqrt is monotone and, below 2^24 / for images up to
// ~2000 px, injective on those integers, so every comparison of the reference is reproduced exactly on the SQUARED
// integer distances: no floating point anywhere in this file.
// Scope of that claim (measured against cv2 4.13 in the build container): for masks of >= 2e4 pixels cv2 returns exactly the
// correctly rounded float sqrt of the exact squared distance (0 deviating pixels over 6.5e6 tested); for smaller masks
// (e.g. 64x64, 100x132) its small-image code path deviates by +-1 ulp at a few pixels, so two pixels with the SAME exact
// distance can compare unequal there and the reference's tie order is an artefact of that path.  The Python wrapper
// (inference/evaluation.py) therefore refuses the device clicker below 2e4 pixels.
//
// Exact EDT as two separable passes: columns (distance to the nearest zero along the column, virtual zeros at rows -1 and H),
// then rows: d2(y, x) = min_x' (x - x')^2 + g(y, x')^2 with g = 0 at the virtual columns -1 and W; the search radius is
// bounded by the running minimum.  All sessions of a micro-batch go through three launches; the host reads back 16 bytes
// per session (the click) and two integers (the IoU counts) instead of a full-resolution probability map, and runs no
// distance transform.
#include "noc.cuh"

namespace vpu {

namespace {

constexpr int ROW_THREADS = 256;

// plane p < S: false negatives of session p; p >= S: false positives of session p - S.  gt: 1 object, 0 background, -1 ignore
__device__ __forceinline__ bool err_at(const int8_t* gt, const uint8_t* pred, size_t i, bool fn) {
    const int g = gt[i];
    const bool p = pred[i] != 0;
    // clicker.py:31-32: fn = (gt == 1) & !pred ; fp = !(gt == 1) & pred, both restricted to gt != -1 (any other label is background)
    return fn ? (g == 1 && !p) : (g != 1 && g != -1 && p);
}

__global__ void __launch_bounds__(128) noc_cols_kernel(const int8_t* __restrict__ gt, const uint8_t* __restrict__ pred, int S, int H,
                                                       int W, uint16_t* __restrict__ g, unsigned long long* __restrict__ keys,
                                                       long long* __restrict__ counts) {
    pdl_launch_dependents();
    pdl_wait();
    const int p = blockIdx.y, s = p < S ? p : p - S, x = blockIdx.x * blockDim.x + threadIdx.x;
    const bool fn = p < S;
    if (blockIdx.x == 0 && threadIdx.x == 0) keys[p] = 0xFFFFFFFFull;      // d2 = 0 at index 0: the argmax of an all-zero map
    long long inter = 0, uni = 0;
    if (x < W) {
        const size_t base = (size_t)s * H * W + x;
        uint16_t* gp = g + (size_t)p * H * W + x;
        int run = 0;                                   // distance to the virtual zero at row -1
        for (int y = 0; y < H; ++y) {
            const size_t i = base + (size_t)y * W;
            run = err_at(gt, pred, i, fn) ? run + 1 : 0;
            gp[(size_t)y * W] = (uint16_t)run;
            if (fn) {                                  // IoU counts once per session (reference utils.py:80-87)
                const int gv = gt[i];
                const bool pr = pred[i] != 0;
                inter += (pr && gv == 1);
                uni += ((pr || gv == 1) && gv != -1);
            }
        }
        run = 0;                                       // virtual zero at row H
        for (int y = H - 1; y >= 0; --y) {
            const int d = gp[(size_t)y * W];
            run = d ? run + 1 : 0;
            if (run < d) gp[(size_t)y * W] = (uint16_t)run;
        }
    }
    if (fn) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            inter += __shfl_xor_sync(0xffffffffu, inter, o);
            uni += __shfl_xor_sync(0xffffffffu, uni, o);
        }
        if ((threadIdx.x & 31) == 0 && (inter | uni)) {
            atomicAdd(reinterpret_cast<unsigned long long*>(counts) + 2 * s, (unsigned long long)inter);
            atomicAdd(reinterpret_cast<unsigned long long*>(counts) + 2 * s + 1, (unsigned long long)uni);
        }
    }
}

// one block per (row y, plane p): exact squared distance of every error pixel of the row, masked by not_clicked, reduced to
// key = d2 << 32 | ~index (larger distance first, then smaller row-major index)
__global__ void __launch_bounds__(ROW_THREADS) noc_rows_kernel(const uint16_t* __restrict__ g, const uint8_t* __restrict__ not_clicked,
                                                               int S, int H, int W, unsigned long long* __restrict__ keys) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ uint16_t row[];                  // g(y, -1 .. W): W + 2 entries, the two virtual columns are 0
    __shared__ unsigned long long red[ROW_THREADS / 32];
    __shared__ int any_s;
    const int y = blockIdx.x, p = blockIdx.y, s = p < S ? p : p - S;
    const uint16_t* gp = g + ((size_t)p * H + y) * W;
    if (threadIdx.x == 0) { row[0] = 0; row[W + 1] = 0; any_s = 0; }
    __syncthreads();
    int any = 0;
    for (int x = threadIdx.x; x < W; x += ROW_THREADS) {
        const uint16_t v = gp[x];
        row[x + 1] = v;
        any |= v;
    }
    if (any) any_s = 1;
    __syncthreads();
    if (!any_s) return;                                // no error pixel in this row
    unsigned long long best_key = 0;
    const uint8_t* nc = not_clicked + ((size_t)s * H + y) * W;
    for (int x = threadIdx.x; x < W; x += ROW_THREADS) {
        const unsigned gx = row[x + 1];
        if (gx == 0 || nc[x] == 0) continue;
        unsigned best = gx * gx;
        for (int r = 1; (unsigned)(r * r) < best; ++r) {
            const int xl = x + 1 - r, xr = x + 1 + r;  // indices into row[]; clamp to the virtual zero columns
            const unsigned gl = xl >= 0 ? row[xl] : 0u, gr = xr <= W + 1 ? row[xr] : 0u;
            // beyond the virtual columns nothing can beat the virtual zero itself, which a smaller r has already covered
            const unsigned cl = (unsigned)(r * r) + (xl >= 0 ? gl * gl : 0u), cr = (unsigned)(r * r) + (xr <= W + 1 ? gr * gr : 0u);
            best = min(best, min(cl, cr));
        }
        const unsigned long long key = ((unsigned long long)best << 32) | (0xFFFFFFFFu - (unsigned)(y * W + x));
        best_key = key > best_key ? key : best_key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best_key, o);
        best_key = other > best_key ? other : best_key;
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best_key;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long k = 0;
        for (int w = 0; w < ROW_THREADS / 32; ++w) k = red[w] > k ? red[w] : k;
        if (k) atomicMax(keys + p, k);
    }
}

__global__ void noc_select_kernel(const unsigned long long* __restrict__ keys, uint8_t* __restrict__ not_clicked, int S, int H, int W,
                                  int32_t* __restrict__ clicks) {
    pdl_launch_dependents();
    pdl_wait();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const unsigned long long fn = keys[s], fp = keys[S + s];
    const bool positive = (fn >> 32) > (fp >> 32);                         // FN strictly larger (clicker.py:52-53)
    const unsigned long long k = positive ? fn : fp;
    const unsigned idx = 0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull);
    const int y = (int)(idx / (unsigned)W), x = (int)(idx % (unsigned)W);
    clicks[4 * s] = positive ? 1 : 0;
    clicks[4 * s + 1] = y;
    clicks[4 * s + 2] = x;
    clicks[4 * s + 3] = (int32_t)(k >> 32);                                // squared distance of the click (diagnostics / tests)
    not_clicked[((size_t)s * H + y) * W + x] = 0;                          // clicker.py:84-85
}

}  // namespace

size_t noc_workspace_bytes(int S, int H, int W) {
    return (size_t)2 * S * H * W * sizeof(uint16_t) + (size_t)2 * S * sizeof(unsigned long long) + 256;
}

int noc_next_clicks_launch(const int8_t* gt, const uint8_t* pred, uint8_t* not_clicked, int S, int H, int W, int32_t* clicks,
                           long long* iou_counts, void* workspace, cudaStream_t stream) {
    VPU_REQUIRE(S > 0 && H > 0 && W > 0 && H <= 4096 && W <= 4096 && S <= 32767, "clicker: bad sizes S=%d H=%d W=%d", S, H, W);
    uint16_t* g = reinterpret_cast<uint16_t*>(workspace);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(workspace) +
                                                                     (((size_t)2 * S * H * W * sizeof(uint16_t) + 255) & ~(size_t)255));
    VPU_CHECK_CUDA(cudaMemsetAsync(iou_counts, 0, (size_t)2 * S * sizeof(long long), stream));
    VPU_CHECK_CUDA(launch_pdl(noc_cols_kernel, dim3((W + 127) / 128, 2 * S), dim3(128), 0, stream, gt, pred, S, H, W, g, keys, iou_counts));
    VPU_CHECK_CUDA(launch_pdl(noc_rows_kernel, dim3(H, 2 * S), dim3(ROW_THREADS), (size_t)(W + 2) * sizeof(uint16_t), stream,
                              static_cast<const uint16_t*>(g), static_cast<const uint8_t*>(not_clicked), S, H, W, keys));
    VPU_CHECK_CUDA(launch_pdl(noc_select_kernel, dim3((S + 63) / 64), dim3(64), 0, stream, static_cast<const unsigned long long*>(keys),
                              not_clicked, S, H, W, clicks));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch(3);
    return 0;
}

}  // namespace vpu
