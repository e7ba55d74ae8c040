// Box / scribble prompt rasteriser: the 3-pixel outline the reference draws into the click planes
// with cv2.rectangle(img, (x0, y0), (x1, y1), 255, 3) and cv2.polylines(img, [curve], False, 255, 3)
// (isegm/model/is_model.py:97-146), bit-exact for vertices anywhere (inside or outside the image).
//
// The arithmetic is OpenCV's (modules/imgproc/src/drawing.cpp, 4.x; the reference pins opencv-python 4.7.0.68, this image has
// 4.13 -- same code path): a thick line of thickness t between integer points is
//   * the convex quadrilateral p0 +- d, p1 +- d with d = round((dy, dx) * r), r = (t * 2^15 + (t & 1) * 2^15) / |p1 - p0| in 16.16
//     fixed point (half-width 2 px for t = 3), filled by FillConvexPoly: its four edges drawn with the fixed-point DDA
//     Line2 (after Cohen-Sutherland clipping to the image in 16.16 coordinates), then a scan conversion that walks the left
//     and right edge with a per-row increment ((xe - xs) * 2 + dy) / (2 * dy) and fills [x_left + 1/2, x_right + 1/2] >> 16;
//   * a filled midpoint circle of radius (t * 2^15 + 2^15) >> 16 = 2 (13 pixels) at the end point, and at the start point of the
//     first segment of an open polyline.
// cv2.rectangle = closed polyline through (x0,y0) (x1,y0) (x1,y1) (x0,y1); a zero-length segment draws only its circle(s).
// One thread rasterises one segment; all threads store the same value, so overlapping stores need no ordering.
// Validated against cv2 in tests/test_kernels_gpu.py (vertices inside the image, on its border and up to 60 px outside).
#include "raster.cuh"

namespace vpu {

namespace {

typedef long long i64;
constexpr int XY_SHIFT = 16;
constexpr i64 XY_ONE = 1 << XY_SHIFT;

struct Plane {
    uint8_t* p;
    int W, H;
    __device__ __forceinline__ void put(i64 x, i64 y) const {
        if (x >= 0 && x < W && y >= 0 && y < H) p[(size_t)y * W + (int)x] = 1;
    }
    __device__ __forceinline__ void hline(int y, int x1, int x2) const {     // caller guarantees 0 <= y < H and clipped x
        uint8_t* r = p + (size_t)y * W;
        for (int x = x1; x <= x2; ++x) r[x] = 1;
    }
};

// cv::clipLine(Size2l, Point2l&, Point2l&): Cohen-Sutherland with truncating 64-bit divisions
__device__ bool clip_line(i64 w, i64 h, i64& x1, i64& y1, i64& x2, i64& y2) {
    const i64 right = w - 1, bottom = h - 1;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        i64 a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (a - y1) * (x2 - x1) / (y2 - y1);
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (a - y2) * (x2 - x1) / (y2 - y1);
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (a - x1) * (y2 - y1) / (x2 - x1);
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (a - x2) * (y2 - y1) / (x2 - x1);
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// Line2: fixed-point DDA between two 16.16 points
__device__ void line2(const Plane& im, i64 x1, i64 y1, i64 x2, i64 y2) {
    if (!clip_line((i64)im.W << XY_SHIFT, (i64)im.H << XY_SHIFT, x1, y1, x2, y2)) return;
    i64 dx = x2 - x1, dy = y2 - y1;
    const i64 ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
    i64 x_step, y_step;
    int ecount;
    if (ax > ay) {
        if (dx < 0) { dy = -dy; i64 t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
        x_step = XY_ONE;
        y_step = (dy << XY_SHIFT) / (ax | 1);
        ecount = (int)((x2 - x1) >> XY_SHIFT);
    } else {
        if (dy < 0) { dx = -dx; i64 t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
        x_step = (dx << XY_SHIFT) / (ay | 1);
        y_step = XY_ONE;
        ecount = (int)((y2 - y1) >> XY_SHIFT);
    }
    x1 += XY_ONE >> 1;
    y1 += XY_ONE >> 1;
    im.put((x2 + (XY_ONE >> 1)) >> XY_SHIFT, (y2 + (XY_ONE >> 1)) >> XY_SHIFT);
    if (ax > ay) {
        x1 >>= XY_SHIFT;
        for (; ecount >= 0; --ecount) { im.put(x1, y1 >> XY_SHIFT); ++x1; y1 += y_step; }
    } else {
        y1 >>= XY_SHIFT;
        for (; ecount >= 0; --ecount) { im.put(x1 >> XY_SHIFT, y1); x1 += x_step; ++y1; }
    }
}

// FillConvexPoly for 4 vertices in 16.16 coordinates (shift == XY_SHIFT, line_type 8)
__device__ void fill_quad(const Plane& im, const i64 (&vx)[4], const i64 (&vy)[4]) {
    constexpr int npts = 4;
    constexpr i64 delta = XY_ONE >> 1;
    i64 xmin = vx[0], xmax = vx[0], ymin = vy[0], ymax = vy[0];
    int imin = 0;
    i64 px = vx[npts - 1], py = vy[npts - 1];
    for (int i = 0; i < npts; ++i) {
        if (vy[i] < ymin) { ymin = vy[i]; imin = i; }
        ymax = vy[i] > ymax ? vy[i] : ymax;
        xmax = vx[i] > xmax ? vx[i] : xmax;
        xmin = vx[i] < xmin ? vx[i] : xmin;
        line2(im, px, py, vx[i], vy[i]);
        px = vx[i]; py = vy[i];
    }
    xmin = (xmin + delta) >> XY_SHIFT; xmax = (xmax + delta) >> XY_SHIFT;
    ymin = (ymin + delta) >> XY_SHIFT; ymax = (ymax + delta) >> XY_SHIFT;
    if (xmax < 0 || ymax < 0 || xmin >= im.W || ymin >= im.H) return;
    if (ymax > im.H - 1) ymax = im.H - 1;
    int e_idx[2] = {imin, imin}, e_di[2] = {1, npts - 1}, e_ye[2] = {(int)ymin, (int)ymin};
    i64 e_x[2] = {-XY_ONE, -XY_ONE}, e_dx[2] = {0, 0};
    int y = (int)ymin, edges = npts;
    do {
        for (int i = 0; i < 2; ++i) {
            if (y >= e_ye[i]) {
                int idx0 = e_idx[i];
                const int di = e_di[i];
                int idx = idx0 + di;
                if (idx >= npts) idx -= npts;
                for (; edges-- > 0;) {
                    const int ty = (int)((vy[idx] + delta) >> XY_SHIFT);
                    if (ty > y) {
                        const i64 xs = vx[idx0], xe = vx[idx];
                        e_ye[i] = ty;
                        e_dx[i] = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y));
                        e_x[i] = xs;
                        e_idx[i] = idx;
                        break;
                    }
                    idx0 = idx;
                    idx += di;
                    if (idx >= npts) idx -= npts;
                }
            }
        }
        if (edges < 0) break;
        if (y >= 0) {
            const int l = e_x[0] > e_x[1] ? 1 : 0, r = 1 - l;
            int xx1 = (int)((e_x[l] + delta) >> XY_SHIFT), xx2 = (int)((e_x[r] + delta) >> XY_SHIFT);
            if (xx2 >= 0 && xx1 < im.W) {
                if (xx1 < 0) xx1 = 0;
                if (xx2 >= im.W) xx2 = im.W - 1;
                im.hline(y, xx1, xx2);
            }
        }
        e_x[0] += e_dx[0];
        e_x[1] += e_dx[1];
    } while (++y <= (int)ymax);
}

// Circle(..., fill = 1): midpoint circle, horizontal spans
__device__ void fill_circle(const Plane& im, int cx, int cy, int radius) {
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    while (dx >= dy) {
        const int y11 = cy - dy, y12 = cy + dy, y21 = cy - dx, y22 = cy + dx;
        int x11 = cx - dx, x12 = cx + dx, x21 = cx - dy, x22 = cx + dy;
        if (x11 < im.W && x12 >= 0 && y21 < im.H && y22 >= 0) {
            x11 = x11 > 0 ? x11 : 0;
            x12 = x12 < im.W - 1 ? x12 : im.W - 1;
            if ((unsigned)y11 < (unsigned)im.H) im.hline(y11, x11, x12);
            if ((unsigned)y12 < (unsigned)im.H) im.hline(y12, x11, x12);
            if (x21 < im.W && x22 >= 0) {
                x21 = x21 > 0 ? x21 : 0;
                x22 = x22 < im.W - 1 ? x22 : im.W - 1;
                if ((unsigned)y21 < (unsigned)im.H) im.hline(y21, x21, x22);
                if ((unsigned)y22 < (unsigned)im.H) im.hline(y22, x21, x22);
            }
        }
        ++dy;
        err += plus;
        plus += 2;
        const int mask = (err <= 0) - 1;
        err -= minus & mask;
        dx += mask;
        minus -= mask & 2;
    }
}

// ThickLine for integer end points (shift 0), thickness > 1, line_type 8; flags: 1 = cap at p0, 2 = cap at p1.
// The segment is first clipped (integer Cohen-Sutherland) to the image rectangle grown by `thickness` pixels on every side --
// measured against cv2 4.13: segments are reproduced bit for bit for end points anywhere with exactly this margin, and only
// with it -- and the quadrilateral and the caps are built on the clipped end points; a segment outside that rectangle draws
// nothing.
__device__ void thick_line(const Plane& im, int ax, int ay, int bx, int by, int thickness, int flags) {
    {
        i64 x1 = (i64)ax + thickness, y1 = (i64)ay + thickness, x2 = (i64)bx + thickness, y2 = (i64)by + thickness;
        if (!clip_line((i64)im.W + 2 * thickness, (i64)im.H + 2 * thickness, x1, y1, x2, y2)) return;
        ax = (int)x1 - thickness; ay = (int)y1 - thickness; bx = (int)x2 - thickness; by = (int)y2 - thickness;
    }
    i64 p0x = (i64)ax << XY_SHIFT, p0y = (i64)ay << XY_SHIFT;
    const i64 p1x = (i64)bx << XY_SHIFT, p1y = (i64)by << XY_SHIFT;
    const double inv = 1.0 / (double)XY_ONE;
    const double dx = (double)(p0x - p1x) * inv, dy = (double)(p1y - p0y) * inv;
    double r = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    const int odd = thickness & 1;
    const i64 th = (i64)thickness << (XY_SHIFT - 1);
    if (fabs(r) > 2.220446049250313e-16) {
        r = __ddiv_rn((double)th + (double)odd * (double)XY_ONE * 0.5, sqrt(r));
        const i64 dpx = (i64)rint(__dmul_rn(dy, r)), dpy = (i64)rint(__dmul_rn(dx, r));     // cvRound: half to even
        const i64 vx[4] = {p0x + dpx, p0x - dpx, p1x - dpx, p1x + dpx};
        const i64 vy[4] = {p0y + dpy, p0y - dpy, p1y - dpy, p1y + dpy};
        fill_quad(im, vx, vy);
    }
    const int radius = (int)((th + (XY_ONE >> 1)) >> XY_SHIFT);
    for (int i = 0; i < 2; ++i) {
        if (flags & (i + 1)) fill_circle(im, (int)((p0x + (XY_ONE >> 1)) >> XY_SHIFT), (int)((p0y + (XY_ONE >> 1)) >> XY_SHIFT), radius);
        p0x = p1x; p0y = p1y;
    }
}

__global__ void __launch_bounds__(128) raster_prompts_kernel(RasterArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y, seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (a.type == 1) {                                                    // is_model.py:97-121
        if (seg >= 4) return;
        const int32_t* bx = a.boxes + 5 * b;
        const int xc = bx[0], yc = bx[1], w = bx[2], h = bx[3], slot = bx[4];
        // Python floor division of the reference (w // 2 on non-negative extents; negative extents floor as Python does)
        const int hw = (w >= 0 ? w : w - 1) / 2, hh = (h >= 0 ? h : h - 1) / 2;
        const int x0 = xc - hw, x1 = xc + hw, y0 = yc - hh, y1 = yc + hh;
        const int vx[4] = {x0, x1, x1, x0}, vy[4] = {y0, y0, y1, y1};
        const int p = (seg + 3) & 3;                                      // closed polyline: segment `seg` runs v[seg-1] -> v[seg]
        Plane im{a.planes + ((size_t)b * 2 + (slot < a.n ? 0 : 1)) * a.size * a.size, a.size, a.size};
        thick_line(im, vx[p], vy[p], vx[seg], vy[seg], 3, 2);
    } else {                                                              // is_model.py:123-146: always plane 0
        if (seg >= a.S - 1) return;
        const int32_t* s = a.scribbles + ((size_t)b * a.S + seg) * 2;
        Plane im{a.planes + (size_t)b * 2 * a.size * a.size, a.size, a.size};
        thick_line(im, s[0], s[1], s[2], s[3], 3, seg == 0 ? 3 : 2);
    }
}

}  // namespace

int raster_prompts_launch(const RasterArgs& a, int B, cudaStream_t stream) {
    VPU_REQUIRE(a.type == 1 || a.type == 2, "raster: type must be 1 (box) or 2 (scribble)");
    VPU_REQUIRE(a.planes && B > 0 && a.size > 0 && a.size <= 16384, "raster: bad planes / sizes");
    VPU_REQUIRE(a.type == 1 ? a.boxes != nullptr : (a.scribbles != nullptr && a.S >= 1), "raster: missing prompt array");
    VPU_CHECK_CUDA(cudaMemsetAsync(a.planes, 0, (size_t)B * 2 * a.size * a.size, stream));
    const int nseg = a.type == 1 ? 4 : a.S - 1;
    if (nseg > 0) {
        VPU_CHECK_CUDA(launch_pdl(raster_prompts_kernel, dim3((nseg + 127) / 128, B), dim3(128), 0, stream, a));
        VPU_CHECK_CUDA(cudaGetLastError());
        count_launch();
    } else {
        // a one-point polyline draws nothing in cv2.polylines (no segment): planes stay zero
    }
    return 0;
}

}  // namespace vpu
