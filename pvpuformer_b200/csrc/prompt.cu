// Prompt-side kernels (HBM-bound, fp32 / integer exact):
//   ppue_kernel          Probabilistic Prompt-unified Encoder rows [B, 48, 899]
//                        (reference is_vpu_model.py:189-352 + ops.py:39-325)
//   coord_features_kernel cat(prev_mask, click disks | box/scribble raster) [B, 3, H, W]
//                        (reference is_model.py:78-95 + ops.py:347-382, use_disks=True)
//   patch_operand_kernel the same planes + the raw RGB image written directly as the bf16
//                        A-operand [B*N, 6*p*p] of the fused (image + coords) patch-embed GEMM
//                        (reference models_vit.py:94-104, is_vpu_model.py:385-386), so the
//                        disk maps never exist in HBM on the forward path.
#include "prompt.cuh"

namespace vpu {

__device__ __forceinline__ int floordiv(int a, int b) {
    int q = a / b, r = a % b;
    return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}
__device__ __forceinline__ bool in_img(int x, int y, int w, int h) {  // ops.py:63-67 ('>' not '>=')
    return !((x < 0) || (x > w) || (y < 0) || (y > h));
}

// One CTA per output row (b, j).  Row = [vec_a(S) | vec_b(S) | label(3)].
__global__ void __launch_bounds__(128) ppue_kernel(const PpueArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const int j = blockIdx.x, b = blockIdx.y;
    const int S = a.size, D = 2 * S + 3, n = a.n, NM = a.num_max_points;
    const int half = j / NM, slot = j % NM;
    float* out = a.out + ((size_t)b * 2 * NM + j) * D;
    __nv_bfloat16* outb = a.out_bf16 ? a.out_bf16 + ((size_t)b * 2 * NM + j) * a.ld_bf16 : nullptr;

    // kind: 0 not-a-point, 1 click, 2 box, 3 scribble, 4 all-zero vectors with label
    int kind = 0, lab = 2;
    int p0 = 0, p1 = 0, r0 = 0, r1 = 0;       // centres and radii per axis
    double two_s0 = 1.0, two_s1 = 1.0;        // 2*sigma^2 per axis (box)
    if (slot < n) {
        const int i = half * n + slot;
        const double* pt = a.points + ((size_t)b * 2 * n + i) * 3;
        lab = half;
        if (a.type == 1 && a.boxes[b * 5 + 4] == i) {
            const int* bx = a.boxes + b * 5;
            lab = bx[4] < n ? 0 : 1;                                   // is_vpu_model.py:270-272
            kind = 4;
            if (bx[0] + bx[1] + bx[2] + bx[3] != 0) {                  // ops.py:142
                const int ksw = floordiv(bx[2], 2) * 2 - 1, ksh = floordiv(bx[3], 2) * 2 - 1;
                r0 = floordiv(ksw - 1, 2);
                r1 = floordiv(ksh - 1, 2);
                const int s0 = floordiv(r0, 3), s1 = floordiv(r1, 3);
                if (s0 != 0 && s1 != 0) {                              // ops.py:150,162
                    p0 = bx[0];
                    p1 = bx[1];
                    two_s0 = 2.0 * s0 * s0;
                    two_s1 = 2.0 * s1 * s1;
                    kind = 2;
                }
            }
        } else if (a.type == 2 && a.scrib_slot[b] == i) {
            kind = 3;
            lab = 0;
        } else if (pt[2] == -1.0) {
            kind = 0;
            lab = 2;
        } else {
            kind = 1;
            p0 = (int)pt[0];                                           // astype('int32'): trunc
            p1 = (int)pt[1];
            r0 = r1 = a.click_radius;
        }
        if (kind == 1 || kind == 2) {                                  // ops.py:90-94 / 178-182
            const bool ul = in_img(p0 - r0, p1 - r1, S, S), br = in_img(p0 + r0 + 1, p1 + r1 + 1, S, S);
            if (!ul && !br) kind = 4;
        }
    }
    for (int c = threadIdx.x; c < (outb ? a.ld_bf16 : D); c += blockDim.x) {
        float v = 0.f;
        if (c < 2 * S) {
            const int axis = c >= S, i = c - axis * S;
            if (kind == 1) {
                const int p = axis ? p1 : p0;
                const int ti = i - (p - r0);
                if (ti >= 0 && ti <= 2 * r0) v = a.click_table[ti];
            } else if (kind == 2) {
                const int p = axis ? p1 : p0, r = axis ? r1 : r0;
                const int dlt = i - p;
                if (dlt >= -r && dlt <= r) v = (float)exp(-((double)dlt * dlt) / (axis ? two_s1 : two_s0));
            } else if (kind == 3) {
                const int off = a.scrib_sel[((size_t)b * 2 + axis) * S + i];
                if (off != INT_MIN) v = (float)exp(-((double)off * off) / 18.0);   // sigma = 3
            }
        } else if (c < D) {
            v = (c - 2 * S == lab) ? 1.f : 0.f;
        }
        if (c < D) out[c] = v;
        if (outb) outb[c] = __float2bfloat16(v);
    }
}

int ppue_launch(const PpueArgs& a, int B, cudaStream_t stream) {
    VPU_REQUIRE(a.n >= 1 && a.n <= a.num_max_points, "PPuE: n=%d outside [1, %d] (reference supports n <= num_max_points)",
                a.n, a.num_max_points);
    VPU_REQUIRE(a.type == 0 || (a.type == 1 && a.boxes) || (a.type == 2 && a.scrib_sel && a.scrib_slot),
                "PPuE: prompt type %d needs its side inputs", a.type);
    VPU_REQUIRE(a.click_radius >= 0 && 2 * a.click_radius + 1 <= 32, "PPuE: click table too large");
    dim3 grid(2 * a.num_max_points, B);
    VPU_CHECK_CUDA(launch_pdl(ppue_kernel, dim3(grid), dim3(128), 0, stream, a));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------------------
// disks: reference ops.py:352-375.  d = fl32(fl32(r - pr)^2 + fl32(c - pc)^2), valid iff
// max(pr, pc) >= 0, disk = any(d <= R^2).  The difference is formed in double and rounded once
// to fp32 (what torch's in-place add_ on a float32 grid does with float64 clicks).
// ------------------------------------------------------------------------------------------
struct SmemPoints {
    double pc[2 * 24];      // column of the clicks that can reach this image row
    float dr2[2 * 24];      // fl32(fl32(r - pr)^2) of the same clicks
    int cnt[2];
};

// A CTA works on one image row r: only clicks with fl32(r - pr)^2 <= R^2 can produce d <= R^2 on it (the second term is
// non-negative and fp32 addition is monotone), so each half's valid clicks are filtered once per CTA -- typically 0 or 1 of up
// to 24 survive -- and the row term is computed once.  The per-pixel loop used to evaluate two double-precision differences per
// click and pixel (the reference's arithmetic: difference in double, rounded once to fp32); the column difference still is.
__device__ __forceinline__ void load_points(SmemPoints& sp, const double* pts, int n, int r, float r2) {
    if (threadIdx.x < 2) {
        const int h = threadIdx.x;
        int k = 0;
        for (int i = 0; i < n; ++i) {
            const double* p = pts + (size_t)(h * n + i) * 3;
            if (fmax(p[0], p[1]) >= 0.0) {
                const float dr = (float)((double)r - p[0]);
                const float d2 = __fmul_rn(dr, dr);
                if (d2 <= r2) { sp.pc[h * 24 + k] = p[1]; sp.dr2[h * 24 + k] = d2; ++k; }
            }
        }
        sp.cnt[h] = k;
    }
}
__device__ __forceinline__ bool disk_hit(const SmemPoints& sp, int h, int c, float r2) {
    bool hit = false;
    for (int k = 0; k < sp.cnt[h]; ++k) {
        const float dc = (float)((double)c - sp.pc[h * 24 + k]);
        const float d = __fadd_rn(sp.dr2[h * 24 + k], __fmul_rn(dc, dc));
        hit |= (d <= r2);
    }
    return hit;
}

// grid (H, B), block 256: one image row per CTA.
__global__ void __launch_bounds__(256) coord_features_kernel(const CoordArgs a, float* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ SmemPoints sp;
    const int y = blockIdx.x, b = blockIdx.y, H = a.H, W = a.W;
    const float r2 = a.radius * a.radius;
    load_points(sp, a.points + (size_t)b * 2 * a.n * 3, a.n, y, r2);
    __syncthreads();
    const float* prev = a.image4 + (((size_t)b * 4 + 3) * H + y) * W;
    float* o = out + ((size_t)b * 3 * H + y) * W;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        o[x] = prev[x];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            bool hit = disk_hit(sp, h, x, r2);
            if (a.extra_mask) hit |= a.extra_mask[(((size_t)b * 2 + h) * H + y) * W + x] != 0;
            o[(size_t)(h + 1) * H * W + x] = hit ? 1.f : 0.f;
        }
    }
}

int coord_features_launch(const CoordArgs& a, int B, float* out, cudaStream_t stream) {
    VPU_REQUIRE(a.n >= 1 && a.n <= 24, "coord features: n=%d outside [1, 24]", a.n);
    dim3 grid(a.H, B);
    VPU_CHECK_CUDA(launch_pdl(coord_features_kernel, dim3(grid), dim3(256), 0, stream, a, out));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

// grid (H, B), block W/2 threads: each thread converts a horizontal pixel pair of all 6 planes.
// A[m = (b, y/p, x/p), k = c*p*p + (y%p)*p + x%p];  planes c = 0..5: R, G, B (raw, normalisation is
// folded into the packed weights), prev mask, positive disks, negative disks, all as bf16 "hi" parts;
// planes 6..9: the bf16 "lo" residuals of R, G, B, prev mask ({0,1} disks are exact).  Row = 10*p*p.
__global__ void __launch_bounds__(256) patch_operand_kernel(const CoordArgs a, __nv_bfloat16* __restrict__ A, int p,
                                                            int lda) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ SmemPoints sp;
    const int y = blockIdx.x, b = blockIdx.y, H = a.H, W = a.W;
    const float r2 = a.radius * a.radius;
    load_points(sp, a.points + (size_t)b * 2 * a.n * 3, a.n, y, r2);
    __syncthreads();
    const int g = W / p, gy = y / p, ph = y % p, pp = p * p;
    for (int x = 2 * threadIdx.x; x < W; x += 2 * blockDim.x) {
        const int gx = x / p, pw = x % p;
        __nv_bfloat16* dst = A + ((size_t)b * g * (H / p) + (size_t)gy * g + gx) * lda + ph * p + pw;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            // split-bf16: v = hi + lo with |lo| <= 2^-9 |v|, so the patch-embed GEMM sees ~16 mantissa bits
            const float2 v = *reinterpret_cast<const float2*>(a.image4 + (((size_t)b * 4 + c) * H + y) * W + x);
            const uint32_t hi = pack_bf16(v.x, v.y);
            const float2 hf = unpack_bf16(hi);
            *reinterpret_cast<uint32_t*>(dst + c * pp) = hi;
            *reinterpret_cast<uint32_t*>(dst + (6 + c) * pp) = pack_bf16(v.x - hf.x, v.y - hf.y);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            bool h0 = disk_hit(sp, h, x, r2), h1 = disk_hit(sp, h, x + 1, r2);
            if (a.extra_mask) {
                const uint8_t* em = a.extra_mask + (((size_t)b * 2 + h) * H + y) * W + x;
                h0 |= em[0] != 0;
                h1 |= em[1] != 0;
            }
            *reinterpret_cast<uint32_t*>(dst + (4 + h) * pp) = pack_bf16(h0 ? 1.f : 0.f, h1 ? 1.f : 0.f);
        }
    }
}


// The same for 16-pixel patches (ViT-B / L): a WARP owns a patch, lane = (patch row, 8-pixel half), so that every store
// instruction of the warp writes the 512 contiguous bytes one plane of one patch occupies in its A row (k = plane * 256 +
// row * 16 + column) as 16-byte pieces.  The kernel above moves 4 bytes per store instruction to 4 different lines and ran at
// 2.3 TB/s (store-instruction-bound: 2 M warp stores for 257 MB); this one issues a quarter of them.
// grid (H / 16, B), block 256: the CTA owns one row of patches, warp w the patches w, w + 8, ...
__global__ void __launch_bounds__(256) patch_operand16_kernel(const CoordArgs a, __nv_bfloat16* __restrict__ A, int lda) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int P = 16, PP = P * P;
    __shared__ SmemPoints sp[P];
    const int gy = blockIdx.x, b = blockIdx.y, H = a.H, W = a.W, g = W / P;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float r2 = a.radius * a.radius;
    if (threadIdx.x < 2 * P) {       // the clicks that can reach each of the 16 image rows, per half (positive / negative)
        const int ry = threadIdx.x >> 1, h = threadIdx.x & 1, yy = gy * P + ry;
        const double* pts = a.points + (size_t)b * 2 * a.n * 3;
        int k = 0;
        for (int i = 0; i < a.n; ++i) {
            const double* p = pts + (size_t)(h * a.n + i) * 3;
            if (fmax(p[0], p[1]) >= 0.0) {
                const float dr = (float)((double)yy - p[0]);
                const float d2 = __fmul_rn(dr, dr);
                if (d2 <= r2) { sp[ry].pc[h * 24 + k] = p[1]; sp[ry].dr2[h * 24 + k] = d2; ++k; }
            }
        }
        sp[ry].cnt[h] = k;
    }
    __syncthreads();
    const int ph = lane >> 1, half = lane & 1, y = gy * P + ph;
    for (int gx = warp; gx < g; gx += 8) {
        const int x0 = gx * P + half * 8;
        __nv_bfloat16* dst = A + ((size_t)b * g * (H / P) + (size_t)gy * g + gx) * lda + ph * P + half * 8;
        float4 v[4][2];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float4* src = reinterpret_cast<const float4*>(a.image4 + (((size_t)b * 4 + c) * H + y) * W + x0);
            v[c][0] = src[0];
            v[c][1] = src[1];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {      // split-bf16: v = hi + lo with |lo| <= 2^-9 |v|
            const float f[8] = {v[c][0].x, v[c][0].y, v[c][0].z, v[c][0].w, v[c][1].x, v[c][1].y, v[c][1].z, v[c][1].w};
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                hi[q] = pack_bf16(f[2 * q], f[2 * q + 1]);
                const float2 hf = unpack_bf16(hi[q]);
                lo[q] = pack_bf16(f[2 * q] - hf.x, f[2 * q + 1] - hf.y);
            }
            *reinterpret_cast<uint4*>(dst + c * PP) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(dst + (6 + c) * PP) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t w[4];
            const uint8_t* em = a.extra_mask ? a.extra_mask + (((size_t)b * 2 + h) * H + y) * W + x0 : nullptr;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                bool h0 = disk_hit(sp[ph], h, x0 + 2 * q, r2), h1 = disk_hit(sp[ph], h, x0 + 2 * q + 1, r2);
                if (em) { h0 |= em[2 * q] != 0; h1 |= em[2 * q + 1] != 0; }
                w[q] = pack_bf16(h0 ? 1.f : 0.f, h1 ? 1.f : 0.f);
            }
            *reinterpret_cast<uint4*>(dst + (4 + h) * PP) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

int patch_operand_launch(const CoordArgs& a, int B, __nv_bfloat16* A, int patch, int lda, cudaStream_t stream) {
    VPU_REQUIRE(a.n >= 1 && a.n <= 24, "patch operand: n=%d outside [1, 24]", a.n);
    VPU_REQUIRE(patch % 2 == 0 && a.W % patch == 0 && a.H % patch == 0 && a.W % 2 == 0, "patch operand: bad geometry");
    if (patch == 16 && lda % 8 == 0 && a.W % 16 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(a.image4) & 15) == 0) {
        VPU_CHECK_CUDA(launch_pdl(patch_operand16_kernel, dim3(a.H / 16, B), dim3(256), 0, stream, a, A, lda));
        VPU_CHECK_CUDA(cudaGetLastError());
        count_launch();
        return 0;
    }
    dim3 grid(a.H, B);
    int threads = a.W / 2;
    threads = threads > 256 ? 256 : ((threads + 31) / 32) * 32;
    VPU_CHECK_CUDA(launch_pdl(patch_operand_kernel, dim3(grid), dim3(threads), 0, stream, a, A, patch, lda));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace vpu
