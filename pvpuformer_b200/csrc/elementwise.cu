// Normalisation, merge and resampling kernels of the path (HBM-bound; warp-level reductions,
// 16-byte vector accesses).  Each cites the reference op it implements.
#include "elementwise.cuh"

namespace vpu {

// ------------------------------------------------------------------------------------------
// LayerNorm over the last dim (reference nn.LayerNorm: eps 1e-6 in the ViT, models_vit.py:126;
// 1e-5 in the DMA blocks, transformer.py:412-422,268).  One warp per row, row kept in registers.
// Optional fused outputs: fp32 copy, bf16 copy, bf16(y + pe) (the with_pos_embed adds of
// transformer.py:439-458), and the per-row max of y (spatial gate of is_vpu_model.py:113-115).
// ------------------------------------------------------------------------------------------
template <int VEC>  // C = 128 * VEC
__global__ void __launch_bounds__(256) layernorm_kernel(const LnArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= a.rows) return;
    constexpr int C = 128 * VEC;
    const float4* in = reinterpret_cast<const float4*>(a.in + (size_t)row * C);
    float4 x[VEC];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        x[i] = in[lane + 32 * i];
        sum += x[i].x + x[i].y + x[i].z + x[i].w;
    }
    const float mean = warp_sum(sum) * (1.0f / C);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float d0 = x[i].x - mean, d1 = x[i].y - mean, d2 = x[i].z - mean, d3 = x[i].w - mean;
        var += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    const float rstd = rsqrtf(warp_sum(var) * (1.0f / C) + a.eps);
    float mx = -INFINITY;
    const int ldb = a.ld_bf16 ? a.ld_bf16 : C;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const int c4 = lane + 32 * i;
        const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma) + c4);
        const float4 b = __ldg(reinterpret_cast<const float4*>(a.beta) + c4);
        float4 y;
        y.x = (x[i].x - mean) * rstd * g.x + b.x;
        y.y = (x[i].y - mean) * rstd * g.y + b.y;
        y.z = (x[i].z - mean) * rstd * g.z + b.z;
        y.w = (x[i].w - mean) * rstd * g.w + b.w;
        mx = fmaxf(mx, fmaxf(fmaxf(y.x, y.y), fmaxf(y.z, y.w)));
        if (a.out_f32) reinterpret_cast<float4*>(a.out_f32 + (size_t)row * C)[c4] = y;
        if (a.out_bf16) {
            uint2 o = make_uint2(pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
            reinterpret_cast<uint2*>(a.out_bf16 + (size_t)row * ldb)[c4] = o;
        }
        if (a.out_pe_bf16) {
            const float4 p = reinterpret_cast<const float4*>(a.pe + (size_t)row * C)[c4];
            uint2 o = make_uint2(pack_bf16(y.x + p.x, y.y + p.y), pack_bf16(y.z + p.z, y.w + p.w));
            reinterpret_cast<uint2*>(a.out_pe_bf16 + (size_t)row * ldb)[c4] = o;
        }
    }
    if (a.rowmax) {
        mx = warp_max(mx);
        if (lane == 0) a.rowmax[row] = mx;
    }
}

int layernorm_launch(const LnArgs& a, int C, cudaStream_t stream) {
    VPU_REQUIRE(a.rows > 0, "layernorm: no rows");
    const int grid = (a.rows + 7) / 8;
    switch (C) {
        case 768: VPU_CHECK_CUDA(launch_pdl(layernorm_kernel<6>, dim3(grid), dim3(256), 0, stream, a)); break;
        case 1024: VPU_CHECK_CUDA(launch_pdl(layernorm_kernel<8>, dim3(grid), dim3(256), 0, stream, a)); break;
        case 1280: VPU_CHECK_CUDA(launch_pdl(layernorm_kernel<10>, dim3(grid), dim3(256), 0, stream, a)); break;
        default: VPU_REQUIRE(false, "layernorm: unsupported width %d (768/1024/1280)", C);
    }
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

// out_bf16 = bf16(a (+ b)); n multiple of 4
__global__ void __launch_bounds__(256) cast_add_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                       __nv_bfloat16* __restrict__ out, size_t n4) {
    pdl_launch_dependents();
    pdl_wait();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 x = reinterpret_cast<const float4*>(a)[i];
        if (b) {
            const float4 y = reinterpret_cast<const float4*>(b)[i];
            x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
        }
        reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w));
    }
}
int cast_add_launch(const float* a, const float* b, __nv_bfloat16* out, size_t n, cudaStream_t stream) {
    VPU_REQUIRE(n % 4 == 0, "cast: element count must be a multiple of 4");
    const size_t n4 = n / 4;
    int grid = (int)((n4 + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    VPU_CHECK_CUDA(launch_pdl(cast_add_kernel, dim3(grid), dim3(256), 0, stream, a, b, out, n4));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------------------
// GroupNorm(1, C) (+ optional GELU) on NHWC bf16, in place (reference is_vpu_model.py:56-86:
// nn.GroupNorm(1, C) eps 1e-5 => statistics over all C*H*W values of one sample).
// stats: grid (chunks, B) fp32 partial sums; finalize in double; apply: elementwise.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_stats_kernel(const __nv_bfloat16* __restrict__ x, size_t per_sample,
                                                       float2* __restrict__ partial) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y, chunks = gridDim.x;
    const uint4* p = reinterpret_cast<const uint4*>(x + (size_t)b * per_sample);
    const size_t n8 = per_sample / 8;
    float s = 0.f, ss = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)chunks * blockDim.x) {
        const uint4 v = p[i];
        const float2 a = unpack_bf16(v.x), b2 = unpack_bf16(v.y), c = unpack_bf16(v.z), d = unpack_bf16(v.w);
        s += a.x + a.y + b2.x + b2.y + c.x + c.y + d.x + d.y;
        ss += a.x * a.x + a.y * a.y + b2.x * b2.x + b2.y * b2.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
    }
    __shared__ float sh[2][8];
    s = warp_sum(s);
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float S = 0.f, SS = 0.f;
        for (int w = 0; w < 8; ++w) { S += sh[0][w]; SS += sh[1][w]; }
        partial[(size_t)b * chunks + blockIdx.x] = make_float2(S, SS);
    }
}
__global__ void gn_finalize_kernel(const float2* __restrict__ partial, int chunks, double count, float eps,
                                   float2* __restrict__ mean_rstd) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.x;
    double s = 0.0, ss = 0.0;
    for (int i = threadIdx.x; i < chunks; i += 32) { s += partial[(size_t)b * chunks + i].x; ss += partial[(size_t)b * chunks + i].y; }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
    if (threadIdx.x == 0) {
        const double mean = s / count;
        double var = ss / count - mean * mean;
        if (var < 0) var = 0;
        mean_rstd[b] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
    }
}
// same fit of the erf form as gelu_fast (common.cuh), evaluated through ONE MUFU (tanh.approx, relative error 2^-11, i.e. a quarter
// of the half-ulp of the bf16 this kernel stores) instead of ex2 + rcp: the apply pass is bound by its instruction stream, not by HBM
// (without the GELU it runs at 4.95 TB/s, with the two-MUFU form at 3.5, with this one at 3.7: 330 -> 291 us over the five passes)
__device__ __forceinline__ float gelu_tanh1(float x) {
    const float x2 = fminf(x * x, 64.0f);
    float t = fmaf(-3.515167e-4f, x2, 3.700565e-2f);
    t = fmaf(t, x2, 7.975079e-1f);
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(x * t));
    const float h = 0.5f * x;
    return fmaf(h, th, h);
}

// apply: every thread owns one channel octet (gamma / beta live in registers) and strides over pixels
__global__ void __launch_bounds__(256) gn_apply_kernel(__nv_bfloat16* __restrict__ x, size_t per_sample, int C,
                                                       const float2* __restrict__ mean_rstd, const long long* __restrict__ sums,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       int gelu) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y, C8 = C / 8, C8z = C8 / gridDim.z;      // wide layers split their channels over grid.z
    const int co = blockIdx.z * C8z + threadIdx.x % C8z, prow = threadIdx.x / C8z, rows_per_block = blockDim.x / C8z;
    float2 mr;
    if (sums) {      // (sum, sum of squares) accumulated by the producing GEMM's epilogue
        const double inv_n = 1.0 / (double)per_sample, mean = (double)sums[2 * b] * (1.0 / (double)GN_SUM_SCALE) * inv_n;
        double var = (double)sums[2 * b + 1] * (1.0 / (double)GN_SQ_SCALE) * inv_n - mean * mean;
        if (var < 0.0) var = 0.0;
        mr = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-5)));
    } else {
        mr = mean_rstd[b];
    }
    float sc[8], sh[8];
    {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + co * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + co * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + co * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + co * 8 + 4));
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) { sc[k] = mr.y * g[k]; sh[k] = bb[k] - mr.x * mr.y * g[k]; }   // y = x*sc + sh
    }
    uint4* p = reinterpret_cast<uint4*>(x + (size_t)b * per_sample);
    const size_t npix = per_sample / C, step = (size_t)gridDim.x * rows_per_block;
    auto apply = [&](uint4 v) {
        float f[8];
        float2 t;
        t = unpack_bf16(v.x); f[0] = t.x; f[1] = t.y;
        t = unpack_bf16(v.y); f[2] = t.x; f[3] = t.y;
        t = unpack_bf16(v.z); f[4] = t.x; f[5] = t.y;
        t = unpack_bf16(v.w); f[6] = t.x; f[7] = t.y;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float y = fmaf(f[k], sc[k], sh[k]);
            f[k] = gelu ? gelu_tanh1(y) : y;
        }
        return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
    };
    // four pixels per iteration, and the loads of the next four are issued before the current four are evaluated and stored: without
    // that a block's warps load, evaluate (MUFU) and store in lockstep and the MUFU time adds to the memory time instead of hiding
    // under it
    size_t pix = (size_t)blockIdx.x * rows_per_block + prow;
    uint4 v[4], nv[4];
    bool have = pix + 3 * step < npix;
    if (have) {
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = p[(pix + u * step) * C8 + co];
    }
    while (have) {
        const size_t npx = pix + 4 * step;
        const bool next = npx + 3 * step < npix;
        if (next) {
#pragma unroll
            for (int u = 0; u < 4; ++u) nv[u] = p[(npx + u * step) * C8 + co];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) p[(pix + u * step) * C8 + co] = apply(v[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = nv[u];
        pix = npx;
        have = next;
    }
    for (; pix < npix; pix += step) p[pix * C8 + co] = apply(p[pix * C8 + co]);
}
int groupnorm_launch(__nv_bfloat16* x, int B, size_t per_sample, int C, const float* gamma, const float* beta,
                     int gelu, float2* partial, float2* mean_rstd, cudaStream_t stream) {
    VPU_REQUIRE(per_sample % 8 == 0 && C % 8 == 0, "groupnorm: sizes must be multiples of 8");
    int chunks = (int)((per_sample / 8 + 255) / 256);
    if (chunks > GN_MAX_CHUNKS) chunks = GN_MAX_CHUNKS;
    VPU_CHECK_CUDA(launch_pdl(gn_stats_kernel, dim3(chunks, B), dim3(256), 0, stream, x, per_sample, partial));
    VPU_CHECK_CUDA(launch_pdl(gn_finalize_kernel, dim3(B), dim3(32), 0, stream, partial, chunks, (double)per_sample, 1e-5f, mean_rstd));
    VPU_REQUIRE(per_sample % C == 0, "groupnorm: C must divide the sample size");
    const int C8 = C / 8;
    int zsplit = 1;
    while (C8 / zsplit > 256 || C8 % zsplit) ++zsplit;
    const int C8z = C8 / zsplit, threads = C8z * (256 / C8z);
    const size_t npix = per_sample / C;
    // 16 pixels per thread (four pipelined iterations; 4 / 8 / 16 / 32 measured 306 / 263 / 249 / 255 us over the neck's five passes at
    // batch 64) unless that leaves SMs without a block (small batches): then 4
    const int rpb = threads / C8z;
    int blocks = (int)((npix + 16 * rpb - 1) / (16 * rpb));
    if ((long long)blocks * B * zsplit < 4 * 148) blocks = (int)((npix + 4 * rpb - 1) / (4 * rpb));
    if (blocks > 148 * 8) blocks = 148 * 8;
    VPU_CHECK_CUDA(launch_pdl(gn_apply_kernel, dim3(blocks, B, zsplit), dim3(threads), 0, stream, x, per_sample, C, mean_rstd, nullptr, gamma, beta, gelu));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch(3);
    return 0;
}

// apply only: the statistics were accumulated by the GEMM epilogue that produced x (Epi::gn_out)
int groupnorm_apply_launch(__nv_bfloat16* x, int B, size_t per_sample, int C, const float* gamma, const float* beta, int gelu,
                           const long long* sums, cudaStream_t stream) {
    VPU_REQUIRE(per_sample % 8 == 0 && C % 8 == 0 && per_sample % C == 0, "groupnorm: sizes must be multiples of 8, C must divide the sample");
    const int C8 = C / 8;
    int zsplit = 1;
    while (C8 / zsplit > 256 || C8 % zsplit) ++zsplit;
    const int C8z = C8 / zsplit, threads = C8z * (256 / C8z);
    const size_t npix = per_sample / C;
    // 16 pixels per thread (four pipelined iterations; 4 / 8 / 16 / 32 measured 306 / 263 / 249 / 255 us over the neck's five passes at
    // batch 64) unless that leaves SMs without a block (small batches): then 4
    const int rpb = threads / C8z;
    int blocks = (int)((npix + 16 * rpb - 1) / (16 * rpb));
    if ((long long)blocks * B * zsplit < 4 * 148) blocks = (int)((npix + 4 * rpb - 1) / (4 * rpb));
    if (blocks > 148 * 8) blocks = 148 * 8;
    VPU_CHECK_CUDA(launch_pdl(gn_apply_kernel, dim3(blocks, B, zsplit), dim3(threads), 0, stream, x, per_sample, C, nullptr, sums, gamma, beta, gelu));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------------------
// q_out = q + q1 + q2 + q_final and the three channel gates sigmoid(max over the 48 prompt
// tokens) (reference is_vpu_model.py:104-108).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) qout_gate_kernel(const float* __restrict__ q0, const float* __restrict__ q1,
                                                        const float* __restrict__ q2, const float* __restrict__ q3,
                                                        int T, int C, float* __restrict__ qout,
                                                        __nv_bfloat16* __restrict__ qout_bf16, float* __restrict__ cg) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x, B = gridDim.y;
    if (c >= C) return;
    float m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
    for (int t = 0; t < T; ++t) {
        const size_t i = ((size_t)b * T + t) * C + c;
        const float a1 = q1[i], a2 = q2[i], a3 = q3[i];
        m1 = fmaxf(m1, a1); m2 = fmaxf(m2, a2); m3 = fmaxf(m3, a3);
        const float s = q0[i] + a1 + a2 + a3;
        qout[i] = s;
        qout_bf16[i] = __float2bfloat16(s);
    }
    cg[((size_t)0 * B + b) * C + c] = 1.0f / (1.0f + expf(-m1));
    cg[((size_t)1 * B + b) * C + c] = 1.0f / (1.0f + expf(-m2));
    cg[((size_t)2 * B + b) * C + c] = 1.0f / (1.0f + expf(-m3));
}
int qout_gate_launch(const float* q0, const float* q1, const float* q2, const float* q3, int B, int T, int C, float* qout,
                     __nv_bfloat16* qout_bf16, float* cg, cudaStream_t stream) {
    VPU_CHECK_CUDA(launch_pdl(qout_gate_kernel, dim3((C + 127) / 128, B), dim3(128), 0, stream, q0, q1, q2, q3, T, C, qout, qout_bf16, cg));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------------------
// DMA merge (reference is_vpu_model.py:109-126): x_l = x + x*cg_l[b,c] + x*sg_l[b,token],
// sg = sigmoid(rowmax of keys_l).  One pass writes bf16 x0 (un-merged, for down_4), x2, x3 and
// x4 in the space-to-depth layout the 2x2/stride-2 conv of down_32 consumes as a plain GEMM.
// ------------------------------------------------------------------------------------------
// Block = one float4 column per thread (C / 4 threads), grid-stride over groups of MERGE_ROWS token rows: the per-row quantities
// -- image index, the three spatial gates sigmoid(rowmax), the space-to-depth row -- are worked out by 3 * MERGE_ROWS threads and
// shared; every thread then has MERGE_ROWS independent 16-byte loads in flight.  The element-indexed version before it paid a 64-bit
// and four 32-bit integer divisions plus three expf + three divisions per float4 and ran at 3.1 TB/s.
constexpr int MERGE_ROWS = 4;
__global__ void __launch_bounds__(320) merge_kernel(const MergeArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sg_s[MERGE_ROWS][3];
    __shared__ int b_s[MERGE_ROWS];
    __shared__ long long x4_s[MERGE_ROWS];
    const int C = a.C, c = threadIdx.x * 4;
    const int g = a.grid, gh = g / 2;
    for (int m0 = blockIdx.x * MERGE_ROWS; m0 < a.M; m0 += gridDim.x * MERGE_ROWS) {
        __syncthreads();                 // the previous group's shared values have been consumed
        if (threadIdx.x < 3 * MERGE_ROWS) {
            const int r = threadIdx.x / 3, l = threadIdx.x % 3, m = m0 + r;
            if (m < a.M) {
                float mx = a.rowmax[(size_t)l * a.rowmax_parts * a.M + m];
                for (int q = 1; q < a.rowmax_parts; ++q) mx = fmaxf(mx, a.rowmax[((size_t)l * a.rowmax_parts + q) * a.M + m]);
                sg_s[r][l] = 1.0f / (1.0f + expf(-mx));
                if (l == 0) {
                    const int b = m / a.N, tok = m % a.N, i = tok / g, j = tok % g;
                    b_s[r] = b;
                    x4_s[r] = ((long long)b * gh * gh + (long long)(i / 2) * gh + j / 2) * 4 * C + (long long)((i & 1) * 2 + (j & 1)) * C;
                }
            }
        }
        __syncthreads();
        float4 x[MERGE_ROWS];
#pragma unroll
        for (int r = 0; r < MERGE_ROWS; ++r)
            if (m0 + r < a.M) x[r] = *reinterpret_cast<const float4*>(a.x + (size_t)(m0 + r) * C + c);
#pragma unroll
        for (int r = 0; r < MERGE_ROWS; ++r) {
            const int m = m0 + r;
            if (m >= a.M) break;
            const int b = b_s[r];
            const float4 xv = x[r];
            auto gate = [&](int l) {
                const float4 cgv = *reinterpret_cast<const float4*>(a.cg + ((size_t)l * a.B + b) * C + c);
                const float sg = sg_s[r][l];
                return make_float4(xv.x + xv.x * cgv.x + xv.x * sg, xv.y + xv.y * cgv.y + xv.y * sg, xv.z + xv.z * cgv.z + xv.z * sg,
                                   xv.w + xv.w * cgv.w + xv.w * sg);
            };
            auto st = [&](__nv_bfloat16* dst, size_t off, const float4& v) {
                *reinterpret_cast<uint2*>(dst + off) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
            };
            if (a.x0) st(a.x0, (size_t)m * C + c, xv);
            st(a.x2, (size_t)m * C + c, gate(0));
            st(a.x3, (size_t)m * C + c, gate(1));
            st(a.x4_s2d, (size_t)x4_s[r] + c, gate(2));
        }
    }
}
int merge_launch(const MergeArgs& a, cudaStream_t stream) {
    VPU_REQUIRE(a.C % 4 == 0 && a.grid % 2 == 0 && a.C / 4 <= 320 && a.C / 4 >= 3 * MERGE_ROWS, "merge: bad geometry");
    int grid = (a.M + MERGE_ROWS - 1) / MERGE_ROWS;
    if (grid > 148 * 16) grid = 148 * 16;
    VPU_CHECK_CUDA(launch_pdl(merge_kernel, dim3(grid), dim3(a.C / 4), 0, stream, a));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------------------
// Head combine.  Reference swin_transformer.py:727-737: four (1x1 conv + ReLU) maps are resized
// (bilinear, align_corners=False) to the 1/4 scale, concatenated and passed through the fusion
// 1x1 conv + ReLU.  A 1x1 conv mixes channels per pixel and bilinear resizing mixes pixels per
// channel, so they commute: fusion(cat_i resize(h_i)) = sum_i resize(W_i h_i).  The GEMMs produce
// y_i = W_i h_i at native resolution; this kernel forms relu(b + y_0 + sum_i resize(y_i)) without
// ever materialising the 1024-channel concat, and emits 1/max(||f||, 1e-12) per pixel for the
// P2CL cosine logits (F.normalize, swin_transformer.py:751).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void src_index(int dst, int in, int out, int& i0, int& i1, float& lam) {
    float s = ((float)dst + 0.5f) * ((float)in / (float)out) - 0.5f;   // align_corners=False
    if (s < 0.f) s = 0.f;
    i0 = (int)s;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    lam = s - (float)i0;
}
__device__ __forceinline__ void acc8(float (&f)[8], const __nv_bfloat16* p, float w) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    float2 t;
    t = unpack_bf16(v.x); f[0] += w * t.x; f[1] += w * t.y;
    t = unpack_bf16(v.y); f[2] += w * t.x; f[3] += w * t.y;
    t = unpack_bf16(v.z); f[4] += w * t.x; f[5] += w * t.y;
    t = unpack_bf16(v.w); f[6] += w * t.x; f[7] += w * t.y;
}
// 256 channels = 8 per lane, one warp per output pixel at a time.  A block owns an 8 x 8 patch of output pixels (warp w =
// row w of the patch, 8 pixels in sequence), so the 2 x 2 bilinear taps of the three coarser levels -- 12 of the 13
// 512-byte loads per pixel -- are shared through L1 by the whole patch (6x6 / 4x4 / 3x3 source pixels per 64 outputs).
// With one pixel row of 8 per block (round 1c) every tap came from L2: 5.3 GB of L2 traffic for 0.96 GB of HBM bytes.
__global__ void __launch_bounds__(256) head_combine_kernel(const HeadCombineArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int R = a.res[0], T = R / 8;
    const int b = blockIdx.x / (T * T), tyx = blockIdx.x % (T * T), y = (tyx / T) * 8 + warp, xb = (tyx % T) * 8;
    const int c = lane * 8;
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.bias + c)), b1 = __ldg(reinterpret_cast<const float4*>(a.bias + c + 4));
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.wseg + c)), w1 = __ldg(reinterpret_cast<const float4*>(a.wseg + c + 4));
    const float ws[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    int y0[4], y1[4];
    float ly[4];
#pragma unroll
    for (int l = 1; l < 4; ++l) src_index(y, a.res[l], R, y0[l], y1[l], ly[l]);
#pragma unroll 2
    for (int i = 0; i < 8; ++i) {
        const int x = xb + i;
        const size_t pix = ((size_t)b * R + y) * R + x;
        float f[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        acc8(f, a.y[0] + pix * 256 + c, 1.0f);
#pragma unroll
        for (int l = 1; l < 4; ++l) {
            const int r = a.res[l];
            int x0, x1;
            float lx;
            src_index(x, r, R, x0, x1, lx);
            const __nv_bfloat16* base = a.y[l] + (size_t)b * r * r * 256 + c;
            acc8(f, base + ((size_t)y0[l] * r + x0) * 256, (1.f - ly[l]) * (1.f - lx));
            acc8(f, base + ((size_t)y0[l] * r + x1) * 256, (1.f - ly[l]) * lx);
            acc8(f, base + ((size_t)y1[l] * r + x0) * 256, ly[l] * (1.f - lx));
            acc8(f, base + ((size_t)y1[l] * r + x1) * 256, ly[l] * lx);
        }
        float ss = 0.f, seg = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            f[k] = fmaxf(f[k], 0.f);
            ss += f[k] * f[k];
            seg += f[k] * ws[k];                     // conv_seg (decode_head.py:210-215) in fp32, before the bf16 store
        }
        ss = warp_sum(ss);
        seg = warp_sum(seg);
        *reinterpret_cast<uint4*>(a.out + pix * 256 + c) =
            make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
        if (lane == 0) {
            a.seg_out[pix] = seg + a.seg_bias;
            a.rnorm[pix] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
        }
    }
}
int head_combine_launch(const HeadCombineArgs& a, cudaStream_t stream) {
    VPU_REQUIRE(a.res[0] % 8 == 0, "head combine: the 1/4-scale resolution (%d) must be a multiple of 8", a.res[0]);
    const int T = a.res[0] / 8;
    VPU_CHECK_CUDA(launch_pdl(head_combine_kernel, dim3((unsigned)(a.B * T * T)), dim3(256), 0, stream, a));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

// P2CL queries (reference swin_transformer.py:745,750): L2-normalise the FFN output rows and lay
// out the per-sample B operand [64, 256] of the head-final GEMM: rows 0..nq-1 = normalised
// queries, row nq = conv_seg weight, remaining rows zero.
__global__ void __launch_bounds__(256) head_queries_kernel(const float* __restrict__ qe, const float* __restrict__ wseg,
                                                           int nq, __nv_bfloat16* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.x, r = blockIdx.y, c = threadIdx.x;  // 256 channels
    __shared__ float sh[8];
    float v = 0.f, scale = 1.f;
    if (r < nq) {
        v = qe[((size_t)b * nq + r) * 256 + c];
        float ss = warp_sum(v * v);
        if ((c & 31) == 0) sh[c >> 5] = ss;
        __syncthreads();
        float tot = 0.f;
        for (int w = 0; w < 8; ++w) tot += sh[w];
        scale = 1.0f / fmaxf(sqrtf(tot), 1e-12f);
    } else if (r == nq) {
        v = wseg[c];
    }
    out[((size_t)b * 64 + r) * 256 + c] = __float2bfloat16(v * scale);
}
int head_queries_launch(const float* qe, const float* wseg, int B, int nq, __nv_bfloat16* out, cudaStream_t stream) {
    VPU_REQUIRE(nq < 64, "head queries: nq must be < 64");
    VPU_CHECK_CUDA(launch_pdl(head_queries_kernel, dim3(B, 64), dim3(256), 0, stream, qe, wseg, nq, out));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------------------
// Final bilinear upsampling, align_corners=True (reference is_vpu_model.py:431-436).
// in [planes, h, w] fp32 -> out [planes, H, W] fp32; each thread writes 4 consecutive x.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample_ac_kernel(const float* __restrict__ in, float* __restrict__ out, int h, int w,
                                                          int H, int W, size_t planes) {
    pdl_launch_dependents();
    pdl_wait();
    const int W4 = W / 4;
    const size_t total = planes * H * W4;
    const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
    const float sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int x4 = (int)(idx % W4);
        const int y = (int)((idx / W4) % H);
        const size_t pl = idx / ((size_t)W4 * H);
        const float fy = sy * (float)y;
        const int y0 = (int)fy, y1 = y0 + (y0 < h - 1 ? 1 : 0);
        const float ly = fy - (float)y0;
        const float* r0 = in + (pl * h + y0) * w;
        const float* r1 = in + (pl * h + y1) * w;
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float fx = sx * (float)(x4 * 4 + k);
            const int x0 = (int)fx, x1 = x0 + (x0 < w - 1 ? 1 : 0);
            const float lx = fx - (float)x0;
            o[k] = (1.f - ly) * ((1.f - lx) * __ldg(r0 + x0) + lx * __ldg(r0 + x1)) +
                   ly * ((1.f - lx) * __ldg(r1 + x0) + lx * __ldg(r1 + x1));
        }
        reinterpret_cast<float4*>(out)[idx] = make_float4(o[0], o[1], o[2], o[3]);
    }
}
// 4 x 4 output block per thread: for the x4 scale of this path (112 -> 448, step 111/447 < 1/3) the 16 outputs read a
// 3 x 3 source patch, so loads drop from 64 to 9 per block and the horizontal interpolation is shared by 4 rows.
__global__ void __launch_bounds__(256) upsample_ac4_kernel(const float* __restrict__ in, float* __restrict__ out, int h, int w,
                                                           int H, int W, size_t planes) {
    pdl_launch_dependents();
    pdl_wait();
    const int W4 = W / 4, H4 = H / 4;
    const size_t total = planes * H4 * W4;
    const float sy = (float)(h - 1) / (float)(H - 1), sx = (float)(w - 1) / (float)(W - 1);
    // 32-bit index arithmetic (total < 2^31 is checked by the launcher): three 64-bit divisions per 4 x 4 block were a third of the
    // kernel's instructions
    const unsigned total32 = (unsigned)total, stride = gridDim.x * blockDim.x, per_plane = (unsigned)(W4 * H4);
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total32; idx += stride) {
        const unsigned plu = idx / per_plane, rem = idx - plu * per_plane;
        const int Y = (int)(rem / (unsigned)W4), x4 = (int)(rem - (unsigned)Y * (unsigned)W4);
        const size_t pl = plu;
        int xo[4], yo[4];
        float lx[4], ly[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float fx = sx * (float)(x4 * 4 + k), fy = sy * (float)(Y * 4 + k);
            xo[k] = (int)fx; lx[k] = fx - (float)xo[k];
            yo[k] = (int)fy; ly[k] = fy - (float)yo[k];
        }
        const int xb = xo[0], yb = yo[0];
        float src[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float* row = in + (pl * h + min(yb + r, h - 1)) * w;
#pragma unroll
            for (int c = 0; c < 3; ++c) src[r][c] = __ldg(row + min(xb + c, w - 1));
        }
        float hz[3][4];     // horizontally interpolated source rows
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool s1 = xo[k] != xb;        // the 4 outputs span at most two source columns
                const float a = s1 ? src[r][1] : src[r][0], b = s1 ? src[r][2] : src[r][1];
                hz[r][k] = (1.f - lx[k]) * a + lx[k] * b;
            }
        float* o = out + (pl * H + (size_t)Y * 4) * W + (size_t)x4 * 4;
#pragma unroll
        for (int yy = 0; yy < 4; ++yy) {
            const bool s1 = yo[yy] != yb;
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float a = s1 ? hz[1][k] : hz[0][k], b = s1 ? hz[2][k] : hz[1][k];
                v[k] = (1.f - ly[yy]) * a + ly[yy] * b;
            }
            __stcs(reinterpret_cast<float4*>(o + (size_t)yy * W), make_float4(v[0], v[1], v[2], v[3]));     // written once, read by the host
        }
    }
}

int upsample_ac_launch(const float* in, float* out, int h, int w, int H, int W, size_t planes, cudaStream_t stream) {
    VPU_REQUIRE(W % 4 == 0, "upsample: output width must be a multiple of 4");
    if (H % 4 == 0 && H > 1 && W > 1 && 3 * (h - 1) < (H - 1) && 3 * (w - 1) < (W - 1) &&
        planes * (H / 4) * (W / 4) < (size_t)1 << 31) {   // 4 outputs within one source step
        const size_t total = planes * (H / 4) * (W / 4);
        size_t grid = (total + 255) / 256;
        if (grid > 148 * 32) grid = 148 * 32;
        VPU_CHECK_CUDA(launch_pdl(upsample_ac4_kernel, dim3((unsigned)grid), dim3(256), 0, stream, in, out, h, w, H, W, planes));
        VPU_CHECK_CUDA(cudaGetLastError());
        count_launch();
        return 0;
    }
    const size_t total = planes * H * (W / 4);
    size_t grid = (total + 255) / 256;
    if (grid > 148 * 32) grid = 148 * 32;
    VPU_CHECK_CUDA(launch_pdl(upsample_ac_kernel, dim3((unsigned)grid), dim3(256), 0, stream, in, out, h, w, H, W, planes));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace vpu
