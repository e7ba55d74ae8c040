// Dispatch of every attention shape of the path to its tcgen05 / TMEM / TMA kernel (attention_tc.cu: ViT window and global
// self-attention, reference models_vit.py:43-56 + :225-255; attention_dma.cu: the three DMA shapes, transformer.py:499-521).
// Q/K/V are read in place from the projection outputs (row stride + column offset), so no head split / window partition
// copies exist.
//
// -DVPU_DEBUG builds also compile the round-1 kernel below -- flash-style online softmax on mma.sync m16n8k16 with cp.async
// double-buffered K/V tiles -- as the A/B baseline (VPU_ATTN_TC=0); the shipped library has no mma.sync attention and
// rejects shapes none of the tcgen05 kernels covers.
#include <cstdlib>

#include "attention.cuh"

namespace vpu {

#ifdef VPU_DEBUG
constexpr int ATT_KROW_MAX = 256;   // keys of one window (16 x 16 for ViT-H)

__device__ __forceinline__ int map_row(const RowMap& rm, int prob, int s) {
    if (rm.mode == 0) return prob * rm.per_prob + s;
    const int nw = rm.grid / rm.win, nw2 = nw * nw;
    const int b = prob / nw2, w = prob % nw2;
    const int wi = w / nw, wj = w % nw;
    const int i = s / rm.win, j = s % rm.win;
    return b * rm.tokens + (wi * rm.win + i) * rm.grid + wj * rm.win + j;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    const uint32_t d = smem_u32(smem_dst);
    const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// one MUFU.EX2: the arguments are <= 0 (or -inf for masked keys), results in [0, 1]; exp2f() costs five instructions per element
// for its denormal range handling, which this softmax never needs
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// BM query rows per CTA (one warp per 16), BN keys per tile: 64 x 64 in general; the DMA shapes have 48 queries and / or 48 keys
// (transformer.py:499-521 with 24 + 24 prompt tokens), where a 64-wide tile spends a quarter of its MMAs and exponentials on
// padding, so those launches use 48-row / 48-key tiles (3 warps / 3 k-steps)
template <int D, int BM, int BN>
__global__ void __launch_bounds__(BM * 2) attention_kernel(const AttnArgs a) {
    constexpr int NT = BM * 2;              // threads: one warp per 16 query rows
    pdl_launch_dependents();
    pdl_wait();
    constexpr int PITCH = D + 8;            // +16 B per row: conflict-free ldmatrix
    constexpr int CHUNKS = D / 8;           // 16-byte chunks per row
    extern __shared__ __align__(16) uint8_t att_smem[];
    __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(att_smem);
    __nv_bfloat16* Ks = Qs + BM * PITCH;            // [2][BN][PITCH]
    __nv_bfloat16* Vs = Ks + 2 * BN * PITCH;        // [2][BN][PITCH]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int q0 = blockIdx.x * BM, h = blockIdx.y, prob = blockIdx.z;
    const __nv_bfloat16* qbase = a.q + a.qoff + h * D;
    const __nv_bfloat16* kbase = a.k + a.koff + h * D;
    const __nv_bfloat16* vbase = a.v + a.voff + h * D;

    // ---- async loads -------------------------------------------------------------------
    for (int idx = tid; idx < BM * CHUNKS; idx += NT) {
        const int r = idx / CHUNKS, c = idx % CHUNKS;
        const int s = q0 + r;
        const bool ok = s < a.Sq;
        const size_t row = ok ? (size_t)map_row(a.qmap, prob, s) : 0;
        cp_async16(Qs + r * PITCH + c * 8, qbase + row * a.ldq + c * 8, ok);
    }
    // window regrouping of the keys (map_row: six integer divisions) worked out once per CTA instead of once per 16-byte
    // chunk of every tile: the divisions were ~40 % of the kernel's instructions on the ViT-H windows (ncu, round 1h)
    __shared__ int krow[ATT_KROW_MAX];
    const bool ktab = a.kmap.mode == 1 && a.Sk <= ATT_KROW_MAX;
    if (ktab) {
        for (int s = tid; s < a.Sk; s += NT) krow[s] = map_row(a.kmap, prob, s);
        __syncthreads();
    }
    auto load_kv = [&](int tile, int buf) {
        for (int idx = tid; idx < BN * CHUNKS; idx += NT) {
            const int r = idx / CHUNKS, c = idx % CHUNKS;
            const int s = tile * BN + r;
            const bool ok = s < a.Sk;
            const size_t row = ok ? (size_t)(ktab ? krow[s] : map_row(a.kmap, prob, s)) : 0;
            cp_async16(Ks + (buf * BN + r) * PITCH + c * 8, kbase + row * a.ldk + c * 8, ok);
            cp_async16(Vs + (buf * BN + r) * PITCH + c * 8, vbase + row * a.ldv + c * 8, ok);
        }
    };
    const int ntiles = (a.Sk + BN - 1) / BN;
    load_kv(0, 0);
    cp_async_commit();

    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
    const float sl2 = a.scale_log2;

    for (int tile = 0; tile < ntiles; ++tile) {
        const int buf = tile & 1;
        if (tile + 1 < ntiles) {
            load_kv(tile + 1, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        // ---- S = Q K^T (16 x 64 per warp) ------------------------------------------------
        float s[BN / 8][4];
#pragma unroll
        for (int i = 0; i < BN / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
        const __nv_bfloat16* Kb = Ks + buf * BN * PITCH;
        const __nv_bfloat16* Vb = Vs + buf * BN * PITCH;
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
            uint32_t a0, a1, a2, a3;
            ldsm_x4(smem_u32(Qs + (warp * 16 + (lane & 15)) * PITCH + ks * 16 + (lane >> 4) * 8), a0, a1, a2, a3);
#pragma unroll
            for (int np = 0; np < BN / 16; ++np) {
                uint32_t b0, b1, b2, b3;
                // matrices: (keys 0-7, d 0-7), (keys 0-7, d 8-15), (keys 8-15, d 0-7), (keys 8-15, d 8-15)
                ldsm_x4(smem_u32(Kb + (np * 16 + (lane >> 4) * 8 + (lane & 7)) * PITCH + ks * 16 + ((lane >> 3) & 1) * 8),
                        b0, b1, b2, b3);
                mma_bf16(s[2 * np], a0, a1, a2, a3, b0, b1);
                mma_bf16(s[2 * np + 1], a0, a1, a2, a3, b2, b3);
            }
        }
        // ---- mask + online softmax (rows g and g+8 of this warp's 16) ------------------------
        // raw scores: the positive scale commutes with the maximum and is applied inside the exponent's FMA
        const int kbase_idx = tile * BN;
        if (kbase_idx + BN > a.Sk) {                 // ragged last tile only
#pragma unroll
            for (int nt = 0; nt < BN / 8; ++nt) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (kbase_idx + nt * 8 + 2 * t + (e & 1) >= a.Sk) s[nt][e] = -INFINITY;
            }
        }
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < BN / 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
        }
        float corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float mnew = fmaxf(mrow[r], mx[r] * sl2);   // finite: every tile has >= 1 valid key
            corr[r] = ex2_fast(mrow[r] - mnew);
            mrow[r] = mnew;
            lrow[r] *= corr[r];
        }
        float psum[2] = {0.f, 0.f};
#pragma unroll
        for (int nt = 0; nt < BN / 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float p = ex2_fast(fmaf(s[nt][e], sl2, -mrow[e >> 1]));
                s[nt][e] = p;
                psum[e >> 1] += p;
            }
        }
        lrow[0] += psum[0];
        lrow[1] += psum[1];
#pragma unroll
        for (int i = 0; i < D / 8; ++i) {
            o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1];
        }
        // ---- O += P V ---------------------------------------------------------------------
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
            const uint32_t a0 = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
            const uint32_t a1 = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
            const uint32_t a2 = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
            const uint32_t a3 = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
            for (int np = 0; np < D / 16; ++np) {
                uint32_t b0, b1, b2, b3;
                // .trans matrices: (keys 0-7, d 0-7), (keys 8-15, d 0-7), (keys 0-7, d 8-15), (keys 8-15, d 8-15)
                ldsm_x4_t(smem_u32(Vb + (kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * PITCH + np * 16 + (lane >> 4) * 8),
                          b0, b1, b2, b3);
                mma_bf16(o[2 * np], a0, a1, a2, a3, b0, b1);
                mma_bf16(o[2 * np + 1], a0, a1, a2, a3, b2, b3);
            }
        }
        __syncthreads();  // everyone done with `buf` before the next iteration's prefetch overwrites it
    }

    // ---- normalise, stage through this warp's own Q rows, coalesced 16-byte stores -------------
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
        lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
        lrow[r] = 1.0f / lrow[r];
    }
    __nv_bfloat16* Os = Qs + warp * 16 * PITCH;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
        *reinterpret_cast<uint32_t*>(Os + g * PITCH + i * 8 + 2 * t) = pack_bf16(o[i][0] * lrow[0], o[i][1] * lrow[0]);
        *reinterpret_cast<uint32_t*>(Os + (g + 8) * PITCH + i * 8 + 2 * t) = pack_bf16(o[i][2] * lrow[1], o[i][3] * lrow[1]);
    }
    __syncwarp();
    __nv_bfloat16* obase = a.o + h * D;
    for (int idx = lane; idx < 16 * CHUNKS; idx += 32) {
        const int r = idx / CHUNKS, c = idx % CHUNKS;
        const int sq = q0 + warp * 16 + r;
        if (sq < a.Sq) {
            const size_t row = (size_t)map_row(a.qmap, prob, sq);
            *reinterpret_cast<uint4*>(obase + row * a.ldo + c * 8) = *reinterpret_cast<const uint4*>(Os + r * PITCH + c * 8);
        }
    }
}

template <int D, int BM, int BN>
static int launch_att_t(const AttnArgs& a, cudaStream_t stream) {
    const int smem = (BM + 4 * BN) * (D + 8) * 2;
    static bool attr_set = false;
    if (!attr_set) {
        VPU_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<D, BM, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    dim3 grid((a.Sq + BM - 1) / BM, a.heads, a.nprob);
    VPU_CHECK_CUDA(launch_pdl(attention_kernel<D, BM, BN>, dim3(grid), dim3(BM * 2), smem, stream, a));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}
template <int D>
static int launch_att(const AttnArgs& a, cudaStream_t stream) {
    static const bool narrow = [] { const char* e = vpu_debug_env("VPU_ATTN_NARROW"); return !(e && e[0] == '0'); }();
    const bool q48 = narrow && a.Sq <= 48, k48 = narrow && a.Sk <= 48;
    if (q48 && k48) return launch_att_t<D, 48, 48>(a, stream);
    if (q48) return launch_att_t<D, 48, 64>(a, stream);
    if (k48) return launch_att_t<D, 64, 48>(a, stream);
    return launch_att_t<D, 64, 64>(a, stream);
}

#endif  // VPU_DEBUG

int attention_launch(const AttnArgs& a, int head_dim, cudaStream_t stream) {
    VPU_REQUIRE(a.Sq > 0 && a.Sk > 0 && a.nprob > 0 && a.heads > 0, "empty attention problem");
    VPU_REQUIRE(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 8 == 0 && a.qoff % 8 == 0 &&
                    a.koff % 8 == 0 && a.voff % 8 == 0,
                "attention strides/offsets must be multiples of 8 elements");
    VPU_REQUIRE(a.nprob <= 65535 && a.heads <= 65535, "attention grid too large");
    static const bool use_tc = [] { const char* e = vpu_debug_env("VPU_ATTN_TC"); return !(e && e[0] == '0'); }();
    if (use_tc && window_attention_tc_supported(a, head_dim)) return window_attention_tc_launch(a, stream);
    if (use_tc && window_attention_tc80_supported(a, head_dim)) return window_attention_tc80_launch(a, stream);
    if (use_tc && global_attention_tc80_supported(a, head_dim)) return global_attention_tc80_launch(a, stream);
    if (use_tc && global_attention_tc_supported(a, head_dim)) return global_attention_tc_launch(a, stream);
    if (use_tc && dma_attention_tc_supported(a, head_dim)) return dma_attention_tc_launch(a, head_dim, stream);
#ifdef VPU_DEBUG
    switch (head_dim) {
        case 48: return launch_att<48>(a, stream);
        case 64: return launch_att<64>(a, stream);
        case 80: return launch_att<80>(a, stream);
        case 96: return launch_att<96>(a, stream);
        case 128: return launch_att<128>(a, stream);
        case 160: return launch_att<160>(a, stream);
        default: VPU_REQUIRE(false, "unsupported head_dim %d (supported: 48,64,80,96,128,160)", head_dim);
    }
    return 0;
#else
    VPU_REQUIRE(false,
                "attention shape Sq=%d Sk=%d head_dim=%d heads=%d window=%d has no tcgen05 kernel (ViT window 14x14 d=64 / 16x16 d=80, "
                "global S %% 112 (d=64) or S %% 128 (d=80), DMA: 48 keys, or <= 128 queries over S %% 112 / S %% 64 keys)",
                a.Sq, a.Sk, head_dim, a.heads, a.qmap.mode == 1 ? a.qmap.win : 0);
    return 1;
#endif
}

}  // namespace vpu
