// C-ABI + host orchestration of the per-click forward (include/vpuformer_b200.h).
// Stage order follows reference is_vpu_model.py:422-438 / 383-419; every stage enqueues
// hand-written sm_100a kernels on the caller's stream.  No allocation, no sync, no fallback.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/vpuformer_b200.h"
#include "attention.cuh"
#include "elementwise.cuh"
#include "gemm.cuh"
#include "head_tail.cuh"
#include "noc.cuh"
#include "prompt.cuh"
#include "raster.cuh"
#include "session.cuh"

namespace vpu {
const char* last_error();

struct Tensor {
    const void* p = nullptr;
    int dtype = 0;
    std::vector<int64_t> shape;
};

struct Buf {
    std::string name;
    size_t off, bytes;
};

struct Plan {
    std::vector<Buf> bufs;
    size_t total = 0;
    size_t add(const char* name, size_t bytes) {
        total = (total + 1023) & ~size_t(1023);
        bufs.push_back({name, total, bytes});
        const size_t o = total;
        total += bytes;
        return o;
    }
    const Buf* find(const char* name) const {
        for (auto& b : bufs)
            if (b.name == name) return &b;
        return nullptr;
    }
};

// Optional per-kernel-class device timing (bench.py roofline): CUDA events recorded on the caller's
// stream around every launch of vpu_forward.  Off by default; event creation happens only here.
struct ProfRec {
    std::string cls;
    double flops, bytes;
    int launches;
    cudaEvent_t e0, e1;
};
struct Profiler {
    bool on = false;
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    cudaEvent_t get() {
        if (used == pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            pool.push_back(e);
        }
        return pool[used++];
    }
    ~Profiler() {
        for (cudaEvent_t e : pool) cudaEventDestroy(e);
    }
};

}  // namespace vpu

using namespace vpu;

struct vpu_context {
    vpu_dims d;
    std::unordered_map<std::string, Tensor> w;
    std::unordered_map<std::string, float> scalars;
    bool finalized = false;
    int gemm_impl = 0;
    float click_table[32];
    int click_radius = 9;
    Profiler prof;
    // small click batches (launch-latency-bound, replayed as CUDA graphs): independent parts of the forward -- the PPuE chain next
    // to the ViT trunk, the image-side K|V|Q projection next to the prompt self-attention, the four pyramid levels -- are issued on
    // side streams forked from / joined to the caller's stream with events, so a captured graph keeps them as parallel branches
    cudaStream_t side[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork[4] = {nullptr, nullptr, nullptr, nullptr}, ev_join[3] = {nullptr, nullptr, nullptr};
    ~vpu_context() {
        for (auto& st : side) if (st) cudaStreamDestroy(st);
        for (auto& e : ev_fork) if (e) cudaEventDestroy(e);
        for (auto& e : ev_join) if (e) cudaEventDestroy(e);
    }

    int C() const { return d.embed_dim; }
    int grid() const { return d.img_size / d.patch; }
    int N() const { return grid() * grid(); }
    int Q() const { return 2 * d.num_max_points; }
    int K0() const { return 6 * d.patch * d.patch; }    // fused (image | coords) patch-embed K
    int K0s() const { return 10 * d.patch * d.patch; }  // + split-bf16 residual planes of image and prev mask
    // row strides of the patch operand and of the two patch-embed weights: K rounded up to 64 elements, so that every row starts on a
    // 128-byte line (patch 14: K = 1960 / 1176 left the rows 16-byte aligned only and the ViT-H patch-embed GEMMs at half rate)
    int K0s_ld() const { return (K0s() + 63) / 64 * 64; }
    int K0_ld() const { return (K0() + 63) / 64 * 64; }
    int ppue_dim() const { return 2 * d.img_size + 3; }
    int ppue_ld() const { return (ppue_dim() + 7) / 8 * 8; }
    int d4() const { return std::max(d.out_dims[0] * 2, d.embed_dim / 2); }
    int d8() const { return std::max(d.out_dims[1], d.embed_dim / 2); }
    int d32() const { return std::max(d.out_dims[3], d.embed_dim * 2); }
    int group() const { return d.depth == 12 ? 6 : d.depth / 4; }
    // "vit.ln_fold" = 1: norm1 / norm2 of the ViT blocks are folded into the qkv / fc1 weights (packing.py) and evaluated in the
    // GEMM epilogues from row statistics the residual GEMMs leave behind (gemm.cuh Epi::ln_*): no LayerNorm pass over the tokens
    bool ln_fold() const { auto it = scalars.find("vit.ln_fold"); return it != scalars.end() && it->second != 0.f; }
};

namespace {

constexpr int TAB_PAD = 128;   // packing.py TAB_PAD: rows the positional tables repeat after their last row (gemm_res.cu MODE_TAB)

Plan make_plan(const vpu_context& h, int B) {
    Plan p;
    const size_t C = h.C(), N = h.N(), M = (size_t)B * N, Q = h.Q(), MQ = (size_t)B * Q, g = h.grid();
    const size_t g2 = 2 * g, g4 = 4 * g, gh = g / 2, hc = h.d.head_channels;
    p.add("ppue", (size_t)B * Q * h.ppue_dim() * 4);
    p.add("ppue_b", MQ * h.ppue_ld() * 2);
    p.add("A0", M * h.K0s_ld() * 2);
    p.add("X", M * C * 4);
    p.add("Xn", M * C * 2);
    p.add("lnstats", M * (size_t)gemm_ln_slots_max((int)C) * sizeof(float2));
    p.add("lnrow", M * sizeof(float2));
    p.add("QKV", M * 3 * C * 2);
    p.add("AO", M * C * 2);
    p.add("H", M * 4 * C * 2);
    // DMA
    p.add("T1", MQ * h.d.ppue_ffn_dim * 2);
    p.add("Q0", MQ * C * 4);
    p.add("Q0b", MQ * C * 2);
    p.add("Qt", MQ * C * 4);
    p.add("QQ", MQ * 2 * C * 2);        // [tokens + PE | tokens] bf16, side by side: the A operand of the merged token projections
    p.add("ql0", MQ * C * 4);
    p.add("ql1", MQ * C * 4);
    p.add("ql2", MQ * C * 4);
    p.add("qfin", MQ * C * 4);
    p.add("TOK", MQ * 4 * C * 2);       // merged token projections: i2t K | i2t V | self-attention Q | K | V (or the final q)
    p.add("SO", MQ * C * 2);
    p.add("T", MQ * C * 4);
    p.add("TQ", MQ * (C / 2) * 2);
    p.add("TO", MQ * (C / 2) * 2);
    p.add("MH", MQ * h.d.dma_mlp_dim * 2);
    p.add("X0b", M * C * 2);
    p.add("Kb", M * C * 2);
    p.add("KVQ", M * (3 * C / 2) * 2);
    p.add("IO", M * (C / 2) * 2);
    p.add("T2", M * C * 4);
    p.add("rowmax", 3 * (size_t)gemm_ln_parts((int)C) * M * 4);   // [layer][column block][token]: partial row maxima (gemm_ln.cu)
    p.add("qout", MQ * C * 4);
    p.add("qout_b", MQ * C * 2);
    p.add("cg", 3 * (size_t)B * C * 4);
    // merge + neck (NHWC bf16)
    p.add("x2", M * C * 2);
    p.add("x3", M * C * 2);
    p.add("x4", M * C * 2);
    p.add("D4a", (size_t)B * g2 * g2 * h.d4() * 2);
    p.add("D4b", (size_t)B * g4 * g4 * (h.d4() / 2) * 2);
    p.add("P4", (size_t)B * g4 * g4 * h.d.out_dims[0] * 2);
    p.add("D8a", (size_t)B * g2 * g2 * h.d8() * 2);
    p.add("P8", (size_t)B * g2 * g2 * h.d.out_dims[1] * 2);
    p.add("P16", M * h.d.out_dims[2] * 2);
    p.add("D32a", (size_t)B * gh * gh * h.d32() * 2);
    p.add("P32", (size_t)B * gh * gh * h.d.out_dims[3] * 2);
    p.add("gn_sums", (size_t)8 * B * 2 * sizeof(long long));    // 8 GroupNorms x [B][sum, sum of squares]
    // head
    const size_t res[4] = {g4, g2, g, gh};
    for (int i = 0; i < 4; ++i) {
        p.add(("HC" + std::to_string(i)).c_str(), (size_t)B * res[i] * res[i] * hc * 2);
        p.add(("Y" + std::to_string(i)).c_str(), (size_t)B * res[i] * res[i] * hc * 2);
    }
    p.add("F", (size_t)B * g4 * g4 * hc * 2);
    p.add("rnorm", (size_t)B * g4 * g4 * 4);
    p.add("QF", MQ * 2 * C * 2);
    p.add("QE", MQ * hc * 4);
    p.add("QN", (size_t)B * 64 * hc * 2);
    p.add("seg_low", (size_t)B * g4 * g4 * 4);
    p.add("aux_low", (size_t)B * Q * g4 * g4 * 4);
    p.total = (p.total + 1023) & ~size_t(1023);
    return p;
}

struct Fwd {
    vpu_context& h;
    cudaStream_t s;
    uint8_t* ws;
    Plan plan;
    int B;
    const char* stage = "";
    bool capturing = false; // the caller's stream is being captured into a CUDA graph: no event brackets
    std::string label;      // per-launch label of the profile's detail mode (scalar "debug.profile_detail" = 1): the weight key

    // run one launch helper; in profiling mode bracket it with events and book flops / algorithmic bytes
    template <typename F> int timed(const char* kind, double flops, double bytes, F&& fn) {
        if (!h.prof.on || capturing) return fn();
        ProfRec r;
        r.cls = std::string(kind) + "." + stage;
        if (!label.empty() && h.scalars.count("debug.profile_detail")) {     // "gemm.vit_window:blk.qkv.w": blocks pooled
            std::string l = label;
            if (l.compare(0, 3, "blk") == 0) l = "blk" + l.substr(l.find('.'));
            r.cls += ":" + l;
        }
        r.flops = flops; r.bytes = bytes;
        r.e0 = h.prof.get(); r.e1 = h.prof.get();
        const unsigned long long l0 = launch_count();
        cudaEventRecord(r.e0, s);
        const int rc = fn();
        cudaEventRecord(r.e1, s);
        r.launches = (int)(launch_count() - l0);
        h.prof.recs.push_back(r);
        return rc;
    }

    template <typename T> T* buf(const char* name) { return reinterpret_cast<T*>(ws + plan.find(name)->off); }
    template <typename T> const T* W(const std::string& key) { return reinterpret_cast<const T*>(h.w.at(key).p); }
    const __nv_bfloat16* Wb(const std::string& key) { return W<__nv_bfloat16>(key); }
    const float* Wf(const std::string& key) { return W<float>(key); }

    // GroupNorm fusion arguments of a neck GEMM (Epi::gn_*): statistics out, and / or the producer's GroupNorm folded in
    struct Gn {
        long long* out = nullptr;
        const long long* in = nullptr;
        const float* wg = nullptr;
        int rows = 0;
        double in_count = 0;
    };
    void set_gn(Epi& e, const Gn& g) {
        e.gn_out = g.out; e.gn_in = g.in; e.gn_wg = g.wg; e.gn_rows = g.rows; e.gn_in_count = (float)g.in_count;
    }
    // LayerNorm fusion arguments of a ViT GEMM (Epi::ln_*)
    struct Ln {
        float2* out = nullptr;
        __nv_bfloat16* out_bf16 = nullptr;
        float2* row = nullptr;         // finalised (rstd, mean * rstd) per row, written after a GEMM with `out`
        const float2* in = nullptr;    // the same buffer on the consuming side
        const float* s = nullptr;
        float eps = 1e-6f;
        bool unused = false;           // producing side: nobody reads these statistics (the last block's fc2): no finaliser launch
    };
    // The row statistics of the LayerNorm fusion are finalised by ln_rowstats_kernel, a launch of its own after every residual GEMM.
    // Finalising them inside the producing GEMM (the last-arriving column tile of a 32-row block adds the slots, arrival counter
    // + __threadfence) was tried in round 2: +0.3 ms per batch-64 step in three A/B alternations (the fence makes every epilogue
    // warp of proj / fc2 wait for its own 40 KB of tile stores) and no gain at batch 2; removed.  Finalising them in the CONSUMER's
    // epilogue for small problems (each lane one row from the raw slots, results exchanged by shuffles; same arithmetic, same bits)
    // was slower too: batch 2 1.297 -> 1.318 ms, batch 8 2.465 -> 2.684 ms.
    // out = act(A W^T + bias [+ tab] [+ res])
    int gemm(const __nv_bfloat16* A, int lda, const std::string& wkey, int M, int Nn, int K, const float* bias, void* out,
             bool out_bf16, int ldo, int act = ACT_NONE, const void* res = nullptr, bool res_bf16 = false, int ldr = 0,
             const float* tab = nullptr, int tab_rows = 0, const Gn* gn = nullptr, const Ln* ln = nullptr, int tab_pad = 0) {
        GemmProblem p;
        p.A = A; p.W = Wb(wkey); p.M = M; p.N = Nn; p.K = K; p.lda = lda;
        p.ldw = (int)h.w.at(wkey).shape[1];
        p.w_rows = Nn;
        p.epi.out = out; p.epi.out_bf16 = out_bf16; p.epi.ldo = ldo; p.epi.bias = bias; p.epi.act = act;
        p.epi.res = res; p.epi.res_bf16 = res_bf16; p.epi.ldr = ldr; p.epi.bias2d = tab; p.epi.bias2d_rows = tab_rows; p.epi.bias2d_pad_rows = tab_pad;
        if (gn) set_gn(p.epi, *gn);
        if (ln) {
            p.epi.ln_out = ln->out; p.epi.ln_out_bf16 = ln->out_bf16; p.epi.ln_in = ln->in; p.epi.ln_s = ln->s;
            p.epi.ln_slots = gemm_ln_slots(M, h.C());
        }
        const double by = 2.0 * ((double)M * K + (double)Nn * K) + (double)M * Nn * ((out_bf16 ? 2 : 4) + (res ? (res_bf16 ? 2 : 4) : 0)) +
                          (ln && ln->out ? 2.0 * M * Nn : 0.0);
        label = wkey + " " + std::to_string(M) + "x" + std::to_string(Nn) + "x" + std::to_string(K);
        const int grc = timed("gemm", 2.0 * M * Nn * K, by, [&] { return gemm_launch(p, s, h.gemm_impl); });
        label.clear();
        if (grc) return grc;
        if (ln && ln->out && !ln->unused)      // finalise the row statistics for the LayerNorm-folded GEMM that follows
            return timed("lnstats", 0, (double)M * (p.epi.ln_slots + 1) * 8.0,
                         [&] { return ln_rowstats_launch(ln->out, M, p.epi.ln_slots, h.C(), ln->eps, ln->row, s); });
        return 0;
    }
    int gemm_ps(const __nv_bfloat16* A, const std::string& wkey, const float* bias4, int M, int cout, int K, int g,
                __nv_bfloat16* out, const Gn* gn = nullptr) {
        GemmProblem p;
        p.A = A; p.W = Wb(wkey); p.M = M; p.N = 4 * cout; p.K = K; p.lda = K; p.ldw = K; p.w_rows = 4 * cout;
        p.epi.out = out; p.epi.out_bf16 = 1; p.epi.ldo = cout; p.epi.bias = bias4; p.epi.mode = EPI_PIXEL_SHUFFLE;
        p.epi.ps_g = g; p.epi.ps_cout = cout;
        if (gn) set_gn(p.epi, *gn);
        const double by = 2.0 * ((double)M * K + 4.0 * cout * K + 4.0 * M * cout);
        label = wkey + " " + std::to_string(M) + "x" + std::to_string(4 * cout) + "x" + std::to_string(K);
        const int grc = timed("gemm", 8.0 * M * cout * K, by, [&] { return gemm_launch(p, s, h.gemm_impl); });
        label.clear();
        return grc;
    }
    int ln(const float* in, const std::string& key, float eps, int rows, float* of, __nv_bfloat16* ob,
           const float* pe = nullptr, __nv_bfloat16* ope = nullptr, float* rowmax = nullptr, int ld_bf16 = 0) {
        LnArgs a;
        a.in = in; a.gamma = Wf(key + ".g"); a.beta = Wf(key + ".b"); a.eps = eps; a.rows = rows;
        a.out_f32 = of; a.out_bf16 = ob; a.pe = pe; a.out_pe_bf16 = ope; a.rowmax = rowmax; a.ld_bf16 = ld_bf16;
        const double by = (double)rows * h.C() * (4 + (of ? 4 : 0) + (ob ? 2 : 0) + (ope ? 6 : 0));
        return timed("ln", 0, by, [&] { return layernorm_launch(a, h.C(), s); });
    }
    // GroupNorm apply (+GELU) with the statistics the producing GEMM accumulated: one read + one write of x
    int gn(__nv_bfloat16* x, size_t per_sample, int Cc, const std::string& key, int gelu, const long long* sums) {
        return timed("gn", 0, 4.0 * B * (double)per_sample, [&] {
            return groupnorm_apply_launch(x, B, per_sample, Cc, Wf(key + ".g"), Wf(key + ".b"), gelu, sums, s);
        });
    }
    int attn(const __nv_bfloat16* q, int ldq, int qoff, const __nv_bfloat16* k, int ldk, int koff, const __nv_bfloat16* v,
             int ldv, int voff, __nv_bfloat16* o, int ldo, int Sq, int Sk, int heads, int hd, int nprob, float scale,
             bool windowed) {
        AttnArgs a;
        a.q = q; a.k = k; a.v = v; a.o = o; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo;
        a.qoff = qoff; a.koff = koff; a.voff = voff; a.Sq = Sq; a.Sk = Sk; a.heads = heads; a.nprob = nprob;
        a.scale_log2 = scale * 1.4426950408889634f;
        if (windowed) {
            a.qmap.mode = 1; a.qmap.tokens = h.N(); a.qmap.grid = h.grid(); a.qmap.win = 224 / h.d.patch;
            a.kmap = a.qmap;
        } else {
            a.qmap.mode = 0; a.qmap.per_prob = Sq;
            a.kmap.mode = 0; a.kmap.per_prob = Sk;
        }
        const double fl = 4.0 * Sq * Sk * hd * heads * nprob;
        const double by = 2.0 * hd * heads * nprob * (2.0 * Sq + 2.0 * Sk);
        return timed("attn", fl, by, [&] { return attention_launch(a, hd, s); });
    }
};

#define RUN(x)                  \
    do {                        \
        if (int _rc = (x)) return _rc; \
    } while (0)

int fill_ppue_args(const vpu_context& h, const vpu_prompts& pr, PpueArgs& a) {
    const bool own = pr.type != 0 && pr.ppue_points != nullptr;
    a.points = own ? pr.ppue_points : pr.points;
    a.n = own ? pr.n_ppue : pr.n;
    a.num_max_points = h.d.num_max_points;
    a.size = h.d.img_size;
    a.type = pr.type;
    a.boxes = pr.boxes; a.scrib_sel = pr.scrib_sel; a.scrib_slot = pr.scrib_slot;
    a.click_radius = h.click_radius;
    memcpy(a.click_table, h.click_table, sizeof(a.click_table));
    VPU_REQUIRE(a.points != nullptr, "prompts.points is NULL");
    VPU_REQUIRE(pr.type >= 0 && pr.type <= 2, "as_prompt_type %d not in {0,1,2}", pr.type);
    return 0;
}

int run_forward(vpu_context& h, const float* image4, const vpu_prompts& pr, int B, float* instances, float* aux,
                uint8_t* ws, cudaStream_t s) {
    pdl_set_auto((size_t)B * h.N() <= 16384);     // small batches are launch-latency-bound (B=2: -16 %, 8: -9 %, 16: -4 %): see pdl_enabled()
    Fwd f{h, s, ws, make_plan(h, B), B};
    {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        VPU_CHECK_CUDA(cudaStreamIsCapturing(s, &cs));
        f.capturing = cs != cudaStreamCaptureStatusNone;
    }
    const int C = h.C(), N = h.N(), M = B * N, Q = h.Q(), MQ = B * Q, g = h.grid(), img = h.d.img_size;
    const int heads = h.d.num_heads, hd = C / heads;
    typedef __nv_bfloat16 bf;

    // side streams for small batches (vpu_context::side): fork = the side stream continues from the main stream's current position,
    // join = the main stream waits for everything issued on the side stream so far
    const cudaStream_t s_main = s;
    static const bool branch_knob = [] { const char* e = vpu_debug_env("VPU_FWD_BRANCHES"); return !(e && e[0] == '0'); }();   // A/B knob, -DVPU_DEBUG builds
    const bool branches = branch_knob && (size_t)B * h.N() <= 16384 && !h.prof.on && h.side[0] && !h.scalars.count("debug.stop_after_block");
    auto on = [&](cudaStream_t t) { s = t; f.s = t; };
    auto fork = [&](int k, int ev) -> int {
        VPU_CHECK_CUDA(cudaEventRecord(h.ev_fork[ev], s_main));
        VPU_CHECK_CUDA(cudaStreamWaitEvent(h.side[k], h.ev_fork[ev], 0));
        return 0;
    };
    auto join = [&](int k) -> int {
        VPU_CHECK_CUDA(cudaEventRecord(h.ev_join[k], h.side[k]));
        VPU_CHECK_CUDA(cudaStreamWaitEvent(s_main, h.ev_join[k], 0));
        return 0;
    };

    // GroupNorm accumulators of the neck: cleared here so that no memset node sits between two kernels later on
    // (a non-kernel node would break the programmatic-dependent-launch chain, common.cuh)
    VPU_CHECK_CUDA(cudaMemsetAsync(f.buf<long long>("gn_sums"), 0, (size_t)8 * B * 2 * sizeof(long long), s));
    if (branches) VPU_CHECK_CUDA(cudaEventRecord(h.ev_fork[0], s_main));      // the PPuE chain depends on nothing in this forward
    // ---- A1-A3, A7: fused image + coord-feature patch operand, one GEMM for both patch embeds ----
    CoordArgs ca;
    ca.image4 = image4; ca.points = pr.points; ca.extra_mask = pr.extra_mask; ca.n = pr.n; ca.H = img; ca.W = img;
    ca.radius = h.d.norm_radius;
    f.stage = "patch_embed";
    RUN(f.timed("patch_operand", 0, (double)B * img * img * (16.0 + 20.0), [&] {
        return patch_operand_launch(ca, B, f.buf<bf>("A0"), h.d.patch, h.K0s_ld(), s);
    }));
    // 3-term split-bf16 product (A_hi + A_lo)(W_hi + W_lo) ~ A_hi W_hi + A_lo W_hi + A_hi W_lo: the token
    // embedding feeds the fp32 residual stream of every block, so it is computed to ~fp32 accuracy.
    float* X = f.buf<float>("X");
    RUN(f.gemm(f.buf<bf>("A0"), h.K0s_ld(), "pe.w", M, C, h.K0s(), nullptr, X, false, C, ACT_NONE, nullptr, false, 0,
               f.Wf("pe.tab"), N));
    const bool fold = h.ln_fold();
    typedef Fwd::Ln Ln;
    float2* lnst = f.buf<float2>("lnstats");
    bf* X0b = f.buf<bf>("X0b");
    Ln ln_out;                       // residual GEMMs: fp32 X + bf16 copy + row statistics for the LayerNorm that follows
    ln_out.out = lnst; ln_out.out_bf16 = f.buf<bf>("Xn"); ln_out.row = f.buf<float2>("lnrow");
    if (fold)
        RUN(f.gemm(f.buf<bf>("A0"), h.K0s_ld(), "pe.w_lo", M, C, h.K0(), f.Wf("pe.zero_b"), X, false, C, ACT_NONE, X, false, C, nullptr, 0,
                   nullptr, &ln_out));
    else
        RUN(f.gemm(f.buf<bf>("A0"), h.K0s_ld(), "pe.w_lo", M, C, h.K0(), nullptr, X, false, C, ACT_NONE, X, false, C));

    // ---- A8-A9: ViT blocks ----
    bf* Xn = f.buf<bf>("Xn");
    bf* QKV = f.buf<bf>("QKV");
    bf* AO = f.buf<bf>("AO");
    bf* Hh = f.buf<bf>("H");
    const int win = 224 / h.d.patch, nwin = (g / win) * (g / win);
    // parity taps only: "debug.stop_after_block" = k ends the forward after k ViT blocks (X holds the tokens)
    const int stop_after = h.scalars.count("debug.stop_after_block") ? (int)h.scalars.at("debug.stop_after_block") : -1;
    if (stop_after == 0) return 0;
    for (int i = 1; i <= h.d.depth; ++i) {
        const std::string k = "blk" + std::to_string(i - 1);
        const bool windowed = (i % h.group()) != 0;
        f.stage = windowed ? "vit_window" : "vit_global";
        Ln ln_in;
        ln_in.in = f.buf<float2>("lnrow");
        if (fold) {
            ln_in.s = f.Wf(k + ".qkv.s");
            RUN(f.gemm(Xn, C, k + ".qkv.w", M, 3 * C, C, f.Wf(k + ".qkv.b"), QKV, true, 3 * C, ACT_NONE, nullptr, false, 0, nullptr, 0,
                       nullptr, &ln_in));
        } else {
            RUN(f.ln(X, k + ".ln1", 1e-6f, M, nullptr, Xn));
            RUN(f.gemm(Xn, C, k + ".qkv.w", M, 3 * C, C, f.Wf(k + ".qkv.b"), QKV, true, 3 * C));
        }
        if (windowed)
            RUN(f.attn(QKV, 3 * C, 0, QKV, 3 * C, C, QKV, 3 * C, 2 * C, AO, C, win * win, win * win, heads, hd, B * nwin,
                       1.0f / sqrtf((float)hd), true));
        else
            RUN(f.attn(QKV, 3 * C, 0, QKV, 3 * C, C, QKV, 3 * C, 2 * C, AO, C, N, N, heads, hd, B, 1.0f / sqrtf((float)hd),
                       false));
        if (fold) {
            RUN(f.gemm(AO, C, k + ".proj.w", M, C, C, f.Wf(k + ".proj.b"), X, false, C, ACT_NONE, X, false, C, nullptr, 0, nullptr, &ln_out));
            ln_in.s = f.Wf(k + ".fc1.s");
            RUN(f.gemm(Xn, C, k + ".fc1.w", M, 4 * C, C, f.Wf(k + ".fc1.b"), Hh, true, 4 * C, ACT_GELU, nullptr, false, 0, nullptr, 0,
                       nullptr, &ln_in));
            // the last block's bf16 copy is the DMA stage's key operand
            Ln lo = ln_out;
            if (i == h.d.depth) { lo.out_bf16 = X0b; lo.unused = true; }
            RUN(f.gemm(Hh, 4 * C, k + ".fc2.w", M, C, 4 * C, f.Wf(k + ".fc2.b"), X, false, C, ACT_NONE, X, false, C, nullptr, 0, nullptr, &lo));
        } else {
            RUN(f.gemm(AO, C, k + ".proj.w", M, C, C, f.Wf(k + ".proj.b"), X, false, C, ACT_NONE, X, false, C));
            RUN(f.ln(X, k + ".ln2", 1e-6f, M, nullptr, Xn));
            RUN(f.gemm(Xn, C, k + ".fc1.w", M, 4 * C, C, f.Wf(k + ".fc1.b"), Hh, true, 4 * C, ACT_GELU));
            RUN(f.gemm(Hh, 4 * C, k + ".fc2.w", M, C, 4 * C, f.Wf(k + ".fc2.b"), X, false, C, ACT_NONE, X, false, C));
        }
        if (stop_after == i) return 0;
    }

    // ---- A4-A6, A10: PPuE rows + FFN (small batches: a branch parallel to the ViT trunk) ----
    if (branches) {
        VPU_CHECK_CUDA(cudaStreamWaitEvent(h.side[0], h.ev_fork[0], 0));
        on(h.side[0]);
    }
    PpueArgs pa;
    RUN(fill_ppue_args(h, pr, pa));
    pa.out = f.buf<float>("ppue");
    pa.out_bf16 = f.buf<bf>("ppue_b");
    pa.ld_bf16 = h.ppue_ld();
    f.stage = "ppue";
    RUN(f.timed("ppue", 0, (double)MQ * (h.ppue_dim() * 4.0 + h.ppue_ld() * 2.0), [&] { return ppue_launch(pa, B, s); }));
    float* Q0 = f.buf<float>("Q0");
    RUN(f.gemm(f.buf<bf>("ppue_b"), h.ppue_ld(), "ffn.w1", MQ, h.d.ppue_ffn_dim, h.ppue_ld(), f.Wf("ffn.b1"), f.buf<bf>("T1"),
               true, h.d.ppue_ffn_dim, ACT_RELU));
    RUN(f.gemm(f.buf<bf>("T1"), h.d.ppue_ffn_dim, "ffn.w2", MQ, C, h.d.ppue_ffn_dim, f.Wf("ffn.b2"), Q0, false, C));

    // ---- A11-A13: Dual-cross Merging Attention ----
    bf* Q0b = f.buf<bf>("Q0b");
    // the two bf16 views of the prompt tokens live side by side, [tokens + PE | tokens] with row stride 2 C: every projection that
    // reads the same token state runs as ONE GEMM over that pair (block weights [W 0] / [0 W], packing.py), 9 launches fewer
    bf* QPb = f.buf<bf>("QQ");
    bf* Qb = QPb + C;
    const int ldq = 2 * (int)C;
    bf* TOK = f.buf<bf>("TOK");
    const int ldt = 4 * (int)C;
    float* Qt = f.buf<float>("Qt");
    float* T = f.buf<float>("T");
    bf* Kb = f.buf<bf>("Kb");
    bf* KVQ = f.buf<bf>("KVQ");
    float* rowmax = f.buf<float>("rowmax");
    f.stage = "dma";
    RUN(f.timed("cast", 0, 6.0 * MQ * C, [&] { return cast_add_launch(Q0, nullptr, Q0b, (size_t)MQ * C, s); }));
    if (branches) {
        on(s_main);
        RUN(join(0));
    }
    if (!fold || stop_after > 0)
        RUN(f.timed("cast", 0, 6.0 * M * C, [&] { return cast_add_launch(X, nullptr, X0b, (size_t)M * C, s); }));
    const int dh = h.d.dma_heads, Ci = C / 2, dself = C / dh, dcross = Ci / dh;
    const float* Qf = Q0;  // current fp32 queries
    const bf* Kin = X0b;   // current bf16 keys
    float* ql[3] = {f.buf<float>("ql0"), f.buf<float>("ql1"), f.buf<float>("ql2")};
    static const bool use_gemm_ln = [] { const char* e = vpu_debug_env("VPU_DMA_GEMM_LN"); return !(e && e[0] == '0'); }();   // A/B knob, -DVPU_DEBUG builds
    const int rm_parts = gemm_ln_parts((int)C);
    // TOK columns: [0, Ci) i2t K, [Ci, C) i2t V, [C, 3C) self-attention Q | K, [3C, 4C) self-attention V (final layer: [C, C + Ci) = final q)
    bf* SQK = TOK + C;
    bf* SV = TOK + 3 * C;
    // layer 0: no positional term, q, k and v all read the PPuE tokens (transformer.py:436-442)
    RUN(f.gemm(Q0b, C, "dma0.sa.qkv.w", MQ, 3 * C, C, f.Wf("dma0.sa.qkv.b"), SQK, true, ldt));
    for (int j = 0; j < h.d.dma_depth; ++j) {
        const std::string k = "dma" + std::to_string(j);
        const bool last = j + 1 == h.d.dma_depth;
        // the image-side K|V (and step 4's Q) of this layer come from one GEMM whose positional term key_pe W^T is a precomputed
        // additive table (transformer.py:444-449); it depends on the keys only: small batches run it beside steps (1) - (2a)
        auto img_projection = [&]() -> int {
            return f.gemm(Kin, C, k + ".img.w", M, 3 * Ci, C, nullptr, KVQ, true, 3 * Ci, ACT_NONE, nullptr, false, 0,
                          f.Wf(k + ".img.tab"), N, nullptr, nullptr, TAB_PAD);
        };
        if (branches) {
            RUN(fork(1, 1));
            on(h.side[1]);
            RUN(img_projection());
            on(s_main);
        }
        // (1) prompt self-attention (layer 0: output replaces the queries)
        RUN(f.attn(SQK, ldt, 0, SQK, ldt, C, SV, ldt, 0, f.buf<bf>("SO"), C, Q, Q, dh, dself, B, 1.0f / sqrtf((float)dself), false));
        RUN(f.gemm(f.buf<bf>("SO"), C, k + ".sa.o.w", MQ, C, C, f.Wf(k + ".sa.o.b"), T, false, C, ACT_NONE,
                   j == 0 ? nullptr : Qf, false, C));
        RUN(f.ln(T, k + ".n1", 1e-5f, MQ, Qt, Qb, Q0, QPb, nullptr, ldq));
        Qf = Qt;
        // (2) tokens -> image cross attention
        RUN(f.gemm(QPb, ldq, k + ".t2i.q.w", MQ, Ci, C, f.Wf(k + ".t2i.q.b"), f.buf<bf>("TQ"), true, Ci));
        if (branches) RUN(join(1));
        else RUN(img_projection());
        RUN(f.attn(f.buf<bf>("TQ"), Ci, 0, KVQ, 3 * Ci, 0, KVQ, 3 * Ci, Ci, f.buf<bf>("TO"), Ci, Q, N, dh, dcross, B,
                   1.0f / sqrtf((float)dcross), false));
        RUN(f.gemm(f.buf<bf>("TO"), Ci, k + ".t2i.o.w", MQ, C, Ci, f.Wf(k + ".t2i.o.b"), T, false, C, ACT_NONE, Qf, false, C));
        RUN(f.ln(T, k + ".n2", 1e-5f, MQ, Qt, Qb, Q0, QPb, nullptr, ldq));
        // (3) MLP (transformer.py:451-454)
        RUN(f.gemm(Qb, ldq, k + ".mlp.w1", MQ, h.d.dma_mlp_dim, C, f.Wf(k + ".mlp.b1"), f.buf<bf>("MH"), true, h.d.dma_mlp_dim,
                   ACT_RELU));
        RUN(f.gemm(f.buf<bf>("MH"), h.d.dma_mlp_dim, k + ".mlp.w2", MQ, C, h.d.dma_mlp_dim, f.Wf(k + ".mlp.b2"), T, false, C,
                   ACT_NONE, Qf, false, C));
        RUN(f.ln(T, k + ".n3", 1e-5f, MQ, ql[j], Qb, Q0, QPb, nullptr, ldq));
        Qf = ql[j];
        // (4) image -> tokens cross attention (transformer.py:456-461).  Its K (tokens + PE) and V (tokens) projections, the next
        //     layer's self-attention q | k (tokens + PE) and v (tokens) -- or the final attention's q -- all read this token state
        RUN(f.gemm(QPb, ldq, k + ".tok.w", MQ, last ? 3 * Ci : 4 * C, 2 * C, f.Wf(k + ".tok.b"), TOK, true, ldt));
        RUN(f.attn(KVQ, 3 * Ci, 2 * Ci, TOK, ldt, 0, TOK, ldt, Ci, f.buf<bf>("IO"), Ci, N, Q, dh, dcross, B,
                   1.0f / sqrtf((float)dcross), false));
        //     out-projection + residual + norm4 in one kernel (gemm_ln.cu); the row maxima the merge gates need come out as one
        //     partial per 256-column block
        GemmLn gl;
        gl.A = f.buf<bf>("IO"); gl.lda = Ci; gl.W = f.Wb(k + ".i2t.o.w"); gl.ldw = Ci; gl.bias = f.Wf(k + ".i2t.o.b");
        gl.res = Kin; gl.ldr = C; gl.gamma = f.Wf(k + ".n4.g"); gl.beta = f.Wf(k + ".n4.b"); gl.out = Kb; gl.ldo = C;
        gl.rowmax_parts = rowmax + (size_t)j * rm_parts * M; gl.M = M; gl.K = Ci; gl.C = C;
        if (use_gemm_ln && gemm_ln_supported(gl)) {
            f.label = k + ".i2t.o+n4 " + std::to_string(M) + "x" + std::to_string(C) + "x" + std::to_string(Ci);
            const int grc = f.timed("gemm_ln", 2.0 * M * C * Ci, (double)M * (2.0 * Ci + 4.0 * C) + 2.0 * C * Ci,
                                    [&] { return gemm_ln_launch(gl, s); });
            f.label.clear();
            if (grc) return grc;
        } else {
            RUN(f.gemm(f.buf<bf>("IO"), Ci, k + ".i2t.o.w", M, C, Ci, f.Wf(k + ".i2t.o.b"), f.buf<float>("T2"), false, C, ACT_NONE,
                       Kin, true, C));
            RUN(f.ln(f.buf<float>("T2"), k + ".n4", 1e-5f, M, nullptr, Kb, nullptr, nullptr, rowmax + (size_t)j * rm_parts * M));
            for (int q = 1; q < rm_parts; ++q)      // the other partial maxima: copies of the full-row maximum
                VPU_CHECK_CUDA(cudaMemcpyAsync(rowmax + ((size_t)j * rm_parts + q) * M, rowmax + (size_t)j * rm_parts * M, (size_t)M * 4,
                                               cudaMemcpyDeviceToDevice, s));
        }
        Kin = Kb;
    }
    // final tokens -> image attention (transformer.py:374-379); its q sits in TOK[:, C : C + Ci)
    RUN(f.gemm(Kin, C, "dmaf.img.w", M, 2 * Ci, C, nullptr, KVQ, true, 2 * Ci, ACT_NONE, nullptr, false, 0,
               f.Wf("dmaf.img.tab"), N, nullptr, nullptr, TAB_PAD));
    RUN(f.attn(TOK, ldt, C, KVQ, 2 * Ci, 0, KVQ, 2 * Ci, Ci, f.buf<bf>("TO"), Ci, Q, N, dh, dcross, B,
               1.0f / sqrtf((float)dcross), false));
    RUN(f.gemm(f.buf<bf>("TO"), Ci, "dmaf.o.w", MQ, C, Ci, f.Wf("dmaf.o.b"), T, false, C, ACT_NONE, Qf, false, C));
    RUN(f.ln(T, "dmaf.n", 1e-5f, MQ, f.buf<float>("qfin"), nullptr));

    // ---- A14: merge ----
    f.stage = "merge";
    RUN(f.timed("qout_gate", 0, 22.0 * MQ * C, [&] {
        return qout_gate_launch(Q0, ql[0], ql[1], f.buf<float>("qfin"), B, Q, C, f.buf<float>("qout"), f.buf<bf>("qout_b"),
                                f.buf<float>("cg"), s);
    }));
    MergeArgs ma;
    ma.x = X; ma.cg = f.buf<float>("cg"); ma.rowmax = rowmax; ma.rowmax_parts = rm_parts; ma.x2 = f.buf<bf>("x2"); ma.x3 = f.buf<bf>("x3");
    ma.x4_s2d = f.buf<bf>("x4"); ma.B = B; ma.N = N; ma.M = M; ma.C = C; ma.grid = g;
    RUN(f.timed("merge", 0, 10.0 * M * C, [&] { return merge_launch(ma, s); }));

    // ---- A15: 4-scale pyramid (NHWC bf16; ConvT/Conv with stride == kernel are plain GEMMs) ----
    // small batches: the four levels (neck chain + head conv pair) are independent until the head tail: one stream each
    cudaStream_t lvl[4] = {s_main, s_main, s_main, s_main};
    if (branches) {
        for (int i = 0; i < 3; ++i) {
            RUN(fork(i, 2));
            lvl[i + 1] = h.side[i];
        }
    }
    f.stage = "neck";
    const int d4 = h.d4(), d8 = h.d8(), d32 = h.d32();
    const int* od = h.d.out_dims;
    const size_t g2 = 2 * g, g4 = 4 * g, gh = g / 2;
    // GroupNorm(1, C) is fused into the GEMMs on both sides (Epi::gn_*): every GEMM accumulates the per-sample sum / sum of
    // squares of its fp32 outputs in its epilogue, the three GroupNorms that are not followed by GELU (d4.gn2, d8.gn1,
    // d32.gn1) are folded into the consuming 1x1 conv, and the other five need one apply (+GELU) pass and no statistics pass.
    long long* sums = f.buf<long long>("gn_sums");     // zeroed at the top of the forward
    auto S = [&](int i) { return sums + (size_t)i * B * 2; };
    typedef Fwd::Gn Gn;
    Gn gn;
    on(lvl[0]);
    gn = Gn(); gn.out = S(0); gn.rows = N;
    RUN(f.gemm_ps(X0b, "d4.a.w", f.Wf("d4.a.b"), M, d4, C, g, f.buf<bf>("D4a"), &gn));
    RUN(f.gn(f.buf<bf>("D4a"), g2 * g2 * d4, d4, "d4.gn1", 1, S(0)));
    gn = Gn(); gn.out = S(1); gn.rows = (int)(g2 * g2);
    RUN(f.gemm_ps(f.buf<bf>("D4a"), "d4.b.w", f.Wf("d4.b.b"), (int)(B * g2 * g2), d4 / 2, d4, (int)g2, f.buf<bf>("D4b"), &gn));
    gn = Gn(); gn.in = S(1); gn.in_count = (double)(g4 * g4) * (d4 / 2); gn.wg = f.Wf("d4.c.wg"); gn.out = S(2); gn.rows = (int)(g4 * g4);
    RUN(f.gemm(f.buf<bf>("D4b"), d4 / 2, "d4.c.w", (int)(B * g4 * g4), od[0], d4 / 2, f.Wf("d4.c.b"), f.buf<bf>("P4"), true, od[0],
               ACT_NONE, nullptr, false, 0, nullptr, 0, &gn));
    RUN(f.gn(f.buf<bf>("P4"), g4 * g4 * od[0], od[0], "d4.gn3", 1, S(2)));

    on(lvl[1]);
    gn = Gn(); gn.out = S(3); gn.rows = N;
    RUN(f.gemm_ps(f.buf<bf>("x2"), "d8.a.w", f.Wf("d8.a.b"), M, d8, C, g, f.buf<bf>("D8a"), &gn));
    gn = Gn(); gn.in = S(3); gn.in_count = (double)(g2 * g2) * d8; gn.wg = f.Wf("d8.b.wg"); gn.out = S(4); gn.rows = (int)(g2 * g2);
    RUN(f.gemm(f.buf<bf>("D8a"), d8, "d8.b.w", (int)(B * g2 * g2), od[1], d8, f.Wf("d8.b.b"), f.buf<bf>("P8"), true, od[1],
               ACT_NONE, nullptr, false, 0, nullptr, 0, &gn));
    RUN(f.gn(f.buf<bf>("P8"), g2 * g2 * od[1], od[1], "d8.gn2", 1, S(4)));

    on(lvl[2]);
    gn = Gn(); gn.out = S(5); gn.rows = N;
    RUN(f.gemm(f.buf<bf>("x3"), C, "d16.a.w", M, od[2], C, f.Wf("d16.a.b"), f.buf<bf>("P16"), true, od[2], ACT_NONE, nullptr, false, 0,
               nullptr, 0, &gn));
    RUN(f.gn(f.buf<bf>("P16"), (size_t)N * od[2], od[2], "d16.gn1", 1, S(5)));

    on(lvl[3]);
    gn = Gn(); gn.out = S(6); gn.rows = (int)(gh * gh);
    RUN(f.gemm(f.buf<bf>("x4"), 4 * C, "d32.a.w", (int)(B * gh * gh), d32, 4 * C, f.Wf("d32.a.b"), f.buf<bf>("D32a"), true, d32,
               ACT_NONE, nullptr, false, 0, nullptr, 0, &gn));
    gn = Gn(); gn.in = S(6); gn.in_count = (double)(gh * gh) * d32; gn.wg = f.Wf("d32.b.wg"); gn.out = S(7); gn.rows = (int)(gh * gh);
    RUN(f.gemm(f.buf<bf>("D32a"), d32, "d32.b.w", (int)(B * gh * gh), od[3], d32, f.Wf("d32.b.b"), f.buf<bf>("P32"), true, od[3],
               ACT_NONE, nullptr, false, 0, nullptr, 0, &gn));
    RUN(f.gn(f.buf<bf>("P32"), gh * gh * od[3], od[3], "d32.gn2", 1, S(7)));

    // ---- A16: head ----
    f.stage = "head";
    const int hc = h.d.head_channels;
    const char* pyr[4] = {"P4", "P8", "P16", "P32"};
    const size_t res[4] = {g4, g2, (size_t)g, gh};
    HeadCombineArgs hca;
    for (int i = 0; i < 4; ++i) {
        const std::string si = std::to_string(i);
        const int rows = (int)(B * res[i] * res[i]);
        on(lvl[i]);
        GemmB2B bb;
        bb.A = f.buf<bf>(pyr[i]); bb.W1 = f.Wb("hd.c" + si + ".w"); bb.W2 = f.Wb("hd.f" + si + ".w"); bb.bias1 = f.Wf("hd.c" + si + ".b");
        bb.out = f.buf<bf>(("Y" + si).c_str()); bb.M = rows; bb.K1 = od[i]; bb.lda = od[i]; bb.ldo = hc;
        static const bool use_b2b = [] { const char* e = vpu_debug_env("VPU_HEAD_B2B"); return !(e && e[0] == '0'); }();   // A/B knob, -DVPU_DEBUG builds
        if (use_b2b && gemm_b2b_supported(bb, hc, hc)) {      // conv + ReLU + fusion-conv slice back to back, the intermediate stays on chip
            f.label = "hd.c+f" + si + " " + std::to_string(rows) + "x" + std::to_string(hc) + "x" + std::to_string(od[i]);
            const double by = 2.0 * ((double)rows * od[i] + (double)hc * od[i] + (double)hc * hc + (double)rows * hc);
            const int brc = f.timed("gemm", 2.0 * rows * hc * ((double)od[i] + hc), by, [&] { return gemm_b2b_launch(bb, s); });
            f.label.clear();
            if (brc) return brc;
        } else {
            RUN(f.gemm(f.buf<bf>(pyr[i]), od[i], "hd.c" + si + ".w", rows, hc, od[i], f.Wf("hd.c" + si + ".b"),
                       f.buf<bf>(("HC" + si).c_str()), true, hc, ACT_RELU));
            RUN(f.gemm(f.buf<bf>(("HC" + si).c_str()), hc, "hd.f" + si + ".w", rows, hc, hc, nullptr, f.buf<bf>(("Y" + si).c_str()),
                       true, hc));
        }
        hca.y[i] = f.buf<bf>(("Y" + si).c_str());
        hca.res[i] = (int)res[i];
    }
    on(s_main);
    if (branches)
        for (int i = 0; i < 3; ++i) RUN(join(i));
    HeadTailArgs ht;
    for (int i = 0; i < 4; ++i) { ht.y[i] = hca.y[i]; ht.res[i] = hca.res[i]; }
    ht.B = B; ht.channels = hc; ht.bias = f.Wf("hd.f.b"); ht.wseg = f.Wf("hd.seg.w"); ht.seg_bias = h.scalars.at("hd.seg.b");
    ht.seg_out = f.buf<float>("seg_low"); ht.nq = Q;
    if (aux) {   // P2CL queries (swin_transformer.py:745-750); skipped when the caller only reads 'instances'
        RUN(f.gemm(f.buf<bf>("qout_b"), C, "hd.q.w1", MQ, 2 * C, C, f.Wf("hd.q.b1"), f.buf<bf>("QF"), true, 2 * C, ACT_RELU));
        RUN(f.gemm(f.buf<bf>("QF"), 2 * C, "hd.q.w2", MQ, hc, 2 * C, f.Wf("hd.q.b2"), f.buf<float>("QE"), false, hc));
        RUN(f.timed("head_queries", 0, (double)B * 64 * hc * 6.0, [&] { return head_queries_launch(f.buf<float>("QE"), f.Wf("hd.seg.w"), B, Q, f.buf<bf>("QN"), s); }));
        ht.qn = f.buf<bf>("QN"); ht.aux_out = f.buf<float>("aux_low");
    }
    static const bool use_head_tail = [] { const char* e = vpu_debug_env("VPU_HEAD_TAIL"); return !(e && e[0] == '0'); }();   // A/B knob, -DVPU_DEBUG builds
    if (use_head_tail && head_tail_supported(ht)) {
        // resize + sum + ReLU + conv_seg + cosine logits on the tensor core: f never exists in memory (head_tail.cu)
        const double px = (double)B * g4 * g4;
        const double by = (double)B * hc * 2.0 * ((double)g4 * g4 + g2 * g2 + (double)g * g + gh * gh) + px * 4.0 + (aux ? px * Q * 4.0 : 0.0);
        RUN(f.timed("head_tail", 2.0 * px * hc * (240.0 + (aux ? Q : 0)), by, [&] { return head_tail_launch(ht, s); }));
    } else {
        hca.bias = f.Wf("hd.f.b"); hca.out = f.buf<bf>("F"); hca.rnorm = f.buf<float>("rnorm"); hca.B = B;
        hca.wseg = f.Wf("hd.seg.w"); hca.seg_bias = h.scalars.at("hd.seg.b"); hca.seg_out = f.buf<float>("seg_low");
        RUN(f.timed("head_combine", 0, (double)B * hc * 2.0 * (2.0 * g4 * g4 + g2 * g2 + (double)g * g + gh * gh), [&] { return head_combine_launch(hca, s); }));
        if (aux) {
            GemmProblem p;
            p.A = f.buf<bf>("F"); p.W = f.buf<bf>("QN"); p.M = (int)(B * g4 * g4); p.N = 64; p.K = hc; p.lda = hc; p.ldw = hc;
            p.w_rows = B * 64;
            p.epi.mode = EPI_HEAD_FINAL; p.epi.m_per_batch = (int)(g4 * g4); p.epi.b_rows_per_batch = 64;
            p.epi.rnorm = f.buf<float>("rnorm"); p.epi.aux_out = f.buf<float>("aux_low");
            p.epi.seg_out = nullptr; p.epi.seg_bias = 0.f; p.epi.nq = Q;
            const double by = (double)p.M * (hc * 2.0 + 4.0 + Q * 4.0);
            RUN(f.timed("gemm", 2.0 * p.M * 64.0 * hc, by, [&] { return gemm_launch(p, s, h.gemm_impl); }));
        }
    }
    // ---- A17: final upsampling (align_corners=True) ----
    f.stage = "final";
    RUN(f.timed("upsample_seg", 0, 4.0 * B * ((double)g4 * g4 + (double)img * img), [&] {
        return upsample_ac_launch(f.buf<float>("seg_low"), instances, (int)g4, (int)g4, img, img, (size_t)B, s);
    }));
    if (aux)
        RUN(f.timed("upsample_aux", 0, 4.0 * B * Q * ((double)g4 * g4 + (double)img * img), [&] {
            return upsample_ac_launch(f.buf<float>("aux_low"), aux, (int)g4, (int)g4, img, img, (size_t)B * Q, s);
        }));
    return 0;
}

struct Need {
    std::string key;
    int dtype;
    std::vector<int64_t> shape;
};

std::vector<Need> needed_weights(const vpu_context& h) {
    std::vector<Need> v;
    const int64_t C = h.C(), N = h.N(), Ci = C / 2;
    auto lin = [&](const std::string& k, int64_t o, int64_t i) {
        v.push_back({k + ".w", VPU_BF16, {o, i}});
        v.push_back({k + ".b", VPU_F32, {o}});
    };
    auto lin2 = [&](const std::string& k, const char* s, int64_t o, int64_t i) {  // key.w1/.b1 style
        v.push_back({k + ".w" + s, VPU_BF16, {o, i}});
        v.push_back({k + ".b" + s, VPU_F32, {o}});
    };
    auto nrm = [&](const std::string& k, int64_t c) {
        v.push_back({k + ".g", VPU_F32, {c}});
        v.push_back({k + ".b", VPU_F32, {c}});
    };
    v.push_back({"pe.w", VPU_BF16, {C, h.K0s_ld()}});        // zero columns past K (packing.py)
    v.push_back({"pe.w_lo", VPU_BF16, {C, h.K0_ld()}});
    v.push_back({"pe.tab", VPU_F32, {N, C}});
    if (h.ln_fold()) v.push_back({"pe.zero_b", VPU_F32, {C}});
    for (int i = 0; i < h.d.depth; ++i) {
        const std::string k = "blk" + std::to_string(i);
        if (h.ln_fold()) {       // qkv / fc1 hold W diag(gamma) and W beta + b; ".s" = row sums of the bf16 W diag(gamma)
            v.push_back({k + ".qkv.s", VPU_F32, {3 * C}});
            v.push_back({k + ".fc1.s", VPU_F32, {4 * C}});
        } else {
            nrm(k + ".ln1", C); nrm(k + ".ln2", C);
        }
        lin(k + ".qkv", 3 * C, C); lin(k + ".proj", C, C); lin(k + ".fc1", 4 * C, C); lin(k + ".fc2", C, 4 * C);
    }
    lin2("ffn", "1", h.d.ppue_ffn_dim, h.ppue_ld());
    lin2("ffn", "2", C, h.d.ppue_ffn_dim);
    for (int j = 0; j < h.d.dma_depth; ++j) {
        const std::string k = "dma" + std::to_string(j);
        if (j == 0) lin(k + ".sa.qkv", 3 * C, C);
        lin(k + ".sa.o", C, C); nrm(k + ".n1", C);
        lin(k + ".t2i.q", Ci, C);
        v.push_back({k + ".img.w", VPU_BF16, {3 * Ci, C}});
        v.push_back({k + ".img.tab", VPU_F32, {N + TAB_PAD, 3 * Ci}});   // rows 0 .. TAB_PAD-1 repeated at the end
        lin(k + ".t2i.o", C, Ci); nrm(k + ".n2", C);
        lin2(k + ".mlp", "1", h.d.dma_mlp_dim, C); lin2(k + ".mlp", "2", C, h.d.dma_mlp_dim); nrm(k + ".n3", C);
        lin(k + ".tok", j + 1 == h.d.dma_depth ? 3 * Ci : 4 * C, 2 * C);   // merged token projections (packing.py) lin(k + ".i2t.o", C, Ci); nrm(k + ".n4", C);
    }
    v.push_back({"dmaf.img.w", VPU_BF16, {2 * Ci, C}});
    v.push_back({"dmaf.img.tab", VPU_F32, {N + TAB_PAD, 2 * Ci}});
    lin("dmaf.o", C, Ci); nrm("dmaf.n", C);
    const int64_t d4 = h.d4(), d8 = h.d8(), d32 = h.d32();
    const int* od = h.d.out_dims;
    // d4.gn2 / d8.gn1 / d32.gn1 are folded into the weights of the conv that follows them (".wg" = row sums of W diag(gamma))
    lin("d4.a", 4 * d4, C); nrm("d4.gn1", d4); lin("d4.b", 4 * (d4 / 2), d4);
    lin("d4.c", od[0], d4 / 2); v.push_back({"d4.c.wg", VPU_F32, {od[0]}}); nrm("d4.gn3", od[0]);
    lin("d8.a", 4 * d8, C); lin("d8.b", od[1], d8); v.push_back({"d8.b.wg", VPU_F32, {od[1]}}); nrm("d8.gn2", od[1]);
    lin("d16.a", od[2], C); nrm("d16.gn1", od[2]);
    lin("d32.a", d32, 4 * C); lin("d32.b", od[3], d32); v.push_back({"d32.b.wg", VPU_F32, {od[3]}}); nrm("d32.gn2", od[3]);
    const int64_t hc = h.d.head_channels;
    for (int i = 0; i < 4; ++i) {
        lin("hd.c" + std::to_string(i), hc, od[i]);
        v.push_back({"hd.f" + std::to_string(i) + ".w", VPU_BF16, {hc, hc}});
    }
    v.push_back({"hd.f.b", VPU_F32, {hc}});
    lin2("hd.q", "1", 2 * C, C);
    lin2("hd.q", "2", hc, 2 * C);
    v.push_back({"hd.seg.w", VPU_F32, {hc}});
    return v;
}

bool valid_handle(vpu_handle h) {
    if (!h) { set_error("null handle"); return false; }
    return true;
}

}  // namespace

extern "C" {

const char* vpu_last_error(void) { return vpu::last_error(); }
int vpu_version(void) { return 1; }
unsigned long long vpu_launch_count(void) { return vpu::launch_count(); }

int vpu_profile_begin(vpu_handle h) {
    if (!valid_handle(h)) return 1;
    h->prof.on = true;
    h->prof.recs.clear();
    h->prof.used = 0;
    return 0;
}

int vpu_profile_end(vpu_handle h, vpu_profile_entry* out, int max_entries, int* n_out) {
    if (!valid_handle(h)) return 1;
    VPU_REQUIRE(out && n_out && max_entries > 0, "vpu_profile_end: bad argument");
    h->prof.on = false;
    std::vector<std::string> order;
    std::unordered_map<std::string, vpu_profile_entry> agg;
    if (!h->prof.recs.empty()) VPU_CHECK_CUDA(cudaEventSynchronize(h->prof.recs.back().e1));
    for (const ProfRec& r : h->prof.recs) {
        float ms = 0.f;
        VPU_CHECK_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
        auto it = agg.find(r.cls);
        if (it == agg.end()) {
            vpu_profile_entry e;
            memset(&e, 0, sizeof(e));
            strncpy(e.name, r.cls.c_str(), sizeof(e.name) - 1);
            it = agg.emplace(r.cls, e).first;
            order.push_back(r.cls);
        }
        it->second.ms += ms; it->second.flops += r.flops; it->second.bytes += r.bytes; it->second.launches += r.launches;
    }
    int n = 0;
    for (const std::string& k : order) {
        if (n == max_entries) break;
        out[n++] = agg[k];
    }
    *n_out = n;
    h->prof.recs.clear();
    h->prof.used = 0;
    return 0;
}

int vpu_create(vpu_handle* out, const vpu_dims* dims) {
    VPU_REQUIRE(out && dims, "vpu_create: null argument");
    const vpu_dims& d = *dims;
    VPU_REQUIRE(d.img_size > 0 && d.patch > 0 && d.img_size % d.patch == 0, "bad image/patch size");
    VPU_REQUIRE(d.embed_dim == 768 || d.embed_dim == 1024 || d.embed_dim == 1280, "embed_dim %d not in {768,1024,1280}", d.embed_dim);
    VPU_REQUIRE(d.depth % 4 == 0, "depth must be a multiple of 4 (reference models_vit.py:264)");
    VPU_REQUIRE(d.embed_dim % d.num_heads == 0 && d.embed_dim % (2 * d.dma_heads) == 0, "head counts must divide embed_dim");
    VPU_REQUIRE(d.head_channels == 256, "head_channels must be 256 (upsample='x1')");
    VPU_REQUIRE(d.num_max_points >= 1 && d.num_max_points <= 24, "num_max_points must be in [1, 24]");
    VPU_REQUIRE((d.img_size / d.patch) % (224 / d.patch) == 0 && (d.img_size / d.patch) % 2 == 0, "grid must tile into 224-px windows");
    vpu_context* h = new vpu_context();
    h->d = d;
    const char* impl = vpu_debug_env("VPU_GEMM_IMPL");
    h->gemm_impl = (impl && impl[0] == '1') ? 1 : 0;
    // click Gaussian table: the reference's float32 formula (ops.py:51-61), sigma = 3, peak + 1
    const int r = 9;
    h->click_radius = r;
    for (int k = 0; k < 32; ++k) h->click_table[k] = 0.f;
    for (int k = 0; k <= 2 * r; ++k) {
        const float kk = (float)k - (float)r;
        const float sq = kk * kk;
        h->click_table[k] = (float)std::exp((double)(-sq / 18.0f));
    }
    h->click_table[r] += 1.0f;
    *out = h;
    return 0;
}

void vpu_destroy(vpu_handle h) { delete h; }

int vpu_bind_weight(vpu_handle h, const char* key, const void* dev_ptr, int dtype, const int64_t* shape, int rank) {
    if (!valid_handle(h)) return 1;
    VPU_REQUIRE(key && dev_ptr && rank >= 0 && rank <= 4, "vpu_bind_weight: bad argument");
    Tensor t;
    t.p = dev_ptr; t.dtype = dtype;
    t.shape.assign(shape, shape + rank);
    h->w[key] = t;
    h->finalized = false;
    return 0;
}

int vpu_set_scalar(vpu_handle h, const char* key, float value) {
    if (!valid_handle(h)) return 1;
    h->scalars[key] = value;
    return 0;
}

int vpu_set_click_table(vpu_handle h, const float* host_table, int taps) {
    if (!valid_handle(h)) return 1;
    VPU_REQUIRE(host_table && taps >= 1 && taps <= 31 && (taps & 1), "click table must have an odd number of taps <= 31");
    for (int k = 0; k < 32; ++k) h->click_table[k] = k < taps ? host_table[k] : 0.f;
    h->click_radius = taps / 2;
    return 0;
}

int vpu_finalize(vpu_handle h) {
    if (!valid_handle(h)) return 1;
    for (const Need& n : needed_weights(*h)) {
        auto it = h->w.find(n.key);
        VPU_REQUIRE(it != h->w.end(), "weight '%s' is not bound", n.key.c_str());
        VPU_REQUIRE(it->second.dtype == n.dtype, "weight '%s' has dtype %d, expected %d", n.key.c_str(), it->second.dtype, n.dtype);
        VPU_REQUIRE(it->second.shape == n.shape, "weight '%s' has the wrong shape", n.key.c_str());
        VPU_REQUIRE((reinterpret_cast<uintptr_t>(it->second.p) & 15) == 0, "weight '%s' is not 16-byte aligned", n.key.c_str());
    }
    VPU_REQUIRE(h->scalars.count("hd.seg.b"), "scalar 'hd.seg.b' is not set");
    if (int rc = gemm_init()) return rc;
    for (auto& st : h->side) if (!st) VPU_CHECK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (auto& e : h->ev_fork) if (!e) VPU_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : h->ev_join) if (!e) VPU_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (h->d.head_channels == 256 && (4 * h->grid()) % 16 == 0)       // interpolation-matrix table of the fused head tail
        if (int rc = head_tail_prepare(4 * h->grid())) return rc;
    h->finalized = true;
    return 0;
}

size_t vpu_workspace_bytes(vpu_handle h, int B) {
    if (!valid_handle(h) || B <= 0) return 0;
    return make_plan(*h, B).total;
}

int vpu_workspace_lookup(vpu_handle h, int B, const char* name, size_t* offset, size_t* bytes) {
    if (!valid_handle(h)) return 1;
    VPU_REQUIRE(B > 0 && name && offset && bytes, "vpu_workspace_lookup: bad argument");
    Plan p = make_plan(*h, B);
    const Buf* b = p.find(name);
    VPU_REQUIRE(b != nullptr, "no workspace buffer named '%s'", name);
    *offset = b->off;
    *bytes = b->bytes;
    return 0;
}

static int check_prompts(vpu_handle h, const vpu_prompts* pr) {
    VPU_REQUIRE(pr && pr->points, "prompts / prompts.points is NULL");
    VPU_REQUIRE(pr->n >= 1 && pr->n <= h->d.num_max_points, "points per half n=%d outside [1, %d]", pr->n, h->d.num_max_points);
    VPU_REQUIRE(pr->type >= 0 && pr->type <= 2, "as_prompt_type %d not in {0,1,2}", pr->type);
    if (pr->type == 1) VPU_REQUIRE(pr->boxes, "as_prompt_type 1 needs boxes");
    if (pr->type == 2) VPU_REQUIRE(pr->scrib_sel && pr->scrib_slot, "as_prompt_type 2 needs scrib_sel and scrib_slot");
    return 0;
}

int vpu_forward(vpu_handle h, const float* image4, const vpu_prompts* prompts, int B, float* instances, float* instances_aux,
                void* workspace, size_t workspace_bytes, void* stream) {
    if (!valid_handle(h)) return 1;
    VPU_REQUIRE(h->finalized, "vpu_forward before vpu_finalize");
    VPU_REQUIRE(image4 && instances && workspace && B > 0, "vpu_forward: null argument");
    if (int rc = check_prompts(h, prompts)) return rc;
    VPU_REQUIRE(workspace_bytes >= vpu_workspace_bytes(h, B), "workspace too small: %zu < %zu", workspace_bytes, vpu_workspace_bytes(h, B));
    VPU_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "workspace must be 1024-byte aligned");
    return run_forward(*h, image4, *prompts, B, instances, instances_aux, reinterpret_cast<uint8_t*>(workspace),
                       reinterpret_cast<cudaStream_t>(stream));
}

int vpu_ppue(vpu_handle h, const vpu_prompts* prompts, int B, float* out, void* stream) {
    if (!valid_handle(h)) return 1;
    if (int rc = check_prompts(h, prompts)) return rc;
    VPU_REQUIRE(out && B > 0, "vpu_ppue: null argument");
    PpueArgs a;
    if (int rc = fill_ppue_args(*h, *prompts, a)) return rc;
    a.out = out;
    return ppue_launch(a, B, reinterpret_cast<cudaStream_t>(stream));
}

int vpu_coord_features(vpu_handle h, const float* image4, const vpu_prompts* prompts, int B, float* out, void* stream) {
    if (!valid_handle(h)) return 1;
    if (int rc = check_prompts(h, prompts)) return rc;
    VPU_REQUIRE(image4 && out && B > 0, "vpu_coord_features: null argument");
    CoordArgs ca;
    ca.image4 = image4; ca.points = prompts->points; ca.extra_mask = prompts->extra_mask; ca.n = prompts->n;
    ca.H = h->d.img_size; ca.W = h->d.img_size; ca.radius = h->d.norm_radius;
    return coord_features_launch(ca, B, out, reinterpret_cast<cudaStream_t>(stream));
}

int vpu_gemm(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, const float* bias2d,
             int bias2d_rows, const void* residual, int residual_dtype, int ldr, int act, void* out, int out_dtype, int ldo,
             int impl, void* stream) {
    VPU_REQUIRE(A && W && out, "vpu_gemm: null argument");
    GemmProblem p;
    p.A = reinterpret_cast<const __nv_bfloat16*>(A); p.W = reinterpret_cast<const __nv_bfloat16*>(W);
    p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldw = ldw; p.w_rows = N;
    p.epi.out = out; p.epi.out_bf16 = out_dtype == VPU_BF16; p.epi.ldo = ldo; p.epi.bias = bias; p.epi.bias2d = bias2d;
    p.epi.bias2d_rows = bias2d_rows; p.epi.res = residual; p.epi.res_bf16 = residual_dtype == VPU_BF16; p.epi.ldr = ldr;
    p.epi.act = act;
    return gemm_launch(p, reinterpret_cast<cudaStream_t>(stream), impl);
}

int vpu_gemm_table(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* table, int table_rows,
                   int table_pad_rows, void* out, int ldo, int impl, void* stream) {
    VPU_REQUIRE(A && W && out && table && table_rows > 0 && table_pad_rows >= 0, "vpu_gemm_table: bad argument");
    GemmProblem p;
    p.A = reinterpret_cast<const __nv_bfloat16*>(A); p.W = reinterpret_cast<const __nv_bfloat16*>(W);
    p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldw = ldw; p.w_rows = N;
    p.epi.out = out; p.epi.out_bf16 = 1; p.epi.ldo = ldo; p.epi.bias2d = table; p.epi.bias2d_rows = table_rows;
    p.epi.bias2d_pad_rows = table_pad_rows;
    return gemm_launch(p, reinterpret_cast<cudaStream_t>(stream), impl);
}

int vpu_gemm_layernorm(const void* A, int lda, const void* W, int ldw, const float* bias, const void* res, int ldr, const float* gamma,
                       const float* beta, float eps, int M, int K, int C, void* out, int ldo, float* rowmax_parts, void* stream) {
    VPU_REQUIRE(A && W && bias && res && gamma && beta && out, "vpu_gemm_layernorm: null argument");
    GemmLn p;
    p.A = reinterpret_cast<const __nv_bfloat16*>(A); p.W = reinterpret_cast<const __nv_bfloat16*>(W); p.bias = bias;
    p.res = reinterpret_cast<const __nv_bfloat16*>(res); p.gamma = gamma; p.beta = beta; p.eps = eps;
    p.out = reinterpret_cast<__nv_bfloat16*>(out); p.rowmax_parts = rowmax_parts;
    p.M = M; p.K = K; p.C = C; p.lda = lda; p.ldw = ldw; p.ldr = ldr; p.ldo = ldo;
    return gemm_ln_launch(p, reinterpret_cast<cudaStream_t>(stream));
}

int vpu_gemm_b2b(const void* A, int lda, const void* W1, const float* bias1, const void* W2, int M, int K1, void* out, int ldo,
                 void* stream) {
    VPU_REQUIRE(A && W1 && W2 && bias1 && out, "vpu_gemm_b2b: null argument");
    GemmB2B p;
    p.A = reinterpret_cast<const __nv_bfloat16*>(A); p.W1 = reinterpret_cast<const __nv_bfloat16*>(W1);
    p.W2 = reinterpret_cast<const __nv_bfloat16*>(W2); p.bias1 = bias1; p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.M = M; p.K1 = K1; p.lda = lda; p.ldo = ldo;
    return gemm_b2b_launch(p, reinterpret_cast<cudaStream_t>(stream));
}

int vpu_head_tail(const void* y0, const void* y1, const void* y2, const void* y3, int B, int res0, const float* bias, const float* wseg,
                  float seg_bias, const void* qn, int nq, float* seg_out, float* aux_out, void* stream) {
    VPU_REQUIRE(y0 && y1 && y2 && y3 && bias && wseg && seg_out, "vpu_head_tail: null argument");
    VPU_REQUIRE(!aux_out || qn, "vpu_head_tail: aux_out needs the normalised queries");
    HeadTailArgs a;
    a.y[0] = reinterpret_cast<const __nv_bfloat16*>(y0); a.y[1] = reinterpret_cast<const __nv_bfloat16*>(y1);
    a.y[2] = reinterpret_cast<const __nv_bfloat16*>(y2); a.y[3] = reinterpret_cast<const __nv_bfloat16*>(y3);
    for (int l = 0; l < 4; ++l) a.res[l] = res0 >> l;
    a.B = B; a.channels = 256; a.bias = bias; a.wseg = wseg; a.seg_bias = seg_bias;
    a.qn = reinterpret_cast<const __nv_bfloat16*>(qn); a.nq = nq; a.seg_out = seg_out; a.aux_out = aux_out;
    return head_tail_launch(a, reinterpret_cast<cudaStream_t>(stream));
}

int vpu_gemm_pixel_shuffle(const void* A, const void* W, int M, int cout, int K, const float* bias4, int g, void* out, int impl,
                           void* stream) {
    VPU_REQUIRE(A && W && out, "vpu_gemm_pixel_shuffle: null argument");
    GemmProblem p;
    p.A = reinterpret_cast<const __nv_bfloat16*>(A); p.W = reinterpret_cast<const __nv_bfloat16*>(W);
    p.M = M; p.N = 4 * cout; p.K = K; p.lda = K; p.ldw = K; p.w_rows = 4 * cout;
    p.epi.out = out; p.epi.out_bf16 = 1; p.epi.ldo = cout; p.epi.bias = bias4; p.epi.mode = EPI_PIXEL_SHUFFLE;
    p.epi.ps_g = g; p.epi.ps_cout = cout;
    return gemm_launch(p, reinterpret_cast<cudaStream_t>(stream), impl);
}

int vpu_attention(const void* q, int ldq, int qoff, const void* k, int ldk, int koff, const void* v, int ldv, int voff, void* o,
                  int ldo, int Sq, int Sk, int heads, int head_dim, int nprob, float scale, int window, int grid, void* stream) {
    VPU_REQUIRE(q && k && v && o, "vpu_attention: null argument");
    AttnArgs a;
    a.q = reinterpret_cast<const __nv_bfloat16*>(q); a.k = reinterpret_cast<const __nv_bfloat16*>(k);
    a.v = reinterpret_cast<const __nv_bfloat16*>(v); a.o = reinterpret_cast<__nv_bfloat16*>(o);
    a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo; a.qoff = qoff; a.koff = koff; a.voff = voff;
    a.Sq = Sq; a.Sk = Sk; a.heads = heads; a.nprob = nprob; a.scale_log2 = scale * 1.4426950408889634f;
    if (window > 0) {
        VPU_REQUIRE(grid % window == 0 && Sq == window * window && Sk == Sq, "windowed attention: Sq must equal window^2");
        a.qmap.mode = 1; a.qmap.tokens = grid * grid; a.qmap.grid = grid; a.qmap.win = window;
        a.kmap = a.qmap;
    } else {
        a.qmap.per_prob = Sq;
        a.kmap.per_prob = Sk;
    }
    return attention_launch(a, head_dim, reinterpret_cast<cudaStream_t>(stream));
}

size_t vpu_noc_workspace_bytes(int S, int H, int W) { return S > 0 && H > 0 && W > 0 ? noc_workspace_bytes(S, H, W) : 0; }

int vpu_noc_next_clicks(const int8_t* gt, const uint8_t* pred, uint8_t* not_clicked, int S, int H, int W, int32_t* clicks,
                        int64_t* iou_counts, void* workspace, size_t workspace_bytes, void* stream) {
    VPU_REQUIRE(gt && pred && not_clicked && clicks && iou_counts && workspace, "vpu_noc_next_clicks: null argument");
    VPU_REQUIRE(workspace_bytes >= vpu_noc_workspace_bytes(S, H, W), "vpu_noc_next_clicks: workspace too small");
    VPU_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "vpu_noc_next_clicks: workspace must be 256-byte aligned");
    return noc_next_clicks_launch(gt, pred, not_clicked, S, H, W, clicks, reinterpret_cast<long long*>(iou_counts), workspace,
                                  reinterpret_cast<cudaStream_t>(stream));
}

int vpu_raster_prompts(int type, const int32_t* boxes, const int32_t* scribbles, int S, int n, int B, int size, uint8_t* planes,
                       void* stream) {
    RasterArgs a;
    a.type = type; a.boxes = boxes; a.scribbles = scribbles; a.S = S; a.n = n; a.size = size; a.planes = planes;
    return raster_prompts_launch(a, B, reinterpret_cast<cudaStream_t>(stream));
}

static SessionState to_session(const vpu_session_state* st) {
    SessionState s;
    s.S = st->S; s.H = st->H; s.W = st->W; s.T = st->T; s.max_clicks = st->max_clicks; s.n_half = st->n_half;
    s.images = st->images; s.prev_probs = st->prev_probs; s.pred = st->pred; s.clicks = st->clicks; s.nclicks = st->nclicks;
    s.roi = st->roi; s.fgbox = st->fgbox; s.pred_thr = st->pred_thr; s.zoom_thr = st->zoom_thr;
    s.expansion_ratio = st->expansion_ratio; s.recompute_thresh_iou = st->recompute_thresh_iou; s.min_crop_size = st->min_crop_size;
    return s;
}

int vpu_session_prepare(const vpu_session_state* st, const int32_t* active, int A, const int32_t* new_clicks, float* net_image,
                        double* net_points, void* stream) {
    VPU_REQUIRE(st, "vpu_session_prepare: null state");
    return session_prepare_launch(to_session(st), active, A, new_clicks, net_image, net_points, reinterpret_cast<cudaStream_t>(stream));
}

int vpu_session_finish(const vpu_session_state* st, const int32_t* active, int A, const float* logits, void* stream) {
    VPU_REQUIRE(st, "vpu_session_finish: null state");
    return session_finish_launch(to_session(st), active, A, logits, reinterpret_cast<cudaStream_t>(stream));
}

int vpu_image_from_u8(const uint8_t* rgb_nhwc, const float* prev_mask, float* image4, int B, int H, int W, void* stream) {
    return image_from_u8_launch(rgb_nhwc, prev_mask, image4, B, H, W, reinterpret_cast<cudaStream_t>(stream));
}

int vpu_debug_attention_trace(void* dev_buf, int cap) {
    attention_debug_trace(reinterpret_cast<unsigned long long*>(dev_buf), dev_buf ? cap : 0);
    return 0;
}

int vpu_layernorm(const float* in, const float* gamma, const float* beta, float eps, int rows, int C, float* out_f32,
                  void* out_bf16, const float* pe, void* out_pe_bf16, float* rowmax, void* stream) {
    VPU_REQUIRE(in && gamma && beta, "vpu_layernorm: null argument");
    LnArgs a;
    a.in = in; a.gamma = gamma; a.beta = beta; a.eps = eps; a.rows = rows; a.out_f32 = out_f32;
    a.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16); a.pe = pe;
    a.out_pe_bf16 = reinterpret_cast<__nv_bfloat16*>(out_pe_bf16); a.rowmax = rowmax;
    return layernorm_launch(a, C, reinterpret_cast<cudaStream_t>(stream));
}

int vpu_groupnorm_nhwc(void* x, int B, int64_t per_sample, int C, const float* gamma, const float* beta, int gelu, void* scratch,
                       void* stream) {
    VPU_REQUIRE(x && gamma && beta && scratch, "vpu_groupnorm_nhwc: null argument");
    float2* partial = reinterpret_cast<float2*>(scratch);
    float2* stats = partial + (size_t)B * GN_MAX_CHUNKS;
    return groupnorm_launch(reinterpret_cast<__nv_bfloat16*>(x), B, (size_t)per_sample, C, gamma, beta, gelu, partial, stats,
                            reinterpret_cast<cudaStream_t>(stream));
}

int vpu_upsample_align_corners(const float* in, float* out, int h, int w, int H, int W, int64_t planes, void* stream) {
    VPU_REQUIRE(in && out, "vpu_upsample_align_corners: null argument");
    return upsample_ac_launch(in, out, h, w, H, W, (size_t)planes, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
