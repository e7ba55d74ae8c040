/* vpuformer_b200 -- C ABI of the B200-native per-click VPUFormer forward.
 *
 * The reference (XuZhang1211/PVPUFormer) has no FFI: its boundary for this path is the Python
 * call  net(image, points, prompts, as_prompt_type) -> {'instances', 'instances_aux'}
 * (reference isegm/model/is_vpu_model.py:422-438, called from
 * isegm/inference/predictors/base.py:104,163,177).  This header is what a binding of that call
 * targets; pvpuformer_b200/model.py is the ctypes binding.
 *
 * Conventions: every function returns 0 on success, non-zero on error with a thread-local
 * message in vpu_last_error().  No exceptions cross the ABI.  All data pointers are DEVICE
 * pointers owned by the caller (torch tensors kept alive by the caller) unless marked host.
 * Nothing allocates, synchronises or uses the default stream inside vpu_forward: all work is
 * enqueued on `stream` (a cudaStream_t passed as void*).  One handle per (device, thread).
 * There is no CPU or library fallback: on a non-sm_100 device every compute entry fails.
 */
#ifndef VPUFORMER_B200_H
#define VPUFORMER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vpu_context* vpu_handle;

/* Model geometry: reference models_vit.py:306-319 (ViT-B/L/H factories),
 * models/iSegNet/vpu_base448_cocolvis.py:11-56, is_vpu_model.py:19-86 (neck), swin_transformer.py:666-721 (head). */
typedef struct vpu_dims {
    int32_t img_size;        /* 448 */
    int32_t patch;           /* 16 (B/L) or 14 (H) */
    int32_t embed_dim;       /* 768 / 1024 / 1280 */
    int32_t depth;           /* 12 / 24 / 32 */
    int32_t num_heads;       /* 12 / 16 / 16 */
    int32_t num_max_points;  /* 24 -> 48 prompt queries */
    int32_t dma_depth;       /* 3 */
    int32_t dma_heads;       /* 8 */
    int32_t dma_mlp_dim;     /* 1024 */
    int32_t ppue_ffn_dim;    /* 2048 */
    int32_t head_channels;   /* 256 */
    int32_t out_dims[4];     /* 128, 256, 512, 1024 */
    float norm_radius;       /* 5 (click disk radius, is_model.py:10, ops.py:375) */
} vpu_dims;

enum { VPU_F32 = 0, VPU_BF16 = 1, VPU_I32 = 2, VPU_F64 = 3, VPU_U8 = 4 };

const char* vpu_last_error(void);
int vpu_version(void);
/* Number of CUDA kernels this library has launched in this process (bench.py "gpu_launches"). */
unsigned long long vpu_launch_count(void);

/* ---- lifecycle (replaces VitMultiGaussianVector_ed_Model.__init__ / load_state_dict,
 *      reference is_vpu_model.py:142-186, inference/utils.py:21-46) ---- */
int vpu_create(vpu_handle* out, const vpu_dims* dims);
void vpu_destroy(vpu_handle h);
/* Bind one packed weight tensor (device memory, kept alive by the caller).  Keys and layouts are
 * listed in DESIGN.md ("packed weights"); pvpuformer_b200/packing.py produces them from a
 * reference state_dict. */
int vpu_bind_weight(vpu_handle h, const char* key, const void* dev_ptr, int dtype, const int64_t* shape, int rank);
int vpu_set_scalar(vpu_handle h, const char* key, float value);
/* Optional: HOST pointer to the PPuE click Gaussian taps (reference ops.py:51-61 computes them with
 * numpy float32; passing numpy's values makes the click rows bit-identical).  Default: same formula in C. */
int vpu_set_click_table(vpu_handle h, const float* host_table, int taps);
/* Check that every weight the forward needs is bound with the right dtype/shape. */
int vpu_finalize(vpu_handle h);

/* ---- the hot path (replaces forward(), reference is_vpu_model.py:422-438) ---- */
size_t vpu_workspace_bytes(vpu_handle h, int B);
/* Offset/size of a named intermediate inside the workspace (parity taps). */
int vpu_workspace_lookup(vpu_handle h, int B, const char* name, size_t* offset, size_t* bytes);

typedef struct vpu_prompts {
    const double* points;        /* [B, 2n, 3] (row, col, order), -1 padded: disks AND (type 0) PPuE */
    const double* ppue_points;   /* [B, 2n_ppue, 3] = prompts[0] for types 1/2 (is_vpu_model.py:396-397); NULL => points */
    int32_t n;                   /* clicks per half in `points` (1..24) */
    int32_t n_ppue;              /* clicks per half in `ppue_points` */
    int32_t type;                /* as_prompt_type: 0 clicks, 1 box, 2 scribble */
    const int32_t* boxes;        /* [B,5] (x_c, y_c, w, h, slot)                      type 1 */
    const int32_t* scrib_sel;    /* [B,2,img] selected offsets or INT32_MIN           type 2 */
    const int32_t* scrib_slot;   /* [B] PPuE row of the scribble or -1                type 2 */
    const uint8_t* extra_mask;   /* [B,2,img,img] box/scribble raster OR-ed into the click disks; may be NULL */
} vpu_prompts;

int vpu_forward(vpu_handle h, const float* image4 /* [B,4,img,img] fp32 */, const vpu_prompts* prompts, int B,
                float* instances /* [B,1,img,img] */, float* instances_aux /* [B,48,img,img] or NULL */,
                void* workspace, size_t workspace_bytes, void* stream);

/* ---- per-kernel-class device timing (bench.py roofline; measurement only) ----
 * Between vpu_profile_begin and vpu_profile_end every launch of vpu_forward is bracketed by CUDA events on the
 * caller's stream.  vpu_profile_end waits for the last event and returns one entry per kernel class
 * ("gemm.vit_window", "attn.dma", "ln.vit_global", ...): summed device ms, algorithmic FLOPs / bytes, launches. */
typedef struct vpu_profile_entry {
    char name[48];
    double ms;
    double flops;
    double bytes;
    int64_t launches;
} vpu_profile_entry;
int vpu_profile_begin(vpu_handle h);
int vpu_profile_end(vpu_handle h, vpu_profile_entry* out, int max_entries, int* n_out);

/* ---- stage entry points (also what the parity tests call) ---- */
/* PPuE rows (reference is_vpu_model.py:189-352, ops.py:39-325) -> out [B, 2*num_max_points, 2*img+3] fp32 */
int vpu_ppue(vpu_handle h, const vpu_prompts* prompts, int B, float* out, void* stream);
/* cat(prev_mask, disks|raster) (reference is_model.py:78-95, ops.py:347-382) -> out [B,3,img,img] fp32 */
int vpu_coord_features(vpu_handle h, const float* image4, const vpu_prompts* prompts, int B, float* out, void* stream);

/* ---- kernel-level entry points (unit parity tests; all operands bf16/fp32 device pointers) ---- */
/* out[M,N] = act(A[M,K] * W[N,K]^T + bias[N] + bias2d[m % rows, N] + residual[M,N]) */
int vpu_gemm(const void* A_bf16, int lda, const void* W_bf16, int ldw, int M, int N, int K, const float* bias,
             const float* bias2d, int bias2d_rows, const void* residual, int residual_dtype, int ldr, int act /*0,1 gelu,2 relu*/,
             void* out, int out_dtype, int ldo, int impl /*0 tcgen05 (2-CTA pairs when the shape allows), 1 mma.sync cross-check, 2 tcgen05 1-CTA*/, void* stream);
/* tail of the head at 1/4 resolution (reference swin_transformer.py:727-767): y_l = per-level fusion-conv slices, NHWC bf16
 * [B, res0 >> l, res0 >> l, 256].  f = relu(bias + y_0 + sum_l resize_bilinear(y_l)) (never stored);
 * seg_out[B, res0, res0] = <f, wseg> + seg_bias;  aux_out[B, nq, res0, res0] = (<f/|f|, qn[b, n]> + 1) / 2 (qn: unit rows,
 * bf16 [B, 64, 256]; aux_out may be NULL) */
int vpu_head_tail(const void* y0_bf16, const void* y1_bf16, const void* y2_bf16, const void* y3_bf16, int B, int res0,
                  const float* bias, const float* wseg, float seg_bias, const void* qn_bf16, int nq, float* seg_out,
                  float* aux_out, void* stream);
/* head pair of one pyramid level (reference swin_transformer.py:723-737), back to back with the [M,256] intermediate on chip:
 * out[M,256] = bf16( bf16(relu(A[M,K1] * W1[256,K1]^T + bias1)) * W2[256,256]^T ) */
int vpu_gemm_b2b(const void* A_bf16, int lda, const void* W1_bf16, const float* bias1, const void* W2_bf16, int M, int K1,
                 void* out_bf16, int ldo, void* stream);
/* image-side K|V|Q projection of the Dual-cross Merging Attention (reference transformer.py:444-449, 456-458): the positional term
 * key_pe W^T + b is a precomputed fp32 table of table_rows + table_pad_rows rows, the first table_pad_rows (>= 128) repeated after the
 * last one.  out[M,N] (bf16) = A[M,K] * W[N,K]^T + table[m % table_rows, N] */
int vpu_gemm_table(const void* A_bf16, int lda, const void* W_bf16, int ldw, int M, int N, int K, const float* table, int table_rows,
                   int table_pad_rows, void* out_bf16, int ldo, int impl, void* stream);
/* image -> tokens out-projection + residual + norm4 of a Dual-cross Merging Attention layer (reference transformer.py:459-463) in one
 * kernel: out[M,C] (bf16) = LayerNorm_C(A[M,K] * W[C,K]^T + bias + res[M,C]) * gamma + beta; C = 512 ... 1280, a multiple of 256.
 * rowmax_parts (may be NULL): fp32 [C / 256, M], the maximum of row m over each 256-column block of out (before bf16 rounding) */
int vpu_gemm_layernorm(const void* A_bf16, int lda, const void* W_bf16, int ldw, const float* bias, const void* res_bf16, int ldr,
                       const float* gamma, const float* beta, float eps, int M, int K, int C, void* out_bf16, int ldo,
                       float* rowmax_parts, void* stream);
/* ConvTranspose2d(k=2,s=2) as GEMM + pixel-shuffle store: A [B*g*g, K] -> out NHWC [B, 2g, 2g, cout] bf16 */
int vpu_gemm_pixel_shuffle(const void* A_bf16, const void* W_bf16, int M, int cout, int K, const float* bias4, int g,
                           void* out_bf16, int impl, void* stream);
/* softmax(q k^T * scale) v; window=0: contiguous sequences; window>0: 224-px window regrouping */
int vpu_attention(const void* q, int ldq, int qoff, const void* k, int ldk, int koff, const void* v, int ldv, int voff,
                  void* o, int ldo, int Sq, int Sk, int heads, int head_dim, int nprob, float scale, int window,
                  int grid, void* stream);
/* NoC evaluation protocol on the device (replaces isegm/inference/clicker.py:29-69 Clicker._get_next_click and
 * isegm/inference/utils.py:80-87 get_iou for S click sessions at once).
 *   gt          int8  [S,H,W]  1 object, 0 background, -1 ignore
 *   pred        uint8 [S,H,W]  thresholded prediction (0/1)
 *   not_clicked uint8 [S,H,W]  1 = not clicked yet; the chosen pixel is cleared
 *   clicks      int32 [S,4]    (is_positive, row, col, squared distance of the click from the error-region border)
 *   iou_counts  int64 [S,2]    (|pred & gt & keep|, |(pred | gt) & keep|): IoU = [0] / [1]
 * Bit-exact with the reference's cv2.distanceTransform(DIST_L2, 0) clicker: integer squared distances throughout. */
size_t vpu_noc_workspace_bytes(int S, int H, int W);
int vpu_noc_next_clicks(const int8_t* gt, const uint8_t* pred, uint8_t* not_clicked, int S, int H, int W, int32_t* clicks,
                        int64_t* iou_counts, void* workspace /* 256-byte aligned */, size_t workspace_bytes, void* stream);
/* Box / scribble outline planes on the device: replaces the host calls
 * cv2.rectangle(img, (x0,y0), (x1,y1), 255, 3) of draw_box and cv2.polylines(img, [curve], False, 255, 3) of draw_scribble
 * (isegm/model/is_model.py:97-146).  Bit-exact with OpenCV's fixed-point thick-line fill (vertices inside or outside the
 * image); the result is the `extra_mask` operand of vpu_prompts.
 *   type 1: boxes int32 [B,5] (x_c, y_c, w, h, slot) -> outline in plane 0 if slot < n else plane 1
 *   type 2: scribbles int32 [B,S,2] (x, y) -> open polyline in plane 0
 *   planes uint8 [B,2,size,size], cleared by the call */
int vpu_raster_prompts(int type, const int32_t* boxes, const int32_t* scribbles, int S, int n, int B, int size, uint8_t* planes,
                       void* stream);
/* Device-resident NoBRS click sessions: the predictor transforms either side of vpu_forward for S sessions at once
 * (SURVEY.md 8(f) rank 2).  Replaces, per click and per session, BasePredictor's input assembly and get_points_nd
 * (isegm/inference/predictors/base.py:106-151,195-213), ZoomIn.transform / _transform_clicks / inv_transform
 * (transforms/zoom_in.py:30-112), AddHorizontalFlip (transforms/flip.py:9-28) and SigmoidForPred (transforms/base.py:29-38)
 * in the configuration of scripts/evaluate_vpumodel.py:187-192 (skip_clicks = -1, fixed square target size, flip TTA).
 * All pointers are device memory owned by the caller; the state persists between clicks.  Initial state: prev_probs = 0,
 * pred = 0, nclicks = 0, roi = -1, fgbox = {INT32_MAX, -1, INT32_MAX, -1, -1}. */
typedef struct vpu_session_state {
    int32_t S, H, W;              /* sessions, full image size (equal for all sessions of the batch) */
    int32_t T;                    /* network input side = ZoomIn target_size (448) */
    int32_t max_clicks;           /* rows of the click table per session (<= 64) */
    int32_t n_half;               /* points per half written for the network (>= max_clicks, <= num_max_points) */
    const float* images;          /* [S,3,H,W] fp32 in [0,1] */
    float* prev_probs;            /* [S,H,W] last full-size probability map (prev_prediction / ZoomIn._prev_probs) */
    uint8_t* pred;                /* [S,H,W] prev_probs > pred_thr: the `pred` operand of vpu_noc_next_clicks */
    int32_t* clicks;              /* [S,max_clicks,3] (is_positive, row, col) in click order */
    int32_t* nclicks;             /* [S] */
    int32_t* roi;                 /* [S,4] zoom-in region rmin,rmax,cmin,cmax (inclusive); roi[0] < 0: none yet */
    int32_t* fgbox;               /* [S,5] bbox of prev_probs > zoom_thr and a state word (-1 no prediction yet, 0 empty, 1 set) */
    float pred_thr;               /* 0.49 (evaluate_vpumodel.py: --thresh) */
    float zoom_thr;               /* 0.5 (ZoomIn.prob_thresh) */
    double expansion_ratio;       /* 1.4 */
    double recompute_thresh_iou;  /* 0.5 */
    int32_t min_crop_size;        /* 200; < 0 = None */
} vpu_session_state;
/* Before the forward: for each a < A, session s = active[a] appends new_clicks[s] = (is_positive,row,col,.) (the row
 * vpu_noc_next_clicks wrote; NULL = no new click), updates its zoom-in region, and writes network rows a (crop) and A + a
 * (mirrored crop): net_image [2A,4,T,T] fp32 (RGB + previous probabilities), net_points [2A, 2*n_half, 3] float64. */
int vpu_session_prepare(const vpu_session_state* st, const int32_t* active, int A, const int32_t* new_clicks /* [S,4] or NULL */,
                        float* net_image, double* net_points, void* stream);
/* After the forward: logits [2A,1,T,T] (rows a and A + a) -> prev_probs, pred and fgbox of session active[a]. */
int vpu_session_finish(const vpu_session_state* st, const int32_t* active, int A, const float* logits, void* stream);
/* ToTensor of the predictor for a whole batch on the device (replaces transforms.ToTensor() of
 * isegm/inference/predictors/base.py:30,45 and the torch.cat with the previous mask of base.py:113-115): rgb_nhwc uint8 [B,H,W,3],
 * prev_mask fp32 [B,H,W] or NULL (zeros) -> image4 fp32 [B,4,H,W] = (rgb / 255 as an IEEE division, prev_mask): the operand of
 * vpu_forward, from 3 uploaded bytes per pixel instead of 12. */
int vpu_image_from_u8(const uint8_t* rgb_nhwc, const float* prev_mask, float* image4, int B, int H, int W, void* stream);
/* measurement only: CTA 0 of the following global-attention launches logs (event << 56 | clock64) per role into
 * dev_buf[4][cap] (uint64; roles: TMA thread, MMA thread, softmax warpgroup 0 / 1); NULL switches it off */
int vpu_debug_attention_trace(void* dev_buf, int cap);
int vpu_layernorm(const float* in, const float* gamma, const float* beta, float eps, int rows, int C, float* out_f32,
                  void* out_bf16, const float* pe, void* out_pe_bf16, float* rowmax, void* stream);
int vpu_groupnorm_nhwc(void* x_bf16, int B, int64_t per_sample, int C, const float* gamma, const float* beta, int gelu,
                       void* scratch /* >= B*8200 bytes */, void* stream);
int vpu_upsample_align_corners(const float* in, float* out, int h, int w, int H, int W, int64_t planes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VPUFORMER_B200_H */
