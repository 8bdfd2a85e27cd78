/* mga_cbam.h -- C ABI of the B200-native mask-guided CBAM hot path.
 *
 * The reference (MarioPasc/MGA-YOLO) has no native code and no FFI for this path: it is
 * ~200 eager ATen calls per forward.  These entry points are what a Python FFI (ctypes /
 * torch custom op) for that path binds, one per reference function group:
 *
 *   mga_cbam_forward    replaces MaskCBAM.forward              mga_yolo/nn/modules/masked_cbam.py:154-171
 *                        (= _masked_avg :87-102, _masked_max :104-121, _cam :123-130,
 *                           _sam :132-148, alpha residual :150-152,166-171, and the
 *                           deterministic ProbMaskGater clamp  probmaskgater.py:77,82-83)
 *   mga_cbam_backward   replaces what torch autograd derives from those lines
 *                        (closed form: SURVEY.md section 8a "Backward")
 *   mga_mask_downsample replaces MaskUtils.downsample_mask      mga_yolo/utils/mask_utils.py:64-141
 *                        and MaskUtils.downsample_mask_prob     mga_yolo/utils/mask_utils.py:14-48
 *                        (nearest / area / maxpool / default "maxpool+close" / avgpool)
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch caching allocator);
 *     the library never allocates, frees or synchronises.
 *   - tensors are dense NCHW; `stream` is a cudaStream_t passed as void*.
 *   - return value: 0 = ok, otherwise an MGA_ERR_* code; mga_last_error() gives text.
 *     Nothing throws.  There is no CPU implementation behind any entry point.
 */
#ifndef MGA_CBAM_H
#define MGA_CBAM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGA_ABI_VERSION 3

enum { MGA_OK = 0, MGA_ERR_ARG = 1, MGA_ERR_UNSUPPORTED = 2, MGA_ERR_CUDA = 3, MGA_ERR_WORKSPACE = 4 };

/* element types of feature map / gradient / mask tensors (parameters are always fp32) */
enum { MGA_F32 = 0, MGA_BF16 = 1, MGA_F16 = 2, MGA_U8 = 3 };

/* mga_cbam_desc.flags */
enum {
    MGA_HAS_MASK = 1 << 0,         /* mask pointer is valid; otherwise vanilla CBAM (masked_cbam.py:90-91,107-108,137-138) */
    MGA_SIGMOID_MASK = 1 << 1,     /* use_sigmoid_mask (masked_cbam.py:93-94,110-111,143-144) */
    MGA_GATE_CLAMP = 1 << 2,       /* eval-mode ProbMaskGater: clamp mask to [0,1] first (probmaskgater.py:77) */
    MGA_SAMCAM_ADD = 1 << 4,       /* sam_cam_fusion = add (build-side mode, parity unpinned); default multiply = reference */
    MGA_PYRAMID_MULTIPLY = 1 << 6, /* mga_pyramid_fusion = multiply (build-side mode); default add = reference alpha-skip */
    MGA_FORCE_SPLIT = 1 << 8,      /* one kernel per phase instead of the cluster-per-sample kernels (bits 9 and 11 are retired) */
    MGA_GATES_ONLY = 1 << 10,      /* internal: compute / differentiate the two gates s(B,C), a(B,HW) only (concat fusion modes) */
    MGA_GATES_ACC = 1 << 16,       /* internal: gates backward adds an upstream feature gradient (mga_cbam_gates_backward_acc) */
    MGA_NO_PERSIST = 1 << 14,      /* tuning library only: skip the persistent shared-memory-resident experiment even when MGA_PF=1 */
    MGA_NO_SAVE = 1 << 12          /* inference (model.eval() + no_grad, predictor.py:7-24): the forward may skip the saved-for-backward
                                      planes; ctx is then NOT valid for mga_cbam_backward (s and a are still written) */
};

typedef struct mga_cbam_desc {
    int32_t B, C, H, W; /* feature map (B,C,H,W) */
    int32_t hidden;     /* max(1, C / r)            masked_cbam.py:53 */
    int32_t ksize;      /* odd spatial kernel, <= 7 masked_cbam.py:47,61 */
    int32_t dtype;      /* MGA_F32 / MGA_BF16 / MGA_F16: x, out, grad_out, grad_x */
    int32_t mask_dtype; /* MGA_F32 / MGA_BF16 / MGA_F16: mask and grad_mask */
    int32_t flags;
    float tiny_mask_thr; /* masked_cbam.py:40,98 */
    float eps;           /* masked_cbam.py:41,99 */
} mga_cbam_desc;

/* parameters, named after the reference state_dict (masked_cbam.py:54-64); all fp32 */
typedef struct mga_cbam_params {
    const float* w1;   /* cam_mlp.0.weight (hidden, C) */
    const float* b1;   /* cam_mlp.0.bias   (hidden)    */
    const float* w2;   /* cam_mlp.2.weight (C, hidden) */
    const float* b2;   /* cam_mlp.2.bias   (C)         */
    const float* wsam; /* sam_conv.weight  (1,3,k,k)   */
    const float* beta; /* beta ()                      */
} mga_cbam_params;

typedef struct mga_cbam_grads {
    float* w1;
    float* b1;
    float* w2;
    float* b2;
    float* wsam;
    float* beta;
} mga_cbam_grads;

int mga_abi_version(void);
const char* mga_last_error(void);

/* launch accounting (not part of the reference surface; used by bench.py):
 *   mga_launch_count  - kernels launched by this library in this process so far
 *   mga_profile_*     - when enabled, every launch is bracketed by CUDA events on its stream;
 *                       read (name, milliseconds) per launch after synchronising.  Not usable
 *                       while a CUDA graph is being captured. */
unsigned long long mga_launch_count(void);
int mga_profile_enable(int on);
int mga_profile_count(void);
int mga_profile_read(int index, const char** name, float* ms);

/* Which launch path mga_cbam_forward (direction 0) / mga_cbam_backward (direction 1) takes for this descriptor, and its geometry.
 * Host-only (no CUDA call): the choice depends on the per-sample shape and element type, never on B.  The device may still
 * send a cluster launch to the per-phase path when no cluster of that shape can be resident (not the case on B200). */
typedef struct mga_cbam_plan_info {
    int32_t path;          /* 0 = one kernel per phase, 1 = one thread-block cluster per sample */
    int32_t cluster_size;  /* CTAs per sample (path 1) */
    int32_t rows_per_cta;  /* image rows per CTA (path 1) */
    int32_t threads;       /* threads per CTA (path 1) */
    int32_t smem_bytes;    /* dynamic shared memory per CTA (path 1) */
    int32_t launches;      /* kernel launches of the call */
} mga_cbam_plan_info;
int mga_cbam_plan(const mga_cbam_desc* d, int direction, mga_cbam_plan_info* info);

/* bytes of the saved-for-backward context and of the transient scratch for this shape */
int mga_cbam_workspace(const mga_cbam_desc* d, size_t* ctx_bytes, size_t* scratch_bytes);

/* out = MaskCBAM([x, mask]); fills ctx (kept by the caller until backward). */
int mga_cbam_forward(const mga_cbam_desc* d, const void* x, const void* mask, const mga_cbam_params* p,
                     void* out, void* ctx, void* scratch, void* stream);

/* grad_x, grad_mask (may be NULL), parameter gradients (overwritten, not accumulated). */
int mga_cbam_backward(const mga_cbam_desc* d, const void* x, const void* mask, const void* grad_out,
                      const mga_cbam_params* p, const void* ctx, void* grad_x, void* grad_mask,
                      const mga_cbam_grads* gp, void* scratch, void* stream);

/* Gates only (building block of the `concat` fusion modes, whose 1x1 convolutions are library GEMMs on the host side):
 * forward fills ctx with s = channel gate (B,C) and a = spatial gate computed from x (B,HW) -- read them with
 * mga_cbam_ctx_view; backward takes dL/ds (B,C) and dL/da (B,HW), both fp32, and returns grad_x, grad_mask and the
 * gradients of w1,b1,w2,b2,wsam (gp->beta is written with an unspecified value). */
int mga_cbam_gates_forward(const mga_cbam_desc* d, const void* x, const void* mask, const mga_cbam_params* p, void* ctx, void* scratch,
                           void* stream);
int mga_cbam_gates_backward(const mga_cbam_desc* d, const void* x, const void* mask, const float* grad_s, const float* grad_a,
                            const mga_cbam_params* p, const void* ctx, void* grad_x, void* grad_mask, const mga_cbam_grads* gp,
                            void* scratch, void* stream);
/* Same, with an upstream feature gradient: grad_x = grad_x_acc + (gate gradients pulled back to x).  grad_x_acc (B,C,H,W, feature dtype)
 * is what another consumer of x already produced (the concat kernels' dx); the sum costs no extra pass (the kernel that writes grad_x reads
 * it in place of grad_out).  grad_x_acc may be NULL (= mga_cbam_gates_backward); it must not alias grad_x. */
int mga_cbam_gates_backward_acc(const mga_cbam_desc* d, const void* x, const void* mask, const float* grad_s, const float* grad_a,
                                const void* grad_x_acc, const mga_cbam_params* p, const void* ctx, void* grad_x, void* grad_mask,
                                const mga_cbam_grads* gp, void* scratch, void* stream);

/* Fused forward of sam_cam_fusion = concat (build-side mode, parity unpinned) on the tcgen05 tensor cores, 16-bit features:
 *   out = k0 * x + k1 * (Wa (x * s) + Wb (x * a) + bias),   W = [Wa | Wb] (C, 2C) fp32 = fuse_sam_cam.weight, alpha = softplus(*beta),
 *   (k0, k1) = (1 - alpha, alpha), or (0, alpha) with MGA_PYRAMID_MULTIPLY in d->flags.
 * s (B,C) and a (B,H*W) are the fp32 gates of mga_cbam_gates_forward.  wscratch: (B + 1) * C * C elements of the feature dtype.
 * Needs C % 128 == 0 and H*W % 8 == 0 (MGA_ERR_UNSUPPORTED otherwise: compose the mode from the gates op and a library GEMM). */
int mga_cbam_concat_forward(const mga_cbam_desc* d, const void* x, const float* s, const float* a, const float* w, const float* bias,
                            const float* beta, void* out, void* wscratch, void* stream);

/* Elementwise / reduction part of the backward of mga_cbam_concat_forward.  The caller provides uv (B, 2C, H*W) = [Wa^T g ; Wb^T g]
 * (one library GEMM on the raw upstream gradient g = grad_out).  Writes grad_x = alpha (s U + a V) + k0 g, ga = g * a (left operand of
 * the dWb GEMM), grad_a (B, H*W) = alpha sum_c x V, and per-tile partial rows: ds_part / dbias_part (B, nTiles, C) (already scaled by
 * alpha; nTiles = ceil(H*W / 256)) and dalpha_part (B, nTiles) (sum over everything = d out / d alpha contracted with g). */
int mga_cbam_concat_backward_elem(const mga_cbam_desc* d, const void* x, const void* grad_out, const void* uv, const float* s, const float* a,
                                  const float* bias, const float* beta, void* grad_x, void* ga, float* ds_part, float* dbias_part, float* grad_a,
                                  float* dalpha_part, void* stream);

/* Backward of mga_cbam_concat_forward with respect to the features and the gates, on the tcgen05 tensor cores: U = Wa^T g and V = Wb^T g
 * are the two accumulators of ONE kernel whose epilogue finishes the closed form (U, V never reach memory):
 *   grad_x = alpha (s_c U + a_p V) + k0 grad_out        ga = grad_out * a_p   (B,C,H*W, feature dtype; left operand of the dWb GEMM)
 *   ds_part, dbias_part (B, 2 nT, C)  per-(128-pixel tile, half) partial sums over pixels:   ds = sum over dim 1, dbias = sum over dims 0, 1
 *   da_part (B, C / 32, H*W)          per-32-channel partial sums:                            grad_a = sum over dim 1
 *   dalpha_part (B, nT, C / 16)       partial sums of d out / d alpha:                        d beta = sigmoid(beta) * sum
 * with nT = ceil(H*W / 128); every partial buffer is written (never accumulated): deterministic.  wscratch: 2 * C * C elements of the
 * feature dtype (the transposed weights).  Same shape / dtype conditions as mga_cbam_concat_forward. */
int mga_cbam_concat_backward_dx(const mga_cbam_desc* d, const void* x, const void* grad_out, const float* s, const float* a, const float* w,
                                const float* bias, const float* beta, void* grad_x, void* ga, float* ds_part, float* dbias_part, float* da_part,
                                float* dalpha_part, void* wscratch, void* stream);

/* Weight gradient of mga_cbam_concat_forward from the per-sample GEMM results Ga = g X^T and Gb = (g * a) X^T, both (B, C, C) of gemm_dtype
 * (MGA_F32 / MGA_BF16 / MGA_F16): grad_w (C, 2C) fp32 = alpha * [ sum_b Ga[b] diag(s_b) | sum_b Gb[b] ], fixed summation order. */
int mga_cbam_concat_wgrad_reduce(const mga_cbam_desc* d, const void* Ga, const void* Gb, int gemm_dtype, const float* s, const float* beta,
                                 float* grad_w, void* stream);

/* read-back of small saved quantities for tests / logging: which = 0 s(B,C), 1 a(B,HW) */
int mga_cbam_ctx_view(const mga_cbam_desc* d, const void* ctx, int which, const float** ptr, size_t* count);

/* mask downsample methods (mask_utils.py:64-141 and :14-48) */
enum {
    MGA_DS_NEAREST = 0,       /* cv2.INTER_NEAREST */
    MGA_DS_AREA = 1,          /* cv2.INTER_AREA on uint8, then > thresh (binary) */
    MGA_DS_MAXPOOL = 2,       /* zero pad + block max */
    MGA_DS_AVGPOOL = 3,       /* zero pad + block mean (float out) */
    MGA_DS_AREA_RAW = 4       /* cv2.INTER_AREA on uint8, no threshold (downsample_mask_prob 'area') */
};

/* src: (B,H,W) uint8 {0,1}; dst: (B,ceil(H/s),ceil(W/s)) uint8 or float32 (out_dtype).
 * close3x3 != 0 applies the 3x3 MORPH_CLOSE "bridge"; tmp must hold 2*B*nh*nw bytes then. */
int mga_mask_downsample(const uint8_t* src, void* dst, void* tmp, int32_t B, int32_t H, int32_t W, int32_t stride,
                        int32_t method, float thresh, int32_t close3x3, int32_t out_dtype, void* stream);

/* All three pyramid masks of MGADataset.__getitem__ (mga_yolo/data/dataset.py:95-103: `for s in (8, 16, 32): downsample_mask[_prob]`)
 * in ONE pass over the (B,H,W) uint8 {0,1} masks: dst8 (B,H/8,W/8), dst16 (B,H/16,W/16), dst32 (B,H/32,W/32), uint8 or float32.
 * Same methods / thresh / close3x3 as mga_mask_downsample, bit-identical results.  Needs H % 32 == 0 and W % 32 == 0
 * (letterboxed inputs); returns MGA_ERR_UNSUPPORTED otherwise (call mga_mask_downsample per stride). */
int mga_masks_multi(const uint8_t* src, void* dst8, void* dst16, void* dst32, int32_t B, int32_t H, int32_t W, int32_t method,
                    float thresh, int32_t close3x3, int32_t out_dtype, void* stream);

/* Same results in two stages when the caller provides tmp (2 * B * (H/8) * (W/8) bytes): stage 1 reads the masks with one thread per
 * 8x8 block over the whole batch (enough CTAs to pull them at HBM speed), stage 2 derives the three strides per image from the block
 * counts.  tmp == NULL is mga_masks_multi. */
int mga_masks_multi_ws(const uint8_t* src, void* dst8, void* dst16, void* dst32, void* tmp, int32_t B, int32_t H, int32_t W,
                       int32_t method, float thresh, int32_t close3x3, int32_t out_dtype, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Components either side of the block
 * ------------------------------------------------------------------------------------------------------------------ */

/* MaskECA (mga_yolo/nn/modules/masked_eca.py:139-193): out = x * (1 + softplus(beta) * (sigmoid(conv1d_k(masked_avg(x, mask))) - 0.5)).
 * The descriptor is mga_cbam_desc with `hidden` = the odd conv1d kernel size k (masked_eca.py:43-52, <= 15) and `ksize` ignored;
 * flags: MGA_HAS_MASK, MGA_SIGMOID_MASK.  w1d = conv1d.weight (k floats), beta ().  Gradients are written, not accumulated. */
int mga_eca_workspace(const mga_cbam_desc* d, size_t* ctx_bytes, size_t* scratch_bytes);
int mga_eca_forward(const mga_cbam_desc* d, const void* x, const void* mask, const float* w1d, const float* beta, void* out, void* ctx,
                    void* scratch, void* stream);
int mga_eca_backward(const mga_cbam_desc* d, const void* x, const void* mask, const void* grad_out, const float* w1d, const void* ctx,
                     void* grad_x, void* grad_mask, float* grad_w1d, float* grad_beta, void* scratch, void* stream);

/* MGAMaskHead tail (mga_yolo/nn/modules/segmentation.py:94,107-110): logits (B,1,H,W) fp32 = Conv2d(C, 1, 3, padding 1, bias)(feat),
 * feat (B,C,H,W) of `dtype`, weight (1,C,3,3) and bias (1) fp32, C <= 256.  Backward: grad_feat (dtype of feat), grad_weight, grad_bias. */
int mga_head_tail_forward(const void* feat, const float* weight, const float* bias, float* logits, int32_t B, int32_t C, int32_t H, int32_t W,
                          int32_t dtype, void* stream);
int mga_head_tail_backward(const void* feat, const float* weight, const float* grad_logits, void* grad_feat, float* grad_weight, float* grad_bias,
                           int32_t B, int32_t C, int32_t H, int32_t W, int32_t dtype, void* stream);

/* ProbMaskGater in train mode (mga_yolo/nn/modules/probmaskgater.py:59-95).  mode: 0 gumbel, 1 hard_st, 2 bernoulli_detach.
 * p: n raw gate inputs (clamped to [0,1], floored at p_min inside); out: the gate; soft: the soft gate saved for backward (modes 0/1).
 * NOISE CONTRACT: when `noise` is NULL, element i draws Philox4x32-10 (Salmon et al. 2011) with key = (seed & 0xffffffff, seed >> 32)
 * and counter = (i & 0xffffffff, i >> 32, offset & 0xffffffff, offset >> 32); u1 = (r0 + 0.5) * 2^-32, u2 = (r1 + 0.5) * 2^-32 (fp32).
 * gumbel / hard_st use logistic noise -log(-log u1) + log(-log u2) (u clamped to [1e-6, 1-1e-6]); bernoulli_detach is [u1 < p].
 * The caller advances `offset` by one per call (the reference re-seeds a generator with seed + call counter, probmaskgater.py:43-49).
 * When `noise` is not NULL it holds the uniforms (u1: noise[0..n), u2: noise[n..2n)) -- how the parity tests feed the reference's own
 * draws.  noise_out (2n floats, may be NULL) receives the uniforms that were used. */
int mga_gate_sample_forward(const float* p, const float* noise, float* out, float* soft, float* noise_out, size_t n, int32_t mode, float tau,
                            float p_min, float threshold, uint64_t seed, uint64_t offset, void* stream);
int mga_gate_sample_backward(const float* grad_out, const float* p, const float* soft, float* grad_p, size_t n, float tau, float p_min, void* stream);

/* Zero-pad + stack of per-sample masks of ONE pyramid stride (mga_yolo/data/dataset.py:149-169): items is a DEVICE array of B records
 * { const void* src; int32_t h, w; } (16 bytes each), src uint8 or float32 (h,w) maps; dst (B,1,H,W) float32 with H >= h_i, W >= w_i. */
int mga_collate_masks(const void* items_dev, float* dst, int32_t B, int32_t H, int32_t W, int32_t src_dtype, void* stream);

/* MaskSPADE, feature side (mga_yolo/nn/modules/masked_spade.py:73,126-143): out = gamma * IN(x) + beta with IN = affine-free
 * InstanceNorm2d (per (sample, channel) mean / biased variance over H*W, rstd = 1/sqrt(var + eps)).  x, out, grad_out, grad_x: (B,C,H,W)
 * of `dtype`; gamma, beta, grad_gamma: (B,C,H,W) of `mod_dtype` = `dtype` or MGA_F32 (the outputs of the block's mask branch
 * conv_gamma / conv_beta, which stay with the caller: dense 3x3 convolutions); gamma == beta == NULL is the mask-less branch
 * (plain normalisation, masked_spade.py:128-130).  stats: (B,C,2) fp32 = mean, rstd, written by forward and read by backward.
 * Backward: grad_gamma = g * xhat (written when not NULL), grad_beta = g (the caller's tensor, nothing to compute),
 * grad_x = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)) with dxhat = g * gamma (g when gamma == NULL). */
int mga_spade_forward(const void* x, const void* gamma, const void* beta, void* out, float* stats, int32_t B, int32_t C, int32_t H, int32_t W, float eps,
                      int32_t dtype, int32_t mod_dtype, void* stream);
int mga_spade_backward(const void* x, const void* grad_out, const void* gamma, const float* stats, void* grad_x, void* grad_gamma, int32_t B, int32_t C,
                       int32_t H, int32_t W, int32_t dtype, int32_t mod_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MGA_CBAM_H */
