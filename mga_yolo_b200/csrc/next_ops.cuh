// next_ops.cuh -- (included at the end of mga_cbam.cu: shares the pooling kernels of cbam_fwd.cuh) the components either side of the mask-guided CBAM block, C ABI in include/mga_cbam.h:
//
//   mga_eca_*          MaskECA: masked average pool -> conv1d over channels -> sigmoid -> x * (1 + alpha (w - 0.5))
//                      /root/reference/mga_yolo/nn/modules/masked_eca.py:139-193 (shares the masked-pool front end of the CBAM block:
//                      the same mask_prep / cam_pool kernels)
//   mga_head_tail_*    MGAMaskHead.head: Conv2d(hidden, 1, 3, padding 1, bias) producing the mask logits, and its backward
//                      /root/reference/mga_yolo/nn/modules/segmentation.py:94,107-110
//   mga_gate_sample_*  ProbMaskGater in train mode (gumbel / hard_st / bernoulli_detach) with a stated Philox4x32-10 noise contract
//                      /root/reference/mga_yolo/nn/modules/probmaskgater.py:59-95
//   mga_collate_masks  zero-pad + stack of per-sample pyramid masks
//                      /root/reference/mga_yolo/data/dataset.py:149-169
// All of them are HBM-bound streaming kernels: 128-bit accesses along the contiguous pixel axis, fp32 accumulation, deterministic
// (no atomics), nothing allocated, everything on the caller's stream.
#pragma once
#include <algorithm>
#include <cstring>

#include "cbam_fwd.cuh"
#include "common.cuh"

namespace mga {

// ================================================================== MaskECA
struct EcaCtx {     // saved for backward (carved from the caller's ctx buffer)
    float* consts;  // [0]=alpha [1]=sigmoid(beta)
    float* m;       // (B,S)
    float* y;       // (B,C) pooled descriptor after the GAP blend
    float* A;       // (B,C) masked average before the blend
    float* w;       // (B,C) sigmoid(conv1d(y))
    float* gate;    // (B,C) 1 + alpha (w - 0.5)
    float* use;     // (B)
    float* den;     // (B)
    float* msum;    // (B)
};
struct EcaScratch {
    FwdScratch fs;   // pooling partials (sxm, sx, best, bidx, mpart)
    float* dg;       // (B,C) sum_p gout x
    float* cA;       // (B,C)
    float* cG;       // (B,C)
    float* kb;       // (B)
    float* wpart;    // (B, kmax + 1): per-sample conv1d-weight gradient taps and d alpha
};
constexpr int kEcaMaxK = 15;

struct Carve2 {
    char* base;
    size_t off = 0;
    template <typename T> T* take(size_t n) {
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += align256(n * sizeof(T));
        return p;
    }
};
static size_t eca_carve_ctx(const Shape& s, void* base, EcaCtx* c) {
    Carve2 k{static_cast<char*>(base)};
    const size_t BS = (size_t)s.B * s.S, BC = (size_t)s.B * s.C;
    c->consts = k.take<float>(8);
    c->m = k.take<float>(BS);
    c->y = k.take<float>(BC);
    c->A = k.take<float>(BC);
    c->w = k.take<float>(BC);
    c->gate = k.take<float>(BC);
    c->use = k.take<float>(s.B);
    c->den = k.take<float>(s.B);
    c->msum = k.take<float>(s.B);
    return k.off;
}
static size_t eca_carve_scratch(const Shape& s, void* base, EcaScratch* e) {
    Carve2 k{static_cast<char*>(base)};
    const size_t BC = (size_t)s.B * s.C;
    e->fs.sxm = k.take<float>(BC);
    e->fs.sx = k.take<float>(BC);
    e->fs.best = k.take<float>(BC);
    e->fs.bidx = k.take<int>(BC);
    e->fs.mpart = k.take<float>((size_t)s.B * ((s.S + kMaskTile - 1) / kMaskTile));
    e->dg = k.take<float>(BC);
    e->cA = k.take<float>(BC);
    e->cG = k.take<float>(BC);
    e->kb = k.take<float>(s.B);
    e->wpart = k.take<float>((size_t)s.B * (kEcaMaxK + 1));
    return k.off;
}

// one CTA per sample: blend, conv1d over the channel axis (zero padded), sigmoid, gate
__global__ void __launch_bounds__(kBlock) eca_gate_kernel(Shape sh, const float* __restrict__ w1d, const float* __restrict__ beta, EcaCtx ctx,
                                                          FwdScratch fs, int nMaskTiles) {
    extern __shared__ float esm[];  // y (C)
    const int C = sh.C, b = blockIdx.x, k = sh.hidden, pad = k / 2;
    const bool has_mask = sh.has_mask();
    float use = 0.0f, den = 1.0f, tot = 0.0f;
    if (has_mask) {
        for (int t = 0; t < nMaskTiles; ++t) tot += __ldcg(fs.mpart + (size_t)b * nMaskTiles + t);  // same order in every thread
        use = (tot / (float)sh.S >= sh.tiny_thr) ? 1.0f : 0.0f;
        den = fmaxf(tot, sh.eps);
    }
    const float invS = 1.0f / (float)sh.S;
    for (int c = threadIdx.x; c < C; c += kBlock) {
        const int i = b * C + c;
        const float G = __ldcg(fs.sx + i) * invS;
        const float A = has_mask ? __ldcg(fs.sxm + i) / den : G;
        const float y = has_mask ? (A * use + G * (1.0f - use)) : G;
        esm[c] = y;
        ctx.y[i] = y;
        ctx.A[i] = A;
    }
    const float bt = __ldg(beta);
    const float alpha = softplusf_acc(bt);
    if (threadIdx.x == 0) {
        ctx.use[b] = use; ctx.den[b] = den; ctx.msum[b] = tot;
        if (b == 0) { ctx.consts[0] = alpha; ctx.consts[1] = sigmoidf_acc(bt); }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kBlock) {
        float z = 0.0f;
        for (int j = 0; j < k; ++j) {
            const int cc = c + j - pad;
            if (cc >= 0 && cc < C) z = fmaf(__ldg(w1d + j), esm[cc], z);
        }
        const float w = sigmoidf_acc(z);
        ctx.w[b * C + c] = w;
        ctx.gate[b * C + c] = fmaf(alpha, w - 0.5f, 1.0f);
    }
}

// out = x * gate[b][c]: grid-stride over 16-byte units
template <typename T, int VEC>
__global__ void __launch_bounds__(kBlock) eca_rescale_kernel(const T* __restrict__ x, T* __restrict__ out, Shape sh, EcaCtx ctx) {
    const int U = sh.S / VEC;
    const size_t total = (size_t)sh.B * sh.C * U;
    for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < total; i += (size_t)gridDim.x * kBlock) {
        const int pl = (int)(i / U);
        float v[VEC];
        ldv<T, VEC, kLdStream>(x + i * VEC, v);
        const float g = __ldg(ctx.gate + pl);
#pragma unroll
        for (int e = 0; e < VEC; ++e) v[e] *= g;
        stv<T, VEC, true>(out + i * VEC, v);
    }
}

// per (b,c) plane: dg = sum_p gout * x   (one warp per plane, 8 planes per CTA)
template <typename T, int VEC>
__global__ void __launch_bounds__(kBlock) eca_bwd_reduce_kernel(const T* __restrict__ x, const T* __restrict__ g, Shape sh, EcaScratch es) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int pl = blockIdx.x * kWarpsPerBlock + w;
    if (pl >= sh.B * sh.C) return;
    const int U = sh.S / VEC;
    const T* xp = x + (size_t)pl * sh.S;
    const T* gp = g + (size_t)pl * sh.S;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int u0 = lane; u0 < U; u0 += 4 * 32) {
        float xv[4][VEC], gv[4][VEC];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (u0 + q * 32 < U) {
                ldv<T, VEC, kLdStream>(xp + (size_t)(u0 + q * 32) * VEC, xv[q]);
                ldv<T, VEC, kLdStream>(gp + (size_t)(u0 + q * 32) * VEC, gv[q]);
            }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (u0 + q * 32 < U) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) acc[q] = fmaf(xv[q][e], gv[q][e], acc[q]);
            }
    }
    const float t = warp_sum((acc[0] + acc[1]) + (acc[2] + acc[3]));
    if (lane == 0) es.dg[pl] = t;
}

// one CTA per sample: d alpha, dz, conv1d-weight taps, dy (transposed conv), pooling coefficients
__global__ void __launch_bounds__(kBlock) eca_bwd_gate_kernel(Shape sh, const float* __restrict__ w1d, EcaCtx ctx, EcaScratch es) {
    extern __shared__ float esm[];  // dz (C) | y (C)
    __shared__ float red[32];
    const int C = sh.C, b = blockIdx.x, k = sh.hidden, pad = k / 2;
    float* dz = esm;
    float* ys = esm + C;
    const float alpha = ctx.consts[0];
    float da = 0.0f;
    for (int c = threadIdx.x; c < C; c += kBlock) {
        const int i = b * C + c;
        const float w = ctx.w[i], dg = es.dg[i];
        dz[c] = alpha * dg * w * (1.0f - w);
        ys[c] = ctx.y[i];
        da = fmaf(dg, w - 0.5f, da);
    }
    const float da_tot = block_sum(da, red);
    if (threadIdx.x == 0) es.wpart[(size_t)b * (kEcaMaxK + 1) + kEcaMaxK] = da_tot;
    // taps: dW[j] = sum_c dz[c] * y[c + j - pad]   (one warp per tap)
    {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        for (int j = w; j < k; j += kWarpsPerBlock) {
            float t = 0.0f;
            for (int c = lane; c < C; c += 32) {
                const int cc = c + j - pad;
                if (cc >= 0 && cc < C) t = fmaf(dz[c], ys[cc], t);
            }
            t = warp_sum(t);
            if (lane == 0) es.wpart[(size_t)b * (kEcaMaxK + 1) + j] = t;
        }
    }
    const bool has_mask = sh.has_mask();
    const float use = has_mask ? ctx.use[b] : 0.0f, den = has_mask ? ctx.den[b] : 1.0f;
    const float pass = (has_mask && ctx.msum[b] >= sh.eps) ? 1.0f : 0.0f;
    const float invS = 1.0f / (float)sh.S;
    float kacc = 0.0f;
    for (int c = threadIdx.x; c < C; c += kBlock) {
        float dy = 0.0f;
        for (int j = 0; j < k; ++j) {  // y[c] feeds z[c - j + pad] with weight w1d[j]
            const int cc = c - j + pad;
            if (cc >= 0 && cc < C) dy = fmaf(__ldg(w1d + j), dz[cc], dy);
        }
        const int i = b * C + c;
        const float cA = has_mask ? use * dy / den : 0.0f;
        es.cA[i] = cA;
        es.cG[i] = (has_mask ? (1.0f - use) : 1.0f) * dy * invS;
        kacc = fmaf(cA, ctx.A[i] * pass, kacc);
    }
    const float kb = block_sum(kacc, red);
    if (threadIdx.x == 0) es.kb[b] = kb;
}

// grid (pixel tiles of 32 units, B): dx = gout * gate_c + cA_c * m_p + cG_c ; dmask = (sum_c cA_c x - kb) [* m (1 - m)]
template <typename T, int VEC>
__global__ void __launch_bounds__(kBlock) eca_bwd_dx_kernel(const T* __restrict__ x, const T* __restrict__ g, const void* __restrict__ mask, int mdt,
                                                            T* __restrict__ dx, void* __restrict__ dmask, Shape sh, EcaCtx ctx, EcaScratch es) {
    __shared__ float part[kWarpsPerBlock][32 * VEC];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.y, U = sh.S / VEC, C = sh.C;
    const int u = blockIdx.x * 32 + lane;
    const bool act = u < U;
    const bool has_mask = sh.has_mask();
    float mv[VEC], racc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) { mv[e] = 0.0f; racc[e] = 0.0f; }
    if (act && has_mask) ldf<VEC>(ctx.m + (size_t)b * sh.S + (size_t)u * VEC, mv);
    if (act) {
        for (int c = w; c < C; c += kWarpsPerBlock) {
            const size_t o = ((size_t)b * C + c) * sh.S + (size_t)u * VEC;
            float xv[VEC], gv[VEC], ov[VEC];
            ldv<T, VEC, kLdStream>(x + o, xv);
            ldv<T, VEC, kLdStream>(g + o, gv);
            const float gate = __ldg(ctx.gate + b * C + c), cA = __ldcg(es.cA + b * C + c), cG = __ldcg(es.cG + b * C + c);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                ov[e] = fmaf(gv[e], gate, fmaf(cA, mv[e], cG));
                racc[e] = fmaf(cA, xv[e], racc[e]);
            }
            stv<T, VEC, true>(dx + o, ov);
        }
    }
    if (!has_mask || dmask == nullptr) return;
#pragma unroll
    for (int e = 0; e < VEC; ++e) part[w][lane * VEC + e] = racc[e];
    __syncthreads();
    const float kb = es.kb[b];
    for (int i = threadIdx.x; i < 32 * VEC; i += kBlock) {
        const int p = blockIdx.x * 32 * VEC + i;
        if (p >= sh.S) continue;
        float r = 0.0f;
#pragma unroll
        for (int q = 0; q < kWarpsPerBlock; ++q) r += part[q][i];
        float dm = r - kb;
        if (sh.sigmoid_mask()) {
            const float m = ctx.m[(size_t)b * sh.S + p];
            dm *= m * (1.0f - m);
        }
        store_mask_any(dmask, mdt, (size_t)b * sh.S + p, dm);
    }
}

// batch sums of the per-sample parameter-gradient pieces (fp64 accumulation, fixed order)
__global__ void __launch_bounds__(32) eca_wgrad_kernel(Shape sh, EcaCtx ctx, EcaScratch es, float* __restrict__ dw1d, float* __restrict__ dbeta) {
    const int j = blockIdx.x;  // tap, or k for d beta
    const int k = sh.hidden;
    double acc = 0.0;
    for (int b = threadIdx.x; b < sh.B; b += 32) acc += (double)es.wpart[(size_t)b * (kEcaMaxK + 1) + (j < k ? j : kEcaMaxK)];
    acc = warp_sum_d(acc);
    if (threadIdx.x == 0) {
        if (j < k) dw1d[j] = (float)acc;
        else dbeta[0] = (float)((double)ctx.consts[1] * acc);  // d beta = sigmoid(beta) * d alpha
    }
}

static int eca_validate(const mga_cbam_desc* d, Shape* sh) {
    if (!d) return fail(MGA_ERR_ARG, "null descriptor");
    if (d->B <= 0 || d->C <= 0 || d->H <= 0 || d->W <= 0) return fail(MGA_ERR_ARG, "bad shape (%d,%d,%d,%d)", d->B, d->C, d->H, d->W);
    if (d->hidden < 1 || d->hidden > kEcaMaxK || d->hidden % 2 == 0) return fail(MGA_ERR_UNSUPPORTED, "ECA kernel size %d: odd sizes up to %d", d->hidden, kEcaMaxK);
    if (d->dtype != MGA_F32 && d->dtype != MGA_BF16 && d->dtype != MGA_F16) return fail(MGA_ERR_ARG, "bad feature dtype %d", d->dtype);
    if (d->mask_dtype != MGA_F32 && d->mask_dtype != MGA_BF16 && d->mask_dtype != MGA_F16) return fail(MGA_ERR_ARG, "bad mask dtype %d", d->mask_dtype);
    if (d->B > 65535) return fail(MGA_ERR_UNSUPPORTED, "batch > 65535");
    sh->B = d->B; sh->C = d->C; sh->H = d->H; sh->W = d->W; sh->S = d->H * d->W;
    sh->hidden = d->hidden; sh->k = 1; sh->flags = d->flags & (MGA_HAS_MASK | MGA_SIGMOID_MASK | MGA_GATE_CLAMP);
    sh->tiny_thr = d->tiny_mask_thr; sh->eps = d->eps;
    return MGA_OK;
}
static int launch_ok(const char* what) {
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? MGA_OK : fail(MGA_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
static bool aligned16(std::initializer_list<const void*> ptrs) {
    for (const void* p : ptrs)
        if (p && (reinterpret_cast<uintptr_t>(p) & 15)) return false;
    return true;
}

template <typename T, int VEC>
static int eca_forward_t(const Shape& sh, const mga_cbam_desc* d, const T* x, const void* mask, const float* w1d, const float* beta, T* out,
                         EcaCtx ctx, EcaScratch es, cudaStream_t st) {
    Ctx c{};  // the pooling kernels only touch m
    c.m = ctx.m;
    const int nMaskTiles = (sh.S + kMaskTile - 1) / kMaskTile;
    if (sh.has_mask()) {
        const dim3 gm(nMaskTiles, sh.B);
        if (d->mask_dtype == MGA_F32) mask_prep_kernel<float><<<gm, kBlock, 0, st>>>(static_cast<const float*>(mask), sh, c, es.fs);
        else if (d->mask_dtype == MGA_BF16) mask_prep_kernel<__nv_bfloat16><<<gm, kBlock, 0, st>>>(static_cast<const __nv_bfloat16*>(mask), sh, c, es.fs);
        else mask_prep_kernel<__half><<<gm, kBlock, 0, st>>>(static_cast<const __half*>(mask), sh, c, es.fs);
    }
    const int planes = sh.B * sh.C;
    cam_pool_kernel<T, VEC, 32><<<(planes + 7) / 8, kBlock, 0, st>>>(x, sh, c, es.fs);
    eca_gate_kernel<<<sh.B, kBlock, (size_t)sh.C * sizeof(float), st>>>(sh, w1d, beta, ctx, es.fs, nMaskTiles);
    const size_t total = (size_t)planes * (sh.S / VEC);
    const int grid = (int)std::min<size_t>((total + kBlock - 1) / kBlock, (size_t)kSMs * 32);
    eca_rescale_kernel<T, VEC><<<grid, kBlock, 0, st>>>(x, out, sh, ctx);
    return launch_ok("mga_eca_forward");
}
template <typename T, int VEC>
static int eca_backward_t(const Shape& sh, const mga_cbam_desc* d, const T* x, const void* mask, const T* g, const float* w1d, T* dx, void* dmask,
                          float* dw1d, float* dbeta, EcaCtx ctx, EcaScratch es, cudaStream_t st) {
    (void)mask;
    const int planes = sh.B * sh.C;
    eca_bwd_reduce_kernel<T, VEC><<<(planes + kWarpsPerBlock - 1) / kWarpsPerBlock, kBlock, 0, st>>>(x, g, sh, es);
    eca_bwd_gate_kernel<<<sh.B, kBlock, (size_t)2 * sh.C * sizeof(float), st>>>(sh, w1d, ctx, es);
    const dim3 gt((sh.S / VEC + 31) / 32, sh.B);
    eca_bwd_dx_kernel<T, VEC><<<gt, kBlock, 0, st>>>(x, g, mask, d->mask_dtype, dx, dmask, sh, ctx, es);
    eca_wgrad_kernel<<<sh.hidden + 1, 32, 0, st>>>(sh, ctx, es, dw1d, dbeta);
    return launch_ok("mga_eca_backward");
}

// ================================================================== MGAMaskHead tail: 3x3 conv, Cin -> 1
// forward: grid (tiles of 32 units, B); a warp owns channels w, w+8, ...; a lane owns one 16-byte unit of a row segment and reads the
// three rows (y-1, y, y+1) of its channel with the left / right neighbour pixels through shuffles... kept simple and exact instead:
// thread = (pixel), loop over channels with coalesced scalar loads of the 3x3 neighbourhood (rows stay in L1/L2 across the 9 taps).
constexpr int kHeadMaxC = 256;
template <typename T>
__global__ void __launch_bounds__(kBlock) head_tail_fwd_kernel(const T* __restrict__ feat, const float* __restrict__ w, const float* __restrict__ bias,
                                                               float* __restrict__ logits, int B, int C, int H, int W) {
    extern __shared__ float hsm[];  // weights [C][9]
    for (int i = threadIdx.x; i < C * 9; i += kBlock) hsm[i] = __ldg(w + i);
    __syncthreads();
    const int S = H * W, b = blockIdx.y;
    const int p = blockIdx.x * kBlock + threadIdx.x;
    if (p >= S) return;
    const int y = p / W, xx = p - y * W;
    const T* fb = feat + (size_t)b * C * S;
    float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f;  // one accumulator per kernel row (shorter chains)
    const bool up = y > 0, dn = y + 1 < H, lf = xx > 0, rt = xx + 1 < W;
    for (int c = 0; c < C; ++c) {
        const T* r1 = fb + (size_t)c * S + p;
        const float* wc = hsm + c * 9;
        if (up) {
            const T* r0 = r1 - W;
            acc0 = fmaf(wc[1], to_f<T>(r0[0]), acc0);
            if (lf) acc0 = fmaf(wc[0], to_f<T>(r0[-1]), acc0);
            if (rt) acc0 = fmaf(wc[2], to_f<T>(r0[1]), acc0);
        }
        acc1 = fmaf(wc[4], to_f<T>(r1[0]), acc1);
        if (lf) acc1 = fmaf(wc[3], to_f<T>(r1[-1]), acc1);
        if (rt) acc1 = fmaf(wc[5], to_f<T>(r1[1]), acc1);
        if (dn) {
            const T* r2 = r1 + W;
            acc2 = fmaf(wc[7], to_f<T>(r2[0]), acc2);
            if (lf) acc2 = fmaf(wc[6], to_f<T>(r2[-1]), acc2);
            if (rt) acc2 = fmaf(wc[8], to_f<T>(r2[1]), acc2);
        }
    }
    logits[(size_t)b * S + p] = (acc0 + acc1) + acc2 + __ldg(bias);
}

// backward wrt the input: dfeat[c][p] = sum_{i,j} w[c][i][j] * g[y - (i-1)][x - (j-1)]   (thread = pixel, loop over channels)
template <typename T>
__global__ void __launch_bounds__(kBlock) head_tail_bwd_in_kernel(const float* __restrict__ g, const float* __restrict__ w, T* __restrict__ dfeat, int B,
                                                                  int C, int H, int W) {
    extern __shared__ float hsm[];
    for (int i = threadIdx.x; i < C * 9; i += kBlock) hsm[i] = __ldg(w + i);
    __syncthreads();
    const int S = H * W, b = blockIdx.y;
    const int p = blockIdx.x * kBlock + threadIdx.x;
    if (p >= S) return;
    const int y = p / W, xx = p - y * W;
    const float* gb = g + (size_t)b * S;
    float gn[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int yy = y - (i - 1), xq = xx - (j - 1);
            gn[i * 3 + j] = (yy >= 0 && yy < H && xq >= 0 && xq < W) ? __ldg(gb + yy * W + xq) : 0.0f;
        }
    T* db = dfeat + (size_t)b * C * S + p;
    for (int c = 0; c < C; ++c) {
        const float* wc = hsm + c * 9;
        float a = 0.0f;
#pragma unroll
        for (int t = 0; t < 9; ++t) a = fmaf(wc[t], gn[t], a);
        db[(size_t)c * S] = from_f<T>(a);
    }
}

// backward wrt the weights: dw[c][i][j] = sum_{b,p} feat[b][c][y+i-1][x+j-1] * g[b][p]; one CTA per channel, fp32 per-thread sums
// (9 taps), fp64 across the block; db = sum g by the extra CTA blockIdx.x == C
template <typename T>
__global__ void __launch_bounds__(kBlock) head_tail_bwd_w_kernel(const T* __restrict__ feat, const float* __restrict__ g, float* __restrict__ dw,
                                                                 float* __restrict__ dbias, int B, int C, int H, int W) {
    __shared__ double red[32];
    const int S = H * W, c = blockIdx.x;
    if (c == C) {
        double a = 0.0;
        for (size_t i = threadIdx.x; i < (size_t)B * S; i += kBlock) a += (double)__ldg(g + i);
        const double t = block_sum_d(a, red);
        if (threadIdx.x == 0) dbias[0] = (float)t;
        return;
    }
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < B; ++b) {
        const T* fc = feat + ((size_t)b * C + c) * S;
        const float* gb = g + (size_t)b * S;
        float a[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int p = threadIdx.x; p < S; p += kBlock) {
            const int y = p / W, xx = p - y * W;
            const float gv = __ldg(gb + p);
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int yy = y + i - 1, xq = xx + j - 1;
                    if (yy >= 0 && yy < H && xq >= 0 && xq < W) a[i * 3 + j] = fmaf(to_f<T>(fc[yy * W + xq]), gv, a[i * 3 + j]);
                }
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[t] += (double)a[t];
    }
    for (int t = 0; t < 9; ++t) {
        const double s = block_sum_d(acc[t], red);
        if (threadIdx.x == 0) dw[c * 9 + t] = (float)s;
    }
}

// ================================================================== ProbMaskGater (train mode)
// Noise contract (include/mga_cbam.h): element i of the call draws Philox4x32-10 with key = (seed lo, seed hi) and counter =
// (i lo, i hi, offset lo, offset hi); u1 = (r0 + 0.5) * 2^-32, u2 = (r1 + 0.5) * 2^-32 in fp32.  The caller advances `offset` per call.
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
enum { kGateGumbel = 0, kGateHardSt = 1, kGateBernoulli = 2 };
__global__ void __launch_bounds__(kBlock) gate_sample_fwd_kernel(const float* __restrict__ p_raw, const float* __restrict__ noise /* (2,n) or null */,
                                                                 float* __restrict__ out, float* __restrict__ soft_out, float* __restrict__ noise_out,
                                                                 size_t n, int mode, float tau, float p_min, float threshold,
                                                                 unsigned long long seed, unsigned long long offset) {
    const size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    float p = fminf(fmaxf(p_raw[i], 0.0f), 1.0f);
    if (p_min > 0.0f) p = fmaxf(p, p_min);
    float u1, u2;
    if (noise != nullptr) { u1 = noise[i]; u2 = noise[n + i]; }
    else {
        uint32_t c[4] = {(uint32_t)i, (uint32_t)(i >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        u1 = ((float)c[0] + 0.5f) * 2.3283064365386963e-10f;
        u2 = ((float)c[1] + 0.5f) * 2.3283064365386963e-10f;
    }
    if (noise_out != nullptr) { noise_out[i] = u1; noise_out[n + i] = u2; }
    if (mode == kGateBernoulli) {  // M = [u < p], no gradient (probmaskgater.py:91-92)
        out[i] = u1 < p ? 1.0f : 0.0f;
        return;
    }
    u1 = fminf(fmaxf(u1, 1e-6f), 1.0f - 1e-6f);
    u2 = fminf(fmaxf(u2, 1e-6f), 1.0f - 1e-6f);
    const float gn = -logf(-logf(u1)) + logf(-logf(u2));  // logistic noise (probmaskgater.py:68)
    const float pc = fminf(fmaxf(p, 1e-6f), 1.0f - 1e-6f);
    const float logit = logf(pc) - log1pf(-pc);
    const float soft = sigmoidf_acc((logit + gn) / tau);
    soft_out[i] = soft;
    out[i] = mode == kGateGumbel ? soft : (soft > threshold ? 1.0f : 0.0f);
}
__global__ void __launch_bounds__(kBlock) gate_sample_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ p_raw,
                                                                 const float* __restrict__ soft, float* __restrict__ dp, size_t n, float tau, float p_min) {
    const size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const float pr = p_raw[i];
    float p = fminf(fmaxf(pr, 0.0f), 1.0f);
    const bool floor_on = p_min > 0.0f && p < p_min;
    if (p_min > 0.0f) p = fmaxf(p, p_min);
    const float s = soft[i];
    float d = 0.0f;
    if (p > 1e-6f && p < 1.0f - 1e-6f && !floor_on && pr >= 0.0f && pr <= 1.0f) d = gout[i] * s * (1.0f - s) / tau / (p * (1.0f - p));
    dp[i] = d;
}

// ================================================================== collate: zero-pad + stack
struct CollateItem { const void* src; int h, w; };
__global__ void __launch_bounds__(kBlock) collate_kernel(const CollateItem* __restrict__ items, float* __restrict__ dst, int H, int W, int src_f32) {
    const CollateItem it = items[blockIdx.y];
    float* d = dst + (size_t)blockIdx.y * H * W;
    for (int p = blockIdx.x * kBlock + threadIdx.x; p < H * W; p += gridDim.x * kBlock) {
        const int y = p / W, x = p - y * W;
        float v = 0.0f;
        if (y < it.h && x < it.w) v = src_f32 ? static_cast<const float*>(it.src)[y * it.w + x] : (float)static_cast<const uint8_t*>(it.src)[y * it.w + x];
        d[p] = v;
    }
}

}  // namespace mga

using namespace mga;

extern "C" {

int mga_eca_workspace(const mga_cbam_desc* d, size_t* ctx_bytes, size_t* scratch_bytes) {
    Shape sh;
    if (int rc = eca_validate(d, &sh)) return rc;
    EcaCtx c;
    EcaScratch e;
    if (ctx_bytes) *ctx_bytes = eca_carve_ctx(sh, nullptr, &c);
    if (scratch_bytes) *scratch_bytes = eca_carve_scratch(sh, nullptr, &e);
    return MGA_OK;
}

int mga_eca_forward(const mga_cbam_desc* d, const void* x, const void* mask, const float* w1d, const float* beta, void* out, void* ctx_buf,
                    void* scratch, void* stream) {
    Shape sh;
    if (int rc = eca_validate(d, &sh)) return rc;
    if (!x || !out || !w1d || !beta || !ctx_buf || !scratch) return fail(MGA_ERR_ARG, "null pointer argument");
    if (sh.has_mask() && !mask) return fail(MGA_ERR_ARG, "MGA_HAS_MASK set but mask is null");
    EcaCtx ctx;
    EcaScratch es;
    eca_carve_ctx(sh, ctx_buf, &ctx);
    eca_carve_scratch(sh, scratch, &es);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int v = d->dtype == MGA_F32 ? 4 : 8;
    const bool vec = sh.S % v == 0 && aligned16({x, out});
#define MGA_ECA_F(T) (vec ? eca_forward_t<T, VecOf<T>::V>(sh, d, static_cast<const T*>(x), mask, w1d, beta, static_cast<T*>(out), ctx, es, st) \
                          : eca_forward_t<T, 1>(sh, d, static_cast<const T*>(x), mask, w1d, beta, static_cast<T*>(out), ctx, es, st))
    switch (d->dtype) {
        case MGA_F32: return MGA_ECA_F(float);
        case MGA_BF16: return MGA_ECA_F(__nv_bfloat16);
        default: return MGA_ECA_F(__half);
    }
#undef MGA_ECA_F
}

int mga_eca_backward(const mga_cbam_desc* d, const void* x, const void* mask, const void* grad_out, const float* w1d, const void* ctx_buf,
                     void* grad_x, void* grad_mask, float* grad_w1d, float* grad_beta, void* scratch, void* stream) {
    Shape sh;
    if (int rc = eca_validate(d, &sh)) return rc;
    if (!x || !grad_out || !w1d || !ctx_buf || !grad_x || !grad_w1d || !grad_beta || !scratch) return fail(MGA_ERR_ARG, "null pointer argument");
    EcaCtx ctx;
    EcaScratch es;
    eca_carve_ctx(sh, const_cast<void*>(ctx_buf), &ctx);
    eca_carve_scratch(sh, scratch, &es);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int v = d->dtype == MGA_F32 ? 4 : 8;
    const bool vec = sh.S % v == 0 && aligned16({x, grad_out, grad_x});
#define MGA_ECA_B(T) (vec ? eca_backward_t<T, VecOf<T>::V>(sh, d, static_cast<const T*>(x), mask, static_cast<const T*>(grad_out), w1d, static_cast<T*>(grad_x), grad_mask, grad_w1d, grad_beta, ctx, es, st) \
                          : eca_backward_t<T, 1>(sh, d, static_cast<const T*>(x), mask, static_cast<const T*>(grad_out), w1d, static_cast<T*>(grad_x), grad_mask, grad_w1d, grad_beta, ctx, es, st))
    switch (d->dtype) {
        case MGA_F32: return MGA_ECA_B(float);
        case MGA_BF16: return MGA_ECA_B(__nv_bfloat16);
        default: return MGA_ECA_B(__half);
    }
#undef MGA_ECA_B
}

int mga_head_tail_forward(const void* feat, const float* weight, const float* bias, float* logits, int32_t B, int32_t C, int32_t H, int32_t W,
                          int32_t dtype, void* stream) {
    if (!feat || !weight || !bias || !logits) return fail(MGA_ERR_ARG, "null pointer argument");
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || B > 65535) return fail(MGA_ERR_ARG, "bad shape");
    if (C > kHeadMaxC) return fail(MGA_ERR_UNSUPPORTED, "hidden channels > %d", kHeadMaxC);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 grid((H * W + kBlock - 1) / kBlock, B);
    const size_t smem = (size_t)C * 9 * sizeof(float);
    switch (dtype) {
        case MGA_F32: head_tail_fwd_kernel<float><<<grid, kBlock, smem, st>>>(static_cast<const float*>(feat), weight, bias, logits, B, C, H, W); break;
        case MGA_BF16: head_tail_fwd_kernel<__nv_bfloat16><<<grid, kBlock, smem, st>>>(static_cast<const __nv_bfloat16*>(feat), weight, bias, logits, B, C, H, W); break;
        case MGA_F16: head_tail_fwd_kernel<__half><<<grid, kBlock, smem, st>>>(static_cast<const __half*>(feat), weight, bias, logits, B, C, H, W); break;
        default: return fail(MGA_ERR_ARG, "bad feature dtype %d", dtype);
    }
    return launch_ok("mga_head_tail_forward");
}

int mga_head_tail_backward(const void* feat, const float* weight, const float* grad_logits, void* grad_feat, float* grad_weight, float* grad_bias,
                           int32_t B, int32_t C, int32_t H, int32_t W, int32_t dtype, void* stream) {
    if (!feat || !weight || !grad_logits || !grad_feat || !grad_weight || !grad_bias) return fail(MGA_ERR_ARG, "null pointer argument");
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || B > 65535) return fail(MGA_ERR_ARG, "bad shape");
    if (C > kHeadMaxC) return fail(MGA_ERR_UNSUPPORTED, "hidden channels > %d", kHeadMaxC);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 grid((H * W + kBlock - 1) / kBlock, B);
    const size_t smem = (size_t)C * 9 * sizeof(float);
#define MGA_HEAD_B(T)                                                                                                               \
    head_tail_bwd_in_kernel<T><<<grid, kBlock, smem, st>>>(grad_logits, weight, static_cast<T*>(grad_feat), B, C, H, W);             \
    head_tail_bwd_w_kernel<T><<<C + 1, kBlock, 0, st>>>(static_cast<const T*>(feat), grad_logits, grad_weight, grad_bias, B, C, H, W)
    switch (dtype) {
        case MGA_F32: MGA_HEAD_B(float); break;
        case MGA_BF16: MGA_HEAD_B(__nv_bfloat16); break;
        case MGA_F16: MGA_HEAD_B(__half); break;
        default: return fail(MGA_ERR_ARG, "bad feature dtype %d", dtype);
    }
#undef MGA_HEAD_B
    return launch_ok("mga_head_tail_backward");
}

int mga_gate_sample_forward(const float* p, const float* noise, float* out, float* soft, float* noise_out, size_t n, int32_t mode, float tau,
                            float p_min, float threshold, uint64_t seed, uint64_t offset, void* stream) {
    if (!p || !out || (mode != kGateBernoulli && !soft)) return fail(MGA_ERR_ARG, "null pointer argument");
    if (mode < kGateGumbel || mode > kGateBernoulli) return fail(MGA_ERR_ARG, "unknown gate mode %d", mode);
    if (!(tau > 0.0f)) return fail(MGA_ERR_ARG, "tau must be > 0");
    if (n == 0) return MGA_OK;
    gate_sample_fwd_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
        p, noise, out, soft, noise_out, n, mode, tau, p_min, threshold, (unsigned long long)seed, (unsigned long long)offset);
    return launch_ok("mga_gate_sample_forward");
}

int mga_gate_sample_backward(const float* grad_out, const float* p, const float* soft, float* grad_p, size_t n, float tau, float p_min, void* stream) {
    if (!grad_out || !p || !soft || !grad_p) return fail(MGA_ERR_ARG, "null pointer argument");
    if (n == 0) return MGA_OK;
    gate_sample_bwd_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(grad_out, p, soft, grad_p, n, tau, p_min);
    return launch_ok("mga_gate_sample_backward");
}

int mga_collate_masks(const void* items_dev, float* dst, int32_t B, int32_t H, int32_t W, int32_t src_dtype, void* stream) {
    if (!items_dev || !dst) return fail(MGA_ERR_ARG, "null pointer argument");
    if (B <= 0 || H <= 0 || W <= 0 || B > 65535) return fail(MGA_ERR_ARG, "bad shape");
    if (src_dtype != MGA_F32 && src_dtype != MGA_U8) return fail(MGA_ERR_ARG, "collate source dtype must be u8 or f32");
    const dim3 grid(std::min((H * W + kBlock - 1) / kBlock, 64), B);
    collate_kernel<<<grid, kBlock, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const CollateItem*>(items_dev), dst, H, W, src_dtype == MGA_F32);
    return launch_ok("mga_collate_masks");
}

}  // extern "C"
