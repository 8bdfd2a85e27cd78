// mga_cbam.cu -- C ABI (include/mga_cbam.h) and launch logic for the mask-guided CBAM path.
#include <cstdio>
#include <cstring>
#include <type_traits>

#include "cbam_bwd.cuh"
#include "cbam_fwd.cuh"
#include "common.cuh"

namespace mga {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static int check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(MGA_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return MGA_OK;
}

static int validate(const mga_cbam_desc* d, Shape* sh) {
    if (!d) return fail(MGA_ERR_ARG, "null descriptor");
    if (d->B <= 0 || d->C <= 0 || d->H <= 0 || d->W <= 0) return fail(MGA_ERR_ARG, "bad shape (%d,%d,%d,%d)", d->B, d->C, d->H, d->W);
    if (d->hidden <= 0) return fail(MGA_ERR_ARG, "hidden must be >= 1");
    if (d->ksize < 1 || d->ksize > kMaxK || d->ksize % 2 == 0) return fail(MGA_ERR_UNSUPPORTED, "spatial kernel %d: odd sizes up to %d are built", d->ksize, kMaxK);
    if (d->dtype != MGA_F32 && d->dtype != MGA_BF16 && d->dtype != MGA_F16) return fail(MGA_ERR_ARG, "bad feature dtype %d", d->dtype);
    if (d->mask_dtype != MGA_F32 && d->mask_dtype != MGA_BF16 && d->mask_dtype != MGA_F16) return fail(MGA_ERR_ARG, "bad mask dtype %d", d->mask_dtype);
    if ((long long)d->H * d->W > (1 << 28) || (long long)d->B * d->C > (1 << 28)) return fail(MGA_ERR_UNSUPPORTED, "shape too large");
    if (d->B > 65535) return fail(MGA_ERR_UNSUPPORTED, "batch > 65535");
    sh->B = d->B; sh->C = d->C; sh->H = d->H; sh->W = d->W; sh->S = d->H * d->W;
    sh->hidden = d->hidden; sh->k = d->ksize; sh->flags = d->flags;
    sh->tiny_thr = d->tiny_mask_thr; sh->eps = d->eps;
    return MGA_OK;
}

// ------------------------------------------------------------------ layouts
struct Carver {
    char* base;
    size_t off = 0;
    template <typename T> T* take(size_t n) {
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += align256(n * sizeof(T));
        return p;
    }
};

static size_t carve_ctx(const Shape& s, void* base, Ctx* c) {
    Carver k{static_cast<char*>(base)};
    const size_t BS = (size_t)s.B * s.S, BC = (size_t)s.B * s.C, BH = (size_t)s.B * s.hidden;
    c->consts = k.take<float>(8);
    c->m = k.take<float>(BS);
    c->a = k.take<float>(BS);
    c->pmax = k.take<float>(BS);
    c->pavg = k.take<float>(BS);
    c->idx = k.take<int>(BS);
    c->s = k.take<float>(BC);
    c->avg = k.take<float>(BC);
    c->mx = k.take<float>(BC);
    c->apool = k.take<float>(BC);
    c->amax = k.take<int>(BC);
    c->use = k.take<float>(s.B);
    c->den = k.take<float>(s.B);
    c->msum = k.take<float>(s.B);
    c->ha = k.take<float>(BH);
    c->hm = k.take<float>(BH);
    return k.off;
}

static int tiles_of(const Shape& s, int vec) { return (s.S / vec + 31) / 32; }
static int conv_ctas(const Shape& s) { return ((s.W + kConvTW - 1) / kConvTW) * ((s.H + kConvTH - 1) / kConvTH) * s.B; }

static size_t carve_fwd(const Shape& s, void* base, FwdScratch* f) {
    Carver k{static_cast<char*>(base)};
    const size_t BC = (size_t)s.B * s.C;
    f->sxm = k.take<float>(BC);
    f->sx = k.take<float>(BC);
    f->best = k.take<float>(BC);
    f->bidx = k.take<int>(BC);
    return k.off;
}

static size_t carve_bwd(const Shape& s, void* base, BwdScratch* b) {
    Carver k{static_cast<char*>(base)};
    const size_t BS = (size_t)s.B * s.S, BC = (size_t)s.B * s.C, BH = (size_t)s.B * s.hidden;
    const size_t nT = tiles_of(s, 1);  // scalar tiling is the largest tile count
    b->T = k.take<float>(BS);
    b->dcat = k.take<float>(3 * BS);
    b->epart = k.take<float>((size_t)s.B * nT * s.C);
    b->gxpart = k.take<float>((size_t)s.B * nT * s.C);
    b->qpart = k.take<float>((size_t)s.B * nT * s.C);
    b->cA = k.take<float>(BC);
    b->cG = k.take<float>(BC);
    b->cM = k.take<float>(BC);
    b->kb = k.take<float>(s.B);
    b->dz = k.take<float>(BC);
    b->dha = k.take<float>(BH);
    b->dhm = k.take<float>(BH);
    b->convpart = k.take<float>((size_t)conv_ctas(s) * (3 * kMaxK * kMaxK + 1));
    b->alphapart = k.take<float>(s.B);
    return k.off;
}

// vector width usable for this call: plane size divisible and every pointer 16-byte aligned
static int pick_vec(const Shape& s, int dtype, std::initializer_list<const void*> ptrs) {
    const int v = dtype == MGA_F32 ? 4 : 8;
    if (s.S % v) return 1;
    for (const void* p : ptrs)
        if (p && (reinterpret_cast<uintptr_t>(p) & 15)) return 1;
    return v;
}

// ------------------------------------------------------------------ forward
template <typename T, int VEC>
static int forward_split(const Shape& sh, const T* x, const void* mask, int mask_dtype, const mga_cbam_params& p, T* out, Ctx ctx,
                         FwdScratch fs, cudaStream_t st) {
    const int U = sh.S / VEC;
    if (sh.has_mask()) {
        if (mask_dtype == MGA_F32) mask_prep_kernel<float><<<sh.B, kBlock, 0, st>>>(static_cast<const float*>(mask), sh, ctx);
        else if (mask_dtype == MGA_BF16) mask_prep_kernel<__nv_bfloat16><<<sh.B, kBlock, 0, st>>>(static_cast<const __nv_bfloat16*>(mask), sh, ctx);
        else mask_prep_kernel<__half><<<sh.B, kBlock, 0, st>>>(static_cast<const __half*>(mask), sh, ctx);
    }
    const int planes = sh.B * sh.C;
    if (U >= 128) cam_pool_kernel<T, VEC, 32><<<(planes + 7) / 8, kBlock, 0, st>>>(x, sh, ctx, fs);
    else if (U >= 48) cam_pool_kernel<T, VEC, 16><<<(planes + 15) / 16, kBlock, 0, st>>>(x, sh, ctx, fs);
    else cam_pool_kernel<T, VEC, 8><<<(planes + 31) / 32, kBlock, 0, st>>>(x, sh, ctx, fs);
    const size_t mlp_smem = (2 * (size_t)sh.C + 2 * sh.hidden) * sizeof(float);
    cam_mlp_kernel<<<sh.B, kBlock, mlp_smem, st>>>(sh, p, ctx, fs);
    const dim3 gtile((U + 31) / 32, sh.B);
    sam_reduce_kernel<T, VEC><<<gtile, kBlock, 0, st>>>(x, sh, ctx);
    const dim3 gconv((sh.W + kConvTW - 1) / kConvTW, (sh.H + kConvTH - 1) / kConvTH, sh.B);
    sam_conv_kernel<<<gconv, kBlock, 0, st>>>(sh, p.wsam, ctx);
    const size_t total = (size_t)planes * U;
    const int grid = (int)std::min<size_t>((total + kBlock - 1) / kBlock, (size_t)kSMs * 32);
    rescale_kernel<T, VEC><<<grid, kBlock, 0, st>>>(x, out, sh, ctx);
    return check_launch("mga_cbam_forward");
}

template <typename T>
static int forward_t(const Shape& sh, const mga_cbam_desc* d, const void* x, const void* mask, const mga_cbam_params& p, void* out,
                     Ctx ctx, FwdScratch fs, cudaStream_t st) {
    const int vec = pick_vec(sh, d->dtype, {x, out});
    if (vec == 1) return forward_split<T, 1>(sh, static_cast<const T*>(x), mask, d->mask_dtype, p, static_cast<T*>(out), ctx, fs, st);
    return forward_split<T, VecOf<T>::V>(sh, static_cast<const T*>(x), mask, d->mask_dtype, p, static_cast<T*>(out), ctx, fs, st);
}

// ------------------------------------------------------------------ backward
template <typename T, int VEC, typename TM>
static int backward_split(const Shape& sh, const T* x, const TM* mask, const T* g, const mga_cbam_params& p, Ctx ctx, T* dx, TM* dmask,
                          const mga_cbam_grads& gp, BwdScratch bs, cudaStream_t st) {
    const int U = sh.S / VEC;
    const int nT = (U + 31) / 32;
    const dim3 gtile(nT, sh.B);
    bwd_reduce1_kernel<T, VEC><<<gtile, kBlock, 0, st>>>(x, g, sh, ctx, bs, nT);
    const dim3 gconv((sh.W + kBT_W - 1) / kBT_W, (sh.H + kBT_H - 1) / kBT_H, sh.B);
    const int nconv = gconv.x * gconv.y * gconv.z;
    bwd_conv_kernel<<<gconv, kBlock, 0, st>>>(sh, p.wsam, ctx, bs);
    if (!sh.samcam_add()) bwd_reduce2_kernel<T, VEC><<<gtile, kBlock, 0, st>>>(x, sh, ctx, bs, nT);
    const size_t mlp_smem = ((size_t)sh.C + 2 * sh.hidden) * sizeof(float);
    bwd_mlp_kernel<<<sh.B, kBlock, mlp_smem, st>>>(sh, p, ctx, bs, nT);
    bwd_dx_kernel<T, VEC, TM><<<gtile, kBlock, 0, st>>>(x, g, mask, dx, dmask, sh, ctx, bs);
    const int nw = 2 * sh.C * sh.hidden + sh.C + sh.hidden + 3 * sh.k * sh.k + 1;
    bwd_wgrad_kernel<<<(nw + kBlock - 1) / kBlock, kBlock, 0, st>>>(sh, ctx, bs, gp, nconv);
    return check_launch("mga_cbam_backward");
}

template <typename T, typename TM>
static int backward_tm(const Shape& sh, const mga_cbam_desc* d, const void* x, const void* mask, const void* g, const mga_cbam_params& p,
                       Ctx ctx, void* dx, void* dmask, const mga_cbam_grads& gp, BwdScratch bs, cudaStream_t st) {
    const int vec = pick_vec(sh, d->dtype, {x, g, dx});
    if (vec == 1)
        return backward_split<T, 1, TM>(sh, static_cast<const T*>(x), static_cast<const TM*>(mask), static_cast<const T*>(g), p, ctx,
                                        static_cast<T*>(dx), static_cast<TM*>(dmask), gp, bs, st);
    return backward_split<T, VecOf<T>::V, TM>(sh, static_cast<const T*>(x), static_cast<const TM*>(mask), static_cast<const T*>(g), p, ctx,
                                              static_cast<T*>(dx), static_cast<TM*>(dmask), gp, bs, st);
}

template <typename T>
static int backward_t(const Shape& sh, const mga_cbam_desc* d, const void* x, const void* mask, const void* g, const mga_cbam_params& p,
                      Ctx ctx, void* dx, void* dmask, const mga_cbam_grads& gp, BwdScratch bs, cudaStream_t st) {
    switch (d->mask_dtype) {
        case MGA_F32: return backward_tm<T, float>(sh, d, x, mask, g, p, ctx, dx, dmask, gp, bs, st);
        case MGA_BF16: return backward_tm<T, __nv_bfloat16>(sh, d, x, mask, g, p, ctx, dx, dmask, gp, bs, st);
        default: return backward_tm<T, __half>(sh, d, x, mask, g, p, ctx, dx, dmask, gp, bs, st);
    }
}

}  // namespace mga

using namespace mga;

extern "C" {

int mga_abi_version(void) { return MGA_ABI_VERSION; }
const char* mga_last_error(void) { return g_err; }

int mga_cbam_workspace(const mga_cbam_desc* d, size_t* ctx_bytes, size_t* scratch_bytes) {
    Shape sh;
    if (int rc = validate(d, &sh)) return rc;
    Ctx c;
    FwdScratch f;
    BwdScratch b;
    if (ctx_bytes) *ctx_bytes = carve_ctx(sh, nullptr, &c);
    if (scratch_bytes) *scratch_bytes = std::max(carve_fwd(sh, nullptr, &f), carve_bwd(sh, nullptr, &b));
    return MGA_OK;
}

int mga_cbam_forward(const mga_cbam_desc* d, const void* x, const void* mask, const mga_cbam_params* p, void* out, void* ctx_buf,
                     void* scratch, void* stream) {
    Shape sh;
    if (int rc = validate(d, &sh)) return rc;
    if (!x || !out || !p || !ctx_buf || !scratch) return fail(MGA_ERR_ARG, "null pointer argument");
    if (!p->w1 || !p->b1 || !p->w2 || !p->b2 || !p->wsam || !p->beta) return fail(MGA_ERR_ARG, "null parameter pointer");
    if (sh.has_mask() && !mask) return fail(MGA_ERR_ARG, "MGA_HAS_MASK set but mask is null");
    Ctx ctx;
    FwdScratch fs;
    carve_ctx(sh, ctx_buf, &ctx);
    carve_fwd(sh, scratch, &fs);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (d->dtype) {
        case MGA_F32: return forward_t<float>(sh, d, x, mask, *p, out, ctx, fs, st);
        case MGA_BF16: return forward_t<__nv_bfloat16>(sh, d, x, mask, *p, out, ctx, fs, st);
        default: return forward_t<__half>(sh, d, x, mask, *p, out, ctx, fs, st);
    }
}

int mga_cbam_backward(const mga_cbam_desc* d, const void* x, const void* mask, const void* grad_out, const mga_cbam_params* p,
                      const void* ctx_buf, void* grad_x, void* grad_mask, const mga_cbam_grads* gp, void* scratch, void* stream) {
    Shape sh;
    if (int rc = validate(d, &sh)) return rc;
    if (!x || !grad_out || !p || !ctx_buf || !grad_x || !gp || !scratch) return fail(MGA_ERR_ARG, "null pointer argument");
    if (!gp->w1 || !gp->b1 || !gp->w2 || !gp->b2 || !gp->wsam || !gp->beta) return fail(MGA_ERR_ARG, "null gradient pointer");
    if (sh.has_mask() && !mask) return fail(MGA_ERR_ARG, "MGA_HAS_MASK set but mask is null");
    Ctx ctx;
    BwdScratch bs;
    carve_ctx(sh, const_cast<void*>(ctx_buf), &ctx);
    carve_bwd(sh, scratch, &bs);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (d->dtype) {
        case MGA_F32: return backward_t<float>(sh, d, x, mask, grad_out, *p, ctx, grad_x, grad_mask, *gp, bs, st);
        case MGA_BF16: return backward_t<__nv_bfloat16>(sh, d, x, mask, grad_out, *p, ctx, grad_x, grad_mask, *gp, bs, st);
        default: return backward_t<__half>(sh, d, x, mask, grad_out, *p, ctx, grad_x, grad_mask, *gp, bs, st);
    }
}

int mga_cbam_ctx_view(const mga_cbam_desc* d, const void* ctx_buf, int which, const float** ptr, size_t* count) {
    Shape sh;
    if (int rc = validate(d, &sh)) return rc;
    if (!ctx_buf || !ptr || !count) return fail(MGA_ERR_ARG, "null pointer argument");
    Ctx ctx;
    carve_ctx(sh, const_cast<void*>(ctx_buf), &ctx);
    switch (which) {
        case 0: *ptr = ctx.s; *count = (size_t)sh.B * sh.C; return MGA_OK;
        case 1: *ptr = ctx.a; *count = (size_t)sh.B * sh.S; return MGA_OK;
        case 2: *ptr = ctx.m; *count = (size_t)sh.B * sh.S; return MGA_OK;
        case 3: *ptr = ctx.pmax; *count = (size_t)sh.B * sh.S; return MGA_OK;
        case 4: *ptr = ctx.pavg; *count = (size_t)sh.B * sh.S; return MGA_OK;
        default: return fail(MGA_ERR_ARG, "unknown ctx view %d", which);
    }
}

}  // extern "C"
