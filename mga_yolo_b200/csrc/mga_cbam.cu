// mga_cbam.cu -- C ABI (include/mga_cbam.h) and launch logic for the mask-guided CBAM path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

#include "cbam_bwd.cuh"
#include "cbam_cluster.cuh"
#include "cbam_concat.cuh"
#include "cbam_conv.cuh"
#include "cbam_fwd.cuh"
#include "common.cuh"
#ifdef MGA_TUNING
#include "cbam_persist.cuh"  // experiment (profiles/r2_persistent_experiment.md): tuning library only, opt-in with MGA_PF=1
#endif

namespace mga {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// ------------------------------------------------------------------ launch accounting / per-kernel CUDA-event timing
// Every kernel launch of the library goes through a LaunchScope: it bumps the launch counter
// (bench.py's "gpu_launches") and, when profiling is switched on, brackets the launch with
// CUDA events on the launching stream so bench.py can report the dominant kernel's duration.
// Forward and autograd's per-device backward threads launch concurrently: the counter is atomic, the record list is
// guarded by a mutex (records are addressed by index, never by reference, so growth of the vector is harmless).
struct ProfRec { const char* name; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;
static std::atomic<bool> g_prof_on{false};
static std::atomic<unsigned long long> g_launches{0};

struct LaunchScope {
    cudaStream_t st;
    cudaEvent_t end = nullptr;
    LaunchScope(const char* name, cudaStream_t s) : st(s) {
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (g_prof_on.load(std::memory_order_relaxed)) {
            ProfRec r{name, nullptr, nullptr};
            cudaEventCreate(&r.a);
            cudaEventCreate(&r.b);
            cudaEventRecord(r.a, st);
            end = r.b;
            std::lock_guard<std::mutex> lk(g_prof_mu);
            g_prof.push_back(r);
        }
    }
    ~LaunchScope() {
        if (end) cudaEventRecord(end, st);
    }
};
#define MGA_LAUNCH(name, st, ...) do { LaunchScope _ls(name, st); __VA_ARGS__; } while (0)

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
static int conv_ctas(const Shape& s) {
    const int generic = std::max(((s.W + kConvTW - 1) / kConvTW) * ((s.H + kConvTH - 1) / kConvTH) * s.B, 64 * s.B);  // cluster path: <= 16 CTAs per sample, persistent path: <= 64
    if (s.W % 4) return generic;
    const ConvGeom cg = conv_geom(s.W);
    return std::max(generic, ((s.H + cg.RB - 1) / cg.RB) * s.B);
}

// ------------------------------------------------------------------ TMA tensor maps of the (B,H,W) fp32 planes
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
// box (boxW, boxH, 1) over a (B,H,W) fp32 plane; zero fill outside the image.  false -> stage with plain loads instead.
static bool make_plane_maps_box(PlaneMaps* out, const float* const (&planes)[3], const Shape& sh, int boxW, int boxH) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc || boxW > 256 || boxH > 256 || (boxW * 4) % 16) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)sh.W, (cuuint64_t)sh.H, (cuuint64_t)sh.B};
    const cuuint64_t strides[2] = {(cuuint64_t)sh.W * 4, (cuuint64_t)sh.S * 4};
    const cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    for (int pl = 0; pl < 3; ++pl) {
        if (planes[pl] == nullptr) { std::memset(&out->m[pl], 0, sizeof(CUtensorMap)); continue; }
        if (reinterpret_cast<uintptr_t>(planes[pl]) & 15) return false;
        const CUresult r = enc(&out->m[pl], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(planes[pl]), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return false;
    }
    return true;
}
static bool make_plane_maps(PlaneMaps* out, const float* const (&planes)[3], const Shape& sh, const ConvGeom& cg) {
    return cg.use_tma && make_plane_maps_box(out, planes, sh, cg.TWp, cg.rowsT);
}

template <typename K>
static void allow_big_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// exchange slots and team counters of the persistent kernels (cbam_persist.cuh); first in the scratch so that forward and backward agree
struct PfScratch { unsigned* ctr; float* xchg; int xstride; };
#ifdef MGA_TUNING
static void carve_pf(const Shape& s, Carver& k, PfScratch* p) {
    p->xstride = 4 * (4 * s.C + 8);  // a team's buffer is [4C+1][TS rounded up to 4] floats: fits TS slots for every TS >= 1
    p->ctr = k.take<unsigned>((size_t)kPfMaxLv * kPfMaxCtas);
    p->xchg = k.take<float>((size_t)kPfMaxCtas * p->xstride);
}
#else
static void carve_pf(const Shape&, Carver&, PfScratch* p) { p->ctr = nullptr; p->xchg = nullptr; p->xstride = 0; }  // (the persistent experiment is not in the product library)
#endif

static size_t carve_fwd(const Shape& s, void* base, FwdScratch* f, PfScratch* pf = nullptr) {
    Carver k{static_cast<char*>(base)};
    PfScratch tmp;
    carve_pf(s, k, pf ? pf : &tmp);
    const size_t BC = (size_t)s.B * s.C;
    f->sxm = k.take<float>(BC);
    f->sx = k.take<float>(BC);
    f->best = k.take<float>(BC);
    f->bidx = k.take<int>(BC);
    f->mpart = k.take<float>((size_t)s.B * ((s.S + kMaskTile - 1) / kMaskTile));
    return k.off;
}

static size_t carve_bwd(const Shape& s, void* base, BwdScratch* b, PfScratch* pf = nullptr) {
    Carver k{static_cast<char*>(base)};
    PfScratch tmp;
    carve_pf(s, k, pf ? pf : &tmp);
    const size_t BS = (size_t)s.B * s.S, BC = (size_t)s.B * s.C, BH = (size_t)s.B * s.hidden;
    const size_t nT = tiles_of(s, 1);  // scalar tiling is the largest tile count
    b->T = k.take<float>(BS);
    b->dcat = k.take<float>(3 * BS);
    b->epart = k.take<float>((size_t)s.B * nT * s.C);
    b->gxpart = k.take<float>((size_t)s.B * nT * s.C);
    b->qpart = k.take<float>((size_t)s.B * nT * s.C);
    b->psum = k.take<float>(3 * BC);
    b->cA = k.take<float>(BC);
    b->cG = k.take<float>(BC);
    b->cM = k.take<float>(BC);
    b->kb = k.take<float>(s.B);
    b->dz = k.take<float>(BC);
    b->dha = k.take<float>(BH);
    b->dhm = k.take<float>(BH);
    b->convpart = k.take<float>((size_t)conv_ctas(s) * (3 * kMaxK * kMaxK + 1));
    b->alphapart = k.take<double>(s.B);
    b->atpart = k.take<double>((size_t)conv_ctas(s));
    return k.off;
}

// Tile shape of the per-pixel / per-channel reduce kernels: LPT lanes per channel row (x UPT units per lane).  Large planes
// take 64-unit tiles; small planes shrink the tile to 16 or 8 units so that the grid still has >= ~6 CTAs per SM.
struct TileCfg { int lpt, upt; };
// Plane class: chosen from the per-sample plane size only (never from B): the summation order inside a sample must not
// depend on how the batch is sharded, so a sample gives bit-identical results alone, in a batch of 8 or in a batch of 64.
static int plane_class(int U) { return U >= 1024 ? 0 : (U >= 256 ? 1 : 2); }
static TileCfg pick_tiles(const Shape& s, int U, int vec) {
    (void)s;
    if (vec == 1) return {32, 1};
    switch (plane_class(U)) {
        case 0: return {32, vec == 8 ? 1 : 2};
        case 1: return {16, 1};
        default: return {8, 1};
    }
}
static int pool_tpp(int U, int vec) {  // threads per (b,c) plane of the pooling phase
    if (vec == 1) return U > 6 * 128 ? 256 : (U > 6 * 64 ? 128 : (U > 6 * 32 ? 64 : 32));
    switch (plane_class(U)) {
        case 0: return 256;
        case 1: return 128;
        default: return 32;
    }
}
// expands to the four tile instantiations; VEC == 1 only ever uses <32,1>
#define MGA_TILE_DISPATCH(cfg, CALL)                                   \
    do {                                                               \
        if ((cfg).lpt == 32 && (cfg).upt == 2) { if constexpr (VEC != 8) { CALL(32, 2); } } \
        else if ((cfg).lpt == 32) { CALL(32, 1); }                     \
        else if ((cfg).lpt == 16) { if constexpr (VEC > 1) { CALL(16, 1); } } \
        else { if constexpr (VEC > 1) { CALL(8, 1); } }                \
    } while (0)

// vector width usable for this call: plane size divisible and every pointer 16-byte aligned
static int pick_vec(const Shape& s, int dtype, std::initializer_list<const void*> ptrs) {
    const int v = dtype == MGA_F32 ? 4 : 8;
    if (s.S % v) return 1;
    for (const void* p : ptrs)
        if (p && (reinterpret_cast<uintptr_t>(p) & 15)) return 1;
    return v;
}

// ------------------------------------------------------------------ cluster-per-sample path (cbam_cluster.cuh): geometry + launch
// Tuning knobs (MGA_CL_* environment variables) exist only in -DMGA_TUNING builds (tools/); the product library has none.
#ifdef MGA_TUNING
static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}
#else
static constexpr int env_int(const char*, int dflt) { return dflt; }
#endif

// Smallest cluster whose per-CTA share of the sample (x, plus g in backward) is <= ~224 KB: with 2 CTAs per SM that keeps
// ~60 MB of feature map in flight on the whole GPU, well inside the 126 MB L2, so the later passes are L2 hits.
// Chosen from the per-sample shape only (never from B).  false -> the one-kernel-per-phase path is used.
static bool cl_geometry(const Shape& sh, int esize, bool bwd, ClGeom* out) {
    static const int enabled = env_int("MGA_CL", 1);
    static const int kb_f = env_int("MGA_CL_KB_F", 224), kb_b = env_int("MGA_CL_KB_B", 448);
    static const int cs_f = env_int("MGA_CL_CS_F", 0), cs_b = env_int("MGA_CL_CS_B", 0);  // tuning overrides
    static const int pf_f = env_int("MGA_CL_PREFETCH_F", 0), pf_b = env_int("MGA_CL_PREFETCH_B", 0);
    if (!enabled || (sh.gates_only() && bwd)) return false;  // gates-only forward = cl_fwd without its last phase; the gates backward stays per phase
    static const int max_kb_f = env_int("MGA_CL_MAXKB_F", 1 << 30), max_kb_b = env_int("MGA_CL_MAXKB_B", 1 << 30);  // tuning: larger samples -> split path
    if ((double)sh.C * sh.S * esize > 1024.0 * (bwd ? max_kb_b : max_kb_f)) return false;
    static const int min_kb_f = env_int("MGA_CL_MINKB_F", 0), min_kb_b = env_int("MGA_CL_MINKB_B", 0);              // tuning: smaller samples -> split path
    if ((double)sh.C * sh.S * esize < 1024.0 * (bwd ? min_kb_b : min_kb_f)) return false;
    const int vec = 16 / esize;
    if (sh.S % vec || sh.W % 4 || sh.C < 1) return false;
    int rowq = 1;
    while ((rowq * sh.W) % vec) ++rowq;
    const double bytes = (double)sh.C * sh.S * esize * (bwd ? 2.0 : 1.0);
    const double target = 1024.0 * (bwd ? kb_b : kb_f);
    const int forced = bwd ? cs_b : cs_f;
    ClGeom best{};
    bool have = false;
    static const int any_cs = env_int("MGA_CL_ANYCS", 0);  // tuning: cluster sizes that are not powers of two
    for (int CS = 1; CS <= 16; CS = any_cs ? CS + 1 : CS * 2) {
        int rowsPer = (sh.H + CS - 1) / CS;
        rowsPer = (rowsPer + rowq - 1) / rowq * rowq;
        if (CS > 1 && (CS - 1) * rowsPer >= sh.H) { if (any_cs) continue; break; }  // trailing ranks would own nothing
        const int nP = rowsPer * sh.W, nU = nP / vec;
        const int NT = bwd ? kClNTB : kClNTF;
        if (nU > NT) continue;
        ClGeom g{};
        g.NT = NT;
        g.CS = CS;
        g.rowsPer = rowsPer;
        g.nUmax = nU;
        g.nPmax = nP;
        g.G = std::max(1, std::min(NT / nU, sh.C));
        g.CG = (((sh.C + g.G - 1) / g.G) + 7) & ~7;
        const int K = bwd ? (esize == 4 ? kClKB : kClKB16) : kClKF;
        g.K = K;
        g.LPT = (nU + K - 1) / K;
        g.slots = NT / g.LPT;
        g.tileRows = rowsPer + kMaxK - 1;
        g.TWp = sh.W + 8;
        g.planeT = (g.tileRows * g.TWp + 31) & ~31;
        g.stage_w = 0;
        g.smem_bytes = 4 * (bwd ? cl_bwd_off(sh.C, sh.hidden, g).total : cl_fwd_off(sh.C, sh.hidden, g).total);
        if (cl_mlp_in_smem(sh.C, sh.hidden) && g.smem_bytes + 8 * sh.C * sh.hidden <= kClTwoCtaSmem) {  // staging must not cost the 2nd CTA per SM
            g.stage_w = 1;
            g.smem_bytes = 4 * (bwd ? cl_bwd_off(sh.C, sh.hidden, g).total : cl_fwd_off(sh.C, sh.hidden, g).total);
        }
        g.prefetch = bwd ? pf_b : pf_f;
        static const int min_smem_f = env_int("MGA_CL_MINSMEM_F", 0), min_smem_b = env_int("MGA_CL_MINSMEM_B", 0);  // tuning: caps the CTAs per SM
        g.smem_bytes = std::max(g.smem_bytes, bwd ? min_smem_b : min_smem_f);
        if (g.smem_bytes > kSmemLimit) continue;
        best = g;
        have = true;
        if (forced ? CS >= forced : bytes / CS <= target) break;
    }
    if (!have) return false;
    if (!forced && bytes / best.CS > 2.0 * target) return false;  // sample too large to stay L2-resident: split path
    // gates-only forward (concat modes): measured on B200 at cfg4 (bf16, B = 128) the cluster kernel wins only for small samples
    // (512x20x20: 67 us vs 161 us per phase); with clusters of 8 / 16 the four per-phase kernels are faster (296 vs 368 us, 470 vs 594 us)
    if (sh.gates_only() && best.CS > 4) return false;
    static const int debug = env_int("MGA_CL_DEBUG", 0);
    if (debug)
        fprintf(stderr, "[mga] cluster %s C=%d %dx%d e=%d: CS=%d rows=%d nU=%d G=%d CG=%d LPT=%d slots=%d smem=%d B\n", bwd ? "bwd" : "fwd", sh.C, sh.H,
                sh.W, esize, best.CS, best.rowsPer, best.nUmax, best.G, best.CG, best.LPT, best.slots, best.smem_bytes);
    *out = best;
    return true;
}

template <typename Kern, typename... Args>
static int launch_cl(const char* name, Kern kernel, const ClGeom& gm, int nClusters, cudaStream_t st, Args... args) {
    // function attributes and the occupancy answer are per-DEVICE state: both caches are keyed by (device, kernel)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail(MGA_ERR_CUDA, "%s: cudaGetDevice failed", name);
    struct Cfg { const void* k; int dev; };
    static thread_local Cfg configured[64] = {};
    bool done = false;
    for (const Cfg& c : configured) done |= (c.k == (const void*)kernel && c.dev == dev);
    if (!done) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return fail(MGA_ERR_CUDA, "%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
        for (auto& c : configured)
            if (!c.k) { c = Cfg{(const void*)kernel, dev}; break; }  // table full: the attributes are simply set again next time
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(nClusters * gm.CS));
    cfg.blockDim = dim3((unsigned)gm.NT);
    cfg.dynamicSmemBytes = (size_t)gm.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)gm.CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static const int policy = env_int("MGA_CL_POLICY", 0);  // tuning: 1 = spread, 2 = load balancing (default: the driver's choice)
    if (policy == 1 || policy == 2) {
        attr[1].id = cudaLaunchAttributeClusterSchedulingPolicyPreference;
        attr[1].val.clusterSchedulingPolicyPreference = policy == 1 ? cudaClusterSchedulingPolicySpread : cudaClusterSchedulingPolicyLoadBalancing;
        cfg.numAttrs = 2;
    }
    cudaError_t e;
    {
        // can at least one cluster of this size / shared-memory footprint be resident?  (cached; 0 -> the caller takes the per-phase path)
        struct Occ { const void* k; int dev, cs, smem, n; };
        static thread_local Occ cache[64] = {};
        int n = -1;
        for (const Occ& c : cache)
            if (c.k == (const void*)kernel && c.dev == dev && c.cs == gm.CS && c.smem == gm.smem_bytes) n = c.n;
        if (n < 0) {
            if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { n = 0; cudaGetLastError(); }
            for (Occ& c : cache)
                if (!c.k) { c = Occ{(const void*)kernel, dev, gm.CS, gm.smem_bytes, n}; break; }
        }
        if (n <= 0) return MGA_ERR_UNSUPPORTED;
    }
    {
        LaunchScope ls(name, st);
        e = cudaLaunchKernelEx(&cfg, kernel, args..., gm);  // the geometry is always the last kernel argument
    }
    if (e != cudaSuccess) return fail(MGA_ERR_CUDA, "%s: launch (cluster %d, %d B smem): %s", name, gm.CS, gm.smem_bytes, cudaGetErrorString(e));
    return MGA_OK;
}


#ifdef MGA_TUNING
// ------------------------------------------------------------------ persistent shared-memory-resident path (cbam_persist.cuh): geometry + launch
struct PfCall {  // one pyramid level of a (multi-level) call, host side
    Shape sh;
    int dtype, mdt;
    const void* x; const void* mask; void* out;
    const void* g; void* dx; void* dmask;
    mga_cbam_params prm;
    Ctx ctx;
    BwdScratch bs;
    PfScratch ps;
};

static int pf_gcd(int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; }

// per-level geometry for `TSwant` CTAs per sample (0: the fewest whose slice is <= kPfSliceBytes).  false -> not a persistent shape.
static bool pf_level_geom(const Shape& sh, int esize, bool bwd, int TSwant, PfLevel* lv) {
    const int vec = 16 / esize, NT = bwd ? kPfNTB : kPfNTF;
    if (sh.gates_only() || sh.C % kPfNCH || sh.C < 16 || sh.C / kPfNCH > 256 || sh.W % 4 || sh.S % vec) return false;
    int rowq = 1;
    while ((rowq * sh.W) % vec) ++rowq;
    const long rowBytes = (long)sh.C * sh.W * esize;
    int rowsPer;
    if (TSwant > 0) {
        rowsPer = ((sh.H + TSwant - 1) / TSwant + rowq - 1) / rowq * rowq;
    } else {
        rowsPer = (int)(kPfSliceBytes / rowBytes) / rowq * rowq;
        if (rowsPer < rowq) return false;
        const int TS0 = (sh.H + rowsPer - 1) / rowsPer;
        rowsPer = ((sh.H + TS0 - 1) / TS0 + rowq - 1) / rowq * rowq;  // same team, balanced rows
    }
    if (rowsPer < 1 || (long)rowsPer * rowBytes > kPfSliceBytes) return false;
    const int TS = (sh.H + rowsPer - 1) / rowsPer;
    if (TS > 64 || (TSwant > 0 && TS != TSwant)) return false;
    const int nP = rowsPer * sh.W, nU = nP / vec;
    if (nU > NT || nU < 1) return false;
    const int nsplit = nP <= 256 ? 1 : 2;
    if (nP % nsplit) return false;
    const int nPbox = nP / nsplit;
    if (nPbox > 256 || nPbox % vec) return false;
    lv->sh = sh;
    lv->TS = TS;
    lv->rowsPer = rowsPer;
    lv->nP = nP;
    lv->nU = nU;
    lv->nsplit = nsplit;
    lv->nPbox = nPbox;
    lv->CCH = sh.C / kPfNCH;
    lv->chunkBytes = lv->CCH * nP * esize;
    // pooling: CPL = 2 channels per lane share the mask-plane loads; ~5 units per lane and channel, every thread busy
    const int cpl = (sh.C % 2 == 0 && lv->CCH % 2 == 0) ? 2 : 1;
    int lpc = 1;
    while (lpc < 32 && (lpc * 5 < nU || sh.C * lpc < NT * cpl)) lpc <<= 1;
    lv->LPC = lpc;
    lv->CPL = cpl;
    // channel max/mean: a warp task = UL units x (32/UL) channel ranges, WS warps per unit group; choose the most even split over the warps
    const int NW = NT / 32;
    int bestUL = 8, bestWS = 1;
    double bestCost = 1e30;
    for (int ul = 8; ul <= 32; ul <<= 1) {
        const int nUG = (nU + ul - 1) / ul, cq = 32 / ul;
        for (int ws = 1; ws <= 16; ++ws) {
            if (ws * 3 * nP * 4 > 12 * 1024 || ws * cq > sh.C) break;
            const int tasks = nUG * ws, rounds = (tasks + NW - 1) / NW;
            const double cost = rounds * ((double)((sh.C + cq * ws - 1) / (cq * ws)) + 6.0) + 0.5 * ws;  // channels per lane + merge overhead
            if (cost < bestCost) { bestCost = cost; bestUL = ul; bestWS = ws; }
        }
    }
    lv->UL = bestUL;
    lv->WS = bestWS;
    int G = std::max(1, std::min(NT / nU, lv->CCH));  // rescale: channel groups; must divide the channels of a chunk
    while (lv->CCH % G) --G;
    lv->G = G;
    lv->TWp = sh.W + 8;
    lv->planeT = ((rowsPer + kMaxK - 1) * lv->TWp + 31) & ~31;
    return true;
}

static PfFwdOff pf_fwd_off(const PfLevel* lv, int n) {
    int C = 0, Hd = 0, nP = 0, mrawF = 0, Rf = 0;
    for (int i = 0; i < n; ++i) {
        C = std::max(C, lv[i].sh.C); Hd = std::max(Hd, lv[i].sh.hidden); nP = std::max(nP, lv[i].nP);
        mrawF = std::max(mrawF, (lv[i].rowsPer + kMaxK - 1) * lv[i].sh.W);
        const int stage = 4 * lv[i].sh.C * (lv[i].LPC | 1) + 8;
        const int tilemg = 3 * lv[i].planeT + std::max(lv[i].WS, 1) * 3 * lv[i].nP;
        Rf = std::max(Rf, std::max(stage, tilemg));
    }
    PfFwdOff o;
    int p = 0;
    auto take = [&p](int n) { const int r = p; p += (n + 3) & ~3; return r; };
    o.bar = take(2 * (kPfNCH + 1) + 2);
    o.red = take(64);
    o.wk = take(3 * kMaxK * kMaxK);
    o.avg = take(C); o.mx = take(C); o.ha = take(Hd); o.hm = take(Hd);
    o.q2 = take(2 * C); o.AB4 = take(4 * C); o.pst = take(4 * C);
    o.mloc = take(nP); o.mb = take(nP); o.aloc = take(nP);
    o.mraw = take(mrawF);
    p = (p + 31) & ~31;
    o.R = take(Rf);
    o.total = p;
    return o;
}

// x (or g) of one level as a 2-D tensor: dim0 = the S pixels of a plane, dim1 = the B*C planes; box = (nPbox, CCH)
static bool pf_make_map(CUtensorMap* out, const void* base, const Shape& sh, int dtype, const PfLevel& lv) {
    EncodeTiledFn enc = tensor_map_encoder();
    const int esize = dtype == MGA_F32 ? 4 : 2;
    if (!enc || (reinterpret_cast<uintptr_t>(base) & 15) || ((size_t)sh.S * esize) % 16) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)sh.S, (cuuint64_t)sh.B * sh.C};
    const cuuint64_t strides[1] = {(cuuint64_t)sh.S * esize};
    const cuuint32_t box[2] = {(cuuint32_t)lv.nPbox, (cuuint32_t)lv.CCH};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = dtype == MGA_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (dtype == MGA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    return enc(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct PfPlan { PfParams P; int smem; int NT; };

// Geometry of a call of n levels (same element type): team sizes, items, shared-memory layout.  false -> some level does not fit.
static bool pf_plan(const PfCall* calls, int n, bool bwd, PfPlan* plan) {
    static const int enabled = env_int("MGA_PF", 0);  // opt-in
    static const int max_st = env_int("MGA_PF_MAXST", 64);
    if (!enabled || n < 1 || n > kPfMaxLv) return false;
    PfParams& P = plan->P;
    std::memset(&P, 0, sizeof(P));
    const int esize = calls[0].dtype == MGA_F32 ? 4 : 2;
    int ST = 1;
    for (int i = 0; i < n; ++i) {
        if (calls[i].dtype != calls[0].dtype) return false;
        if (!pf_level_geom(calls[i].sh, esize, bwd, 0, &P.lv[i])) return false;
        ST = std::max(ST, P.lv[i].TS);
    }
    for (int i = 0; i < n; ++i) {  // every team size must divide the super-team: grow the smaller teams to the next divisor
        int ts = P.lv[i].TS;
        while (ts <= ST && (ST % ts || !pf_level_geom(calls[i].sh, esize, bwd, ts, &P.lv[i]))) ++ts;
        if (ts > ST) return false;
    }
    if (ST > max_st) return false;
    P.n_levels = n;
    P.ST = ST;
    int xbytes = 0, items = 0;
    for (int i = 0; i < n; ++i) {
        PfLevel& lv = P.lv[i];
        lv.spi = ST / lv.TS;
        lv.first_item = items;
        lv.n_items = (lv.sh.B + lv.spi - 1) / lv.spi;
        items += lv.n_items;
        xbytes = std::max(xbytes, lv.sh.C * lv.nP * esize);
    }
    P.total_items = items;
    P.xbytes = (xbytes + 127) & ~127;
    plan->NT = bwd ? kPfNTB : kPfNTF;
    static const int stagger = env_int("MGA_PF_STAGGER_NS", 8000);
    P.stagger_ns = stagger;
    P.fo = pf_fwd_off(P.lv, n);
    plan->smem = P.xbytes + 4 * P.fo.total;
    return plan->smem <= kClTwoCtaSmem;
}

template <typename Kern>
static int pf_launch(const char* name, Kern kernel, PfPlan& plan, const PfMaps& maps, cudaStream_t st) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail(MGA_ERR_CUDA, "%s: cudaGetDevice failed", name);
    struct Cfg { const void* k; int dev, smem, ctas; };
    static thread_local Cfg cache[32] = {};
    int ctas = -1;
    for (const Cfg& c : cache)
        if (c.k == (const void*)kernel && c.dev == dev && c.smem == plan.smem) ctas = c.ctas;
    if (ctas < 0) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        int occ = 0, sms = 0;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, plan.NT, (size_t)plan.smem);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) { cudaGetLastError(); return MGA_ERR_UNSUPPORTED; }
        ctas = std::min(occ * sms, kPfMaxCtas);
        for (Cfg& c : cache)
            if (!c.k) { c = Cfg{(const void*)kernel, dev, plan.smem, ctas}; break; }
    }
    PfParams& P = plan.P;
    P.nST = std::min(ctas / P.ST, P.total_items);
    if (P.nST < 1) return MGA_ERR_UNSUPPORTED;
    // team counters start at zero: one memset node per launch (all levels' counters live in level 0's scratch)
    if (cudaMemsetAsync(P.lv[0].ctr, 0, sizeof(unsigned) * kPfMaxLv * kPfMaxCtas, st) != cudaSuccess) return fail(MGA_ERR_CUDA, "%s: memset failed", name);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(P.nST * P.ST));
    cfg.blockDim = dim3((unsigned)plan.NT);
    cfg.dynamicSmemBytes = (size_t)plan.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;  // every CTA of the grid is co-resident: the team barriers cannot deadlock
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e;
    {
        LaunchScope ls(name, st);
        e = cudaLaunchKernelEx(&cfg, kernel, P, maps);
    }
    if (e != cudaSuccess) return fail(MGA_ERR_CUDA, "%s: launch (%d CTAs, %d B smem): %s", name, P.nST * P.ST, plan.smem, cudaGetErrorString(e));
    return MGA_OK;
}

// MGA_ERR_UNSUPPORTED: not a persistent shape (the caller takes the cluster / per-phase path)
static int pf_forward(const PfCall* calls, int n, cudaStream_t st) {
    PfPlan plan;
    if (!pf_plan(calls, n, false, &plan)) return MGA_ERR_UNSUPPORTED;
    PfMaps maps;
    for (int i = 0; i < n; ++i) {
        PfLevel& lv = plan.P.lv[i];
        const PfCall& c = calls[i];
        if ((reinterpret_cast<uintptr_t>(c.out) & 15) || !pf_make_map(&maps.x[i], c.x, c.sh, c.dtype, lv)) return MGA_ERR_UNSUPPORTED;
        if (c.sh.has_mask() && ((reinterpret_cast<uintptr_t>(c.mask) & 15) || (c.sh.W * (c.mdt == MGA_F32 ? 4 : 2)) % 16)) return MGA_ERR_UNSUPPORTED;  // bulk copies of mask rows
        lv.x = c.x; lv.mask = c.mask; lv.out = c.out; lv.mdt = c.mdt;
        lv.prm = c.prm; lv.ctx = c.ctx;
        lv.xchg = c.ps.xchg; lv.xstride = c.ps.xstride;
        lv.ctr = calls[0].ps.ctr + (size_t)i * kPfMaxCtas;
    }
    switch (calls[0].dtype) {
        case MGA_F32: return pf_launch("pf_fwd", pf_fwd_kernel<float>, plan, maps, st);
        case MGA_BF16: return pf_launch("pf_fwd", pf_fwd_kernel<__nv_bfloat16>, plan, maps, st);
        default: return pf_launch("pf_fwd", pf_fwd_kernel<__half>, plan, maps, st);
    }
}

#endif  // MGA_TUNING

// ------------------------------------------------------------------ forward
template <typename T, int VEC>
static int forward_split(const Shape& sh, const T* x, const void* mask, int mask_dtype, const mga_cbam_params& p, T* out, Ctx ctx,
                         FwdScratch fs, cudaStream_t st) {
    const int U = sh.S / VEC;
    const int nMaskTiles = (sh.S + kMaskTile - 1) / kMaskTile;
    const dim3 gmask(nMaskTiles, sh.B);
    if (sh.has_mask()) {
        MGA_LAUNCH("mask_prep", st,
            if (mask_dtype == MGA_F32) mask_prep_kernel<float><<<gmask, kBlock, 0, st>>>(static_cast<const float*>(mask), sh, ctx, fs);
            else if (mask_dtype == MGA_BF16) mask_prep_kernel<__nv_bfloat16><<<gmask, kBlock, 0, st>>>(static_cast<const __nv_bfloat16*>(mask), sh, ctx, fs);
            else mask_prep_kernel<__half><<<gmask, kBlock, 0, st>>>(static_cast<const __half*>(mask), sh, ctx, fs));
    }
    const int planes = sh.B * sh.C;
    MGA_LAUNCH("cam_pool", st,
        const int tpp = pool_tpp(U, VEC);
        if (VEC > 1 && sh.C % 4 == 0 && U >= kBlock && sh.B <= 65535) cam_pool4_kernel<T, VEC><<<dim3(sh.C / 4, sh.B), kBlock, 0, st>>>(x, sh, ctx, fs);
        else if (tpp == 256) cam_pool_kernel<T, VEC, 256><<<planes, kBlock, 0, st>>>(x, sh, ctx, fs);
        else if (tpp == 128) cam_pool_kernel<T, VEC, 128><<<(planes + 1) / 2, kBlock, 0, st>>>(x, sh, ctx, fs);
        else if (tpp == 64) cam_pool_kernel<T, VEC, 64><<<(planes + 3) / 4, kBlock, 0, st>>>(x, sh, ctx, fs);
        else cam_pool_kernel<T, VEC, 32><<<(planes + 7) / 8, kBlock, 0, st>>>(x, sh, ctx, fs));
    const size_t mlp_smem = (2 * (size_t)sh.C + 2 * sh.hidden) * sizeof(float);
    MGA_LAUNCH("cam_mlp", st, cam_mlp_kernel<<<sh.B, kBlock, mlp_smem, st>>>(sh, p, ctx, fs, nMaskTiles));
    {
        const TileCfg tc = pick_tiles(sh, U, VEC);
        const dim3 gtile((U + tc.lpt * tc.upt - 1) / (tc.lpt * tc.upt), sh.B);
#define MGA_CALL(L, P) MGA_LAUNCH("sam_reduce", st, (sam_reduce_kernel<T, VEC, L, P><<<gtile, kBlock, 0, st>>>(x, sh, ctx)))
        MGA_TILE_DISPATCH(tc, MGA_CALL);
#undef MGA_CALL
    }
    if (sh.W % 4 == 0) {
        ConvGeom cg = conv_geom(sh.W);
        PlaneMaps maps;
        const float* const planes[3] = {ctx.pmax, ctx.pavg, sh.has_mask() ? ctx.m : nullptr};
        if (!make_plane_maps(&maps, planes, sh, cg)) cg.use_tma = 0;
        const size_t smem = ((size_t)3 * cg.planeT + (size_t)3 * cg.nStrips * 4 + 3 * kMaxK * kMaxK) * sizeof(float);
        allow_big_smem(sam_conv4_kernel, smem);
        const dim3 gconv((sh.H + cg.RB - 1) / cg.RB, sh.B);
        MGA_LAUNCH("sam_conv", st, sam_conv4_kernel<<<gconv, kBlock, smem, st>>>(sh, p.wsam, ctx, cg, maps));
    } else {
        const dim3 gconv((sh.W + kConvTW - 1) / kConvTW, (sh.H + kConvTH - 1) / kConvTH, sh.B);
        MGA_LAUNCH("sam_conv", st, sam_conv_kernel<<<gconv, kBlock, 0, st>>>(sh, p.wsam, ctx));
    }
    if (!sh.gates_only()) {
        const size_t total = (size_t)planes * U;
        const int grid = (int)std::min<size_t>((total + kBlock - 1) / kBlock, (size_t)kSMs * 32);
        MGA_LAUNCH("rescale", st, (rescale_kernel<T, VEC><<<grid, kBlock, 0, st>>>(x, out, sh, ctx)));
    }
    return check_launch("mga_cbam_forward");
}

template <typename T>
static int forward_t(const Shape& sh, const mga_cbam_desc* d, const void* x, const void* mask, const mga_cbam_params& p, void* out,
                     Ctx ctx, FwdScratch fs, const PfScratch& ps, cudaStream_t st) {
    const int vec = pick_vec(sh, d->dtype, {x, out});
#ifdef MGA_TUNING
    if (vec > 1 && !(d->flags & (MGA_FORCE_SPLIT | MGA_NO_PERSIST))) {
        PfCall c{};
        c.sh = sh; c.dtype = d->dtype; c.mdt = d->mask_dtype; c.x = x; c.mask = mask; c.out = out; c.prm = p; c.ctx = ctx; c.ps = ps;
        const int rc = pf_forward(&c, 1, st);
        if (rc != MGA_ERR_UNSUPPORTED) return rc;
    }
#else
    (void)ps;
#endif
    ClGeom cgm;
    if (vec > 1 && !(d->flags & MGA_FORCE_SPLIT) && cl_geometry(sh, (int)sizeof(T), false, &cgm)) {
        const int rc = launch_cl("cl_fwd", cl_fwd_kernel<T>, cgm, sh.B, st, static_cast<const T*>(x), mask, d->mask_dtype, static_cast<T*>(out), sh, p, ctx,
                                 cl_fwd_off(sh.C, sh.hidden, cgm));
        if (rc != MGA_ERR_UNSUPPORTED) return rc;  // no cluster of that shape fits this device: one kernel per phase
    }
    if (vec == 1) return forward_split<T, 1>(sh, static_cast<const T*>(x), mask, d->mask_dtype, p, static_cast<T*>(out), ctx, fs, st);
    return forward_split<T, VecOf<T>::V>(sh, static_cast<const T*>(x), mask, d->mask_dtype, p, static_cast<T*>(out), ctx, fs, st);
}

// ------------------------------------------------------------------ backward
template <typename T, int VEC>
static int backward_split(const Shape& sh, const T* x, const void* mask, int mask_dtype, const T* g, const mga_cbam_params& p, Ctx ctx, T* dx,
                          void* dmask, const mga_cbam_grads& gp, BwdScratch bs, cudaStream_t st) {
    const int U = sh.S / VEC;
    const bool gates = sh.gates_only();  // dL/da is already in bs.T and dL/ds in bs.epart (one "tile"): no first reduce
    const TileCfg tc = pick_tiles(sh, U, VEC);
    const TileCfg tc1 = {tc.lpt, 1};                       // x-only reduce and dx: one unit per lane (measured faster)
    const int nT = gates ? 1 : (U + tc.lpt * tc.upt - 1) / (tc.lpt * tc.upt);  // tiles (= per-tile partial rows) of bwd_reduce1
    const int nT2 = (U + tc.lpt - 1) / tc.lpt;             // tiles of bwd_reduce2 and bwd_dx
    const dim3 gred(nT, sh.B);
    const dim3 gtile(nT2, sh.B);
    if (!gates) {
#define MGA_CALL(L, P) MGA_LAUNCH("bwd_reduce1", st, (bwd_reduce1_kernel<T, VEC, L, P><<<gred, kBlock, 0, st>>>(x, g, sh, ctx, bs, nT)))
        MGA_TILE_DISPATCH(tc, MGA_CALL);
#undef MGA_CALL
    }
    int nconv;
    if (sh.W % 4 == 0) {
        ConvGeom cg = conv_geom(sh.W);
        PlaneMaps maps;
        const float* const planes[3] = {ctx.pmax, ctx.pavg, sh.has_mask() ? ctx.m : nullptr};
        if (!make_plane_maps(&maps, planes, sh, cg)) cg.use_tma = 0;
        const size_t smem = ((size_t)4 * cg.planeT + 3 * kMaxK * kMaxK + 21 * 12 * kMaxK) * sizeof(float);
        allow_big_smem(bwd_conv4_kernel, smem);
        const dim3 gconv((sh.H + cg.RB - 1) / cg.RB, sh.B);
        nconv = gconv.x * gconv.y;
        MGA_LAUNCH("bwd_conv", st, bwd_conv4_kernel<<<gconv, kBlock, smem, st>>>(sh, p.wsam, ctx, bs, cg, maps));
    } else {
        const dim3 gconv((sh.W + kBT_W - 1) / kBT_W, (sh.H + kBT_H - 1) / kBT_H, sh.B);
        nconv = gconv.x * gconv.y * gconv.z;
        MGA_LAUNCH("bwd_conv", st, bwd_conv_kernel<<<gconv, kBlock, 0, st>>>(sh, p.wsam, ctx, bs));
    }
    if (!sh.samcam_add()) {
#define MGA_CALL(L, P) MGA_LAUNCH("bwd_reduce2", st, (bwd_reduce2_kernel<T, VEC, L, P><<<gtile, kBlock, 0, st>>>(x, sh, ctx, bs, nT2)))
        MGA_TILE_DISPATCH(tc1, MGA_CALL);
#undef MGA_CALL
    }
    const size_t mlp_smem = ((size_t)sh.C + 2 * sh.hidden + 3 * kBlock) * sizeof(float);
    if (!gates && std::max(nT, nT2) > 8) {
        // large planes: many per-tile partial rows per channel -- sum them with a grid of their own first (one CTA per 32 channels and
        // sample) instead of inside the one-CTA-per-sample MLP kernel
        const dim3 gsum((sh.C + 31) / 32, sh.B);
        MGA_LAUNCH("bwd_partsum", st, bwd_partsum_kernel<<<gsum, kBlock, 0, st>>>(sh, bs, nT, sh.samcam_add() ? 0 : nT2));
        BwdScratch b2 = bs;
        const size_t BC = (size_t)sh.B * sh.C;
        b2.epart = bs.psum; b2.gxpart = bs.psum + BC; b2.qpart = bs.psum + 2 * BC;
        MGA_LAUNCH("bwd_mlp", st, bwd_mlp_kernel<<<sh.B, kBlock, mlp_smem, st>>>(sh, p, ctx, b2, 1, 1));
    } else {
        MGA_LAUNCH("bwd_mlp", st, bwd_mlp_kernel<<<sh.B, kBlock, mlp_smem, st>>>(sh, p, ctx, bs, nT, nT2));
    }
#define MGA_CALL(L, P) MGA_LAUNCH("bwd_dx", st, (bwd_dx_kernel<T, VEC, L><<<gtile, kBlock, 0, st>>>(x, g, mask, mask_dtype, dx, dmask, sh, ctx, bs)))
    MGA_TILE_DISPATCH(tc1, MGA_CALL);
#undef MGA_CALL
    const int nw = 2 * sh.C * sh.hidden + sh.C + sh.hidden;
    const int nMlpBlocks = (nw + 31) / 32;  // 32 elements per block
    MGA_LAUNCH("bwd_wgrad", st, bwd_wgrad_kernel<<<nMlpBlocks + 3 * sh.k * sh.k + 1, kBlock, 0, st>>>(sh, ctx, bs, gp, nconv, nMlpBlocks, sh.B));
    return check_launch("mga_cbam_backward");
}

template <typename T>
static int backward_t(const Shape& sh, const mga_cbam_desc* d, const void* x, const void* mask, const void* g, const mga_cbam_params& p,
                      Ctx ctx, void* dx, void* dmask, const mga_cbam_grads& gp, BwdScratch bs, cudaStream_t st) {
    const int vec = pick_vec(sh, d->dtype, {x, g, dx});
    ClGeom cgm;
    if (vec > 1 && !(d->flags & MGA_FORCE_SPLIT) && cl_geometry(sh, (int)sizeof(T), true, &cgm)) {
        PlaneMaps maps;
        const float* const planes[3] = {ctx.pmax, ctx.pavg, sh.has_mask() ? ctx.m : nullptr};
        cgm.use_tma = make_plane_maps_box(&maps, planes, sh, cgm.TWp, cgm.tileRows) ? 1 : 0;
        if (!cgm.use_tma) std::memset(&maps, 0, sizeof(maps));
        const int rc = launch_cl("cl_bwd", cl_bwd_kernel<T>, cgm, sh.B, st, static_cast<const T*>(x), static_cast<const T*>(g), mask, d->mask_dtype,
                                 static_cast<T*>(dx), dmask, sh, p, ctx, bs, maps, cl_bwd_off(sh.C, sh.hidden, cgm));
        if (rc != MGA_OK && rc != MGA_ERR_UNSUPPORTED) return rc;
        if (rc == MGA_OK) {
        const int nw = 2 * sh.C * sh.hidden + sh.C + sh.hidden;
        const int nMlpBlocks = (nw + 31) / 32;
        MGA_LAUNCH("bwd_wgrad", st, bwd_wgrad_kernel<<<nMlpBlocks + 3 * sh.k * sh.k + 1, kBlock, 0, st>>>(sh, ctx, bs, gp, sh.B * cgm.CS, nMlpBlocks, sh.B));
        return check_launch("mga_cbam_backward(cluster)");
        }
    }
    if (vec == 1)
        return backward_split<T, 1>(sh, static_cast<const T*>(x), mask, d->mask_dtype, static_cast<const T*>(g), p, ctx, static_cast<T*>(dx),
                                    dmask, gp, bs, st);
    return backward_split<T, VecOf<T>::V>(sh, static_cast<const T*>(x), mask, d->mask_dtype, static_cast<const T*>(g), p, ctx,
                                          static_cast<T*>(dx), dmask, gp, bs, st);
}


// ------------------------------------------------------------------ fused concat forward (cbam_concat.cuh)
static bool concat_map(CUtensorMap* out, const void* base, CUtensorMapDataType dt, uint64_t inner, uint64_t rows, uint32_t box_inner, uint32_t box_rows) {
    // The driver entry point needs a current context on the CALLING thread.  A thread that has made no runtime call yet (autograd's
    // backward thread when every allocation came from the caching allocator) has none: bind the primary context first.
    static thread_local bool bound = false;
    if (!bound) { cudaFree(nullptr); bound = true; }
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc || (reinterpret_cast<uintptr_t>(base) & 15) || (inner * 2) % 16) return false;
    const cuuint64_t dims[2] = {inner, rows};
    const cuuint64_t strides[1] = {inner * 2};
    const cuuint32_t box[2] = {box_inner, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// launch of the GEMM kernel pair (MODE 0: forward, 1: backward): resident-weight persistent kernel for C <= 256, streaming kernel otherwise
template <typename T, int MODE>
static int concat_launch(const mga_cbam_desc* d, const ConcatMaps& maps, const T* epi_feat, const float* a, const float* bias, const float* beta,
                         T* out, const ConcatBwd& bw, cudaStream_t st, const char* what) {
    const int B = d->B, C = d->C, S = d->H * d->W;
    const int pm = (d->flags & MGA_PYRAMID_MULTIPLY) ? 1 : 0;
    static thread_local int configured_dev[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
    int dev = 0;
    cudaGetDevice(&dev);
    bool done = false;
    for (int v : configured_dev) done |= (v == dev);
    if (!done) {
        if (cudaFuncSetAttribute(concat_fwd_kernel<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCcSmemBytes) != cudaSuccess)
            return fail(MGA_ERR_CUDA, "%s: cudaFuncSetAttribute", what);
        for (int& v : configured_dev)
            if (v < 0) { v = dev; break; }
    }
    static const int use_res = env_int("MGA_CONCAT_RES", 1);
    if (use_res && concat_res_smem(C) <= kSmemLimit) {  // C <= 256: resident weights, persistent CTAs, double-buffered accumulators
        static thread_local int res_dev[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
        bool rdone = false;
        for (int v : res_dev) rdone |= (v == dev);
        if (!rdone) {
            if (cudaFuncSetAttribute(concat_fwd_res_kernel<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit) != cudaSuccess)
                return fail(MGA_ERR_CUDA, "%s: cudaFuncSetAttribute", what);
            for (int& v : res_dev)
                if (v < 0) { v = dev; break; }
        }
        const int nTilesN = (S + kCcBN - 1) / kCcBN, nMt = C / kCcBM;
        // pixel tiles per item: the fewest items that still give every SM >= ~4 of them (weights are reloaded per item)
        int chunks = 1;
        while (chunks < nTilesN && (long)B * nMt * chunks < 4L * kSMs) ++chunks;
        const int tpi = (nTilesN + chunks - 1) / chunks;
        const int items = B * nMt * ((nTilesN + tpi - 1) / tpi);
        MGA_LAUNCH(MODE == 0 ? "concat_fwd" : "concat_bwd_dx", st, (concat_fwd_res_kernel<T, MODE><<<std::min(items, kSMs), kCcThreads, concat_res_smem(C), st>>>(
            maps, epi_feat, a, bias, beta, out, B, C, S, tpi, pm, bw)));
        return check_launch(what);
    }
    const dim3 grid((S + kCcBN - 1) / kCcBN, C / kCcBM, B);
    MGA_LAUNCH(MODE == 0 ? "concat_fwd" : "concat_bwd_dx", st, (concat_fwd_kernel<T, MODE><<<grid, kCcThreads, kCcSmemBytes, st>>>(maps, epi_feat, a, bias, beta, out, C, S, pm, bw)));
    return check_launch(what);
}

template <typename T>
static int concat_forward_t(const mga_cbam_desc* d, const void* x, const float* s, const float* a, const float* w, const float* bias, const float* beta,
                            void* out, void* wscratch, cudaStream_t st) {
    const int B = d->B, C = d->C, S = d->H * d->W;
    T* wa = static_cast<T*>(wscratch);
    T* wb = wa + (size_t)B * C * C;
    const CUtensorMapDataType dt = std::is_same<T, __nv_bfloat16>::value ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    ConcatMaps maps;
    if (!concat_map(&maps.wa, wa, dt, C, (uint64_t)B * C, kCcBK, kCcBM) || !concat_map(&maps.wb, wb, dt, C, C, kCcBK, kCcBM) ||
        !concat_map(&maps.x, x, dt, S, (uint64_t)B * C, 64, kCcBK))
        return fail(MGA_ERR_UNSUPPORTED, "mga_cbam_concat_forward: tensor maps (16-byte aligned pointers, H*W %% 8 == 0)");
    MGA_LAUNCH("concat_fold", st, (concat_fold_kernel<T><<<dim3((C + kWarpsPerBlock - 1) / kWarpsPerBlock, B + 1), kBlock, 0, st>>>(w, s, wa, wb, B, C)));
    return concat_launch<T, 0>(d, maps, static_cast<const T*>(x), a, bias, beta, static_cast<T*>(out), ConcatBwd{}, st, "mga_cbam_concat_forward");
}

// backward: U = Wa^T g, V = Wb^T g on the tensor cores, the rest of the closed form in the epilogue (cbam_concat.cuh, MODE 1)
template <typename T>
static int concat_backward_dx_t(const mga_cbam_desc* d, const void* x, const void* g, const float* s, const float* a, const float* w, const float* bias,
                                const float* beta, void* dx, void* ga, float* ds_part, float* db_part, float* da_part, float* dal_part, void* wscratch,
                                cudaStream_t st) {
    const int B = d->B, C = d->C, S = d->H * d->W;
    T* wat = static_cast<T*>(wscratch);
    T* wbt = wat + (size_t)C * C;
    const CUtensorMapDataType dt = std::is_same<T, __nv_bfloat16>::value ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    ConcatMaps maps;
    const bool m1 = concat_map(&maps.wa, wat, dt, C, C, kCcBK, kCcBM), m2 = concat_map(&maps.wb, wbt, dt, C, C, kCcBK, kCcBM),
               m3 = concat_map(&maps.x, g, dt, S, (uint64_t)B * C, 64, kCcBK);
    if (!m1 || !m2 || !m3)
        return fail(MGA_ERR_UNSUPPORTED, "mga_cbam_concat_backward_dx: tensor maps (16-byte aligned pointers, H*W %% 8 == 0) [wa %d wb %d g %d; g=%p ws=%p]",
                    (int)m1, (int)m2, (int)m3, g, wscratch);
    MGA_LAUNCH("concat_transpose", st, (concat_transpose_kernel<T><<<(C * C + kBlock - 1) / kBlock, kBlock, 0, st>>>(w, wat, wbt, C)));
    const ConcatBwd bw{x, s, ga, ds_part, db_part, da_part, dal_part};
    return concat_launch<T, 1>(d, maps, static_cast<const T*>(g), a, bias, beta, static_cast<T*>(dx), bw, st, "mga_cbam_concat_backward_dx");
}

}  // namespace mga

using namespace mga;

extern "C" {

int mga_abi_version(void) { return MGA_ABI_VERSION; }
const char* mga_last_error(void) { return g_err; }

unsigned long long mga_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

/* debug (tools/timeline.py): device buffer of 16 uint64 per CTA that the cluster kernels stamp with %globaltimer per phase (NULL = off).
   Only -DMGA_TUNING builds carry the stamps; the product library answers MGA_ERR_UNSUPPORTED. */
int mga_debug_timeline(void* device_buffer) {
#ifdef MGA_TUNING
    unsigned long long* p = static_cast<unsigned long long*>(device_buffer);
    const cudaError_t e = cudaMemcpyToSymbol(g_timeline, &p, sizeof(p));
    return e == cudaSuccess ? MGA_OK : fail(MGA_ERR_CUDA, "mga_debug_timeline: %s", cudaGetErrorString(e));
#else
    (void)device_buffer;
    return fail(MGA_ERR_UNSUPPORTED, "mga_debug_timeline: library built without -DMGA_TUNING");
#endif
}

int mga_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_on.store(on != 0);
    return MGA_OK;
}
int mga_profile_count(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    return (int)g_prof.size();
}
int mga_profile_read(int i, const char** name, float* ms) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (i < 0 || i >= (int)g_prof.size() || !name || !ms) return fail(MGA_ERR_ARG, "bad profile record index");
    *name = g_prof[i].name;
    const cudaError_t e = cudaEventElapsedTime(ms, g_prof[i].a, g_prof[i].b);
    if (e != cudaSuccess) return fail(MGA_ERR_CUDA, "profile record %d: %s", i, cudaGetErrorString(e));
    return MGA_OK;
}

int mga_cbam_plan(const mga_cbam_desc* d, int direction, mga_cbam_plan_info* info) {
    Shape sh;
    if (int rc = validate(d, &sh)) return rc;
    if (!info || (direction != 0 && direction != 1)) return fail(MGA_ERR_ARG, "mga_cbam_plan: bad argument");
    const int esize = d->dtype == MGA_F32 ? 4 : 2;
    const bool bwd = direction == 1;
    ClGeom g{};
    std::memset(info, 0, sizeof(*info));
    const bool vec_ok = sh.S % (16 / esize) == 0;  // (pointer alignment is checked at call time)
    if (vec_ok && !(d->flags & MGA_FORCE_SPLIT) && cl_geometry(sh, esize, bwd, &g)) {
        info->path = 1;
        info->cluster_size = g.CS;
        info->rows_per_cta = g.rowsPer;
        info->threads = g.NT;
        info->smem_bytes = g.smem_bytes;
        info->launches = bwd ? 2 : 1;
        return MGA_OK;
    }
    info->path = 0;
    {
        const int vec = vec_ok ? 16 / esize : 1, U = sh.S / vec;
        const TileCfg tc = pick_tiles(sh, U, vec);
        const bool partsum = (U + tc.lpt - 1) / tc.lpt > 8;  // bwd_partsum_kernel (large planes)
        info->launches = bwd ? 5 + (sh.samcam_add() ? 0 : 1) + (partsum ? 1 : 0) : 5 + (sh.has_mask() ? 1 : 0);
    }
    return MGA_OK;
}

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
    PfScratch ps;
    carve_ctx(sh, ctx_buf, &ctx);
    carve_fwd(sh, scratch, &fs, &ps);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (d->dtype) {
        case MGA_F32: return forward_t<float>(sh, d, x, mask, *p, out, ctx, fs, ps, st);
        case MGA_BF16: return forward_t<__nv_bfloat16>(sh, d, x, mask, *p, out, ctx, fs, ps, st);
        default: return forward_t<__half>(sh, d, x, mask, *p, out, ctx, fs, ps, st);
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

int mga_cbam_gates_forward(const mga_cbam_desc* d, const void* x, const void* mask, const mga_cbam_params* p, void* ctx_buf, void* scratch,
                           void* stream) {
    if (!d) return fail(MGA_ERR_ARG, "null descriptor");
    mga_cbam_desc g = *d;
    g.flags |= MGA_GATES_ONLY | MGA_SAMCAM_ADD;  // the spatial gate is computed from x itself (a'); cluster kernel when the shape allows
    return mga_cbam_forward(&g, x, mask, p, ctx_buf /* `out` is never written in gates mode */, ctx_buf, scratch, stream);
}

int mga_cbam_gates_backward(const mga_cbam_desc* d, const void* x, const void* mask, const float* grad_s, const float* grad_a,
                            const mga_cbam_params* p, const void* ctx_buf, void* grad_x, void* grad_mask, const mga_cbam_grads* gp,
                            void* scratch, void* stream) {
    return mga_cbam_gates_backward_acc(d, x, mask, grad_s, grad_a, nullptr, p, ctx_buf, grad_x, grad_mask, gp, scratch, stream);
}

int mga_cbam_gates_backward_acc(const mga_cbam_desc* d, const void* x, const void* mask, const float* grad_s, const float* grad_a,
                                const void* grad_x_acc, const mga_cbam_params* p, const void* ctx_buf, void* grad_x, void* grad_mask,
                                const mga_cbam_grads* gp, void* scratch, void* stream) {
    if (!d) return fail(MGA_ERR_ARG, "null descriptor");
    if (!grad_s || !grad_a) return fail(MGA_ERR_ARG, "null gate gradient");
    mga_cbam_desc g = *d;
    g.flags |= MGA_GATES_ONLY | MGA_SAMCAM_ADD | MGA_FORCE_SPLIT;
    if (grad_x_acc) g.flags |= MGA_GATES_ACC;
    if (grad_x_acc && grad_x_acc == grad_x) return fail(MGA_ERR_ARG, "grad_x_acc must not alias grad_x");
    const void* gsrc = grad_x_acc ? grad_x_acc : x;  // read as grad_out by the last kernel: weight 1 (accumulate) or 0
    Shape sh;
    if (int rc = validate(&g, &sh)) return rc;
    if (!x || !p || !ctx_buf || !grad_x || !gp || !scratch) return fail(MGA_ERR_ARG, "null pointer argument");
    if (sh.has_mask() && !mask) return fail(MGA_ERR_ARG, "MGA_HAS_MASK set but mask is null");
    Ctx ctx;
    BwdScratch bs;
    carve_ctx(sh, const_cast<void*>(ctx_buf), &ctx);
    carve_bwd(sh, scratch, &bs);
    bs.T = const_cast<float*>(grad_a);      // (B,S): plays sum_c g x q_c with k1 = 1
    bs.epart = const_cast<float*>(grad_s);  // (B,1,C): the single "tile" of channel-gate gradients
    bs.gxpart = const_cast<float*>(grad_s); // read but unused (feeds d beta only)
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (g.dtype) {  // grad_out := x (multiplied by k0 = k1 = 0 inside bwd_dx) or the upstream gradient (k0 = 1)
        case MGA_F32: return backward_t<float>(sh, &g, x, mask, gsrc, *p, ctx, grad_x, grad_mask, *gp, bs, st);
        case MGA_BF16: return backward_t<__nv_bfloat16>(sh, &g, x, mask, gsrc, *p, ctx, grad_x, grad_mask, *gp, bs, st);
        default: return backward_t<__half>(sh, &g, x, mask, gsrc, *p, ctx, grad_x, grad_mask, *gp, bs, st);
    }
}


int mga_cbam_concat_forward(const mga_cbam_desc* d, const void* x, const float* s, const float* a, const float* w, const float* bias,
                            const float* beta, void* out, void* wscratch, void* stream) {
    if (!d || !x || !s || !a || !w || !bias || !beta || !out || !wscratch) return fail(MGA_ERR_ARG, "null pointer argument");
    if (d->B <= 0 || d->C <= 0 || d->H <= 0 || d->W <= 0 || d->B > 65535) return fail(MGA_ERR_ARG, "bad shape");
    if (d->dtype != MGA_BF16 && d->dtype != MGA_F16) return fail(MGA_ERR_UNSUPPORTED, "the tensor-core concat forward takes bf16 / f16 features");
    if (d->C % kCcBM || (d->H * d->W) % 8 || (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(a) & 15))
        return fail(MGA_ERR_UNSUPPORTED, "needs C %% 128 == 0, H*W %% 8 == 0 and 16-byte aligned tensors");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (d->dtype == MGA_BF16) return concat_forward_t<__nv_bfloat16>(d, x, s, a, w, bias, beta, out, wscratch, st);
    return concat_forward_t<__half>(d, x, s, a, w, bias, beta, out, wscratch, st);
}

int mga_cbam_concat_backward_elem(const mga_cbam_desc* d, const void* x, const void* grad_out, const void* uv, const float* s, const float* a,
                                  const float* bias, const float* beta, void* grad_x, void* ga, float* ds_part, float* dbias_part, float* grad_a,
                                  float* dalpha_part, void* stream) {
    if (!d || !x || !grad_out || !uv || !s || !a || !bias || !beta || !grad_x || !ga || !ds_part || !dbias_part || !grad_a || !dalpha_part)
        return fail(MGA_ERR_ARG, "null pointer argument");
    if (d->B <= 0 || d->C <= 0 || d->H <= 0 || d->W <= 0 || d->B > 65535) return fail(MGA_ERR_ARG, "bad shape");
    if (d->dtype != MGA_BF16 && d->dtype != MGA_F16) return fail(MGA_ERR_UNSUPPORTED, "16-bit features only");
    const int S = d->H * d->W;
    if (S % 8) return fail(MGA_ERR_UNSUPPORTED, "needs H*W %% 8 == 0");
    for (const void* p : {x, grad_out, uv, (const void*)grad_x, (const void*)ga, (const void*)a})
        if (reinterpret_cast<uintptr_t>(p) & 15) return fail(MGA_ERR_UNSUPPORTED, "16-byte aligned tensors only");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 grid((S / 8 + 31) / 32, d->B);
    const int pm = (d->flags & MGA_PYRAMID_MULTIPLY) ? 1 : 0;
    if (d->dtype == MGA_BF16) {
        using T = __nv_bfloat16;
        MGA_LAUNCH("concat_bwd_elem", st, (concat_bwd_elem_kernel<T><<<grid, kBlock, 0, st>>>(static_cast<const T*>(x), static_cast<const T*>(grad_out), static_cast<const T*>(uv), s, a, bias, beta,
            static_cast<T*>(grad_x), static_cast<T*>(ga), ds_part, dbias_part, grad_a, dalpha_part, d->C, S, pm)));
    } else {
        using T = __half;
        MGA_LAUNCH("concat_bwd_elem", st, (concat_bwd_elem_kernel<T><<<grid, kBlock, 0, st>>>(static_cast<const T*>(x), static_cast<const T*>(grad_out), static_cast<const T*>(uv), s, a, bias, beta,
            static_cast<T*>(grad_x), static_cast<T*>(ga), ds_part, dbias_part, grad_a, dalpha_part, d->C, S, pm)));
    }
    return check_launch("mga_cbam_concat_backward_elem");
}

int mga_cbam_concat_backward_dx(const mga_cbam_desc* d, const void* x, const void* grad_out, const float* s, const float* a, const float* w,
                                const float* bias, const float* beta, void* grad_x, void* ga, float* ds_part, float* dbias_part, float* da_part,
                                float* dalpha_part, void* wscratch, void* stream) {
    if (!d || !x || !grad_out || !s || !a || !w || !bias || !beta || !grad_x || !ga || !ds_part || !dbias_part || !da_part || !dalpha_part || !wscratch)
        return fail(MGA_ERR_ARG, "null pointer argument");
    if (d->B <= 0 || d->C <= 0 || d->H <= 0 || d->W <= 0 || d->B > 65535) return fail(MGA_ERR_ARG, "bad shape");
    if (d->dtype != MGA_BF16 && d->dtype != MGA_F16) return fail(MGA_ERR_UNSUPPORTED, "the tensor-core concat backward takes bf16 / f16 features");
    if (d->C % kCcBM || (d->H * d->W) % 8) return fail(MGA_ERR_UNSUPPORTED, "needs C %% 128 == 0 and H*W %% 8 == 0");
    for (const void* p : {x, grad_out, (const void*)grad_x, (const void*)ga, (const void*)a, (const void*)wscratch})
        if (reinterpret_cast<uintptr_t>(p) & 15) return fail(MGA_ERR_UNSUPPORTED, "16-byte aligned tensors only");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (d->dtype == MGA_BF16)
        return concat_backward_dx_t<__nv_bfloat16>(d, x, grad_out, s, a, w, bias, beta, grad_x, ga, ds_part, dbias_part, da_part, dalpha_part, wscratch, st);
    return concat_backward_dx_t<__half>(d, x, grad_out, s, a, w, bias, beta, grad_x, ga, ds_part, dbias_part, da_part, dalpha_part, wscratch, st);
}

int mga_cbam_concat_wgrad_reduce(const mga_cbam_desc* d, const void* Ga, const void* Gb, int gemm_dtype, const float* s, const float* beta,
                                 float* grad_w, void* stream) {
    if (!d || !Ga || !Gb || !s || !beta || !grad_w) return fail(MGA_ERR_ARG, "null pointer argument");
    if (d->B <= 0 || d->C <= 0 || (long long)d->C * d->C > (1LL << 30)) return fail(MGA_ERR_ARG, "bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int C = d->C, grid = (C * C + kBlock - 1) / kBlock;
    switch (gemm_dtype) {
        case MGA_F32:
            MGA_LAUNCH("concat_wgrad", st, (concat_wgrad_reduce_kernel<float><<<grid, kBlock, 0, st>>>(static_cast<const float*>(Ga), static_cast<const float*>(Gb), s, beta, grad_w, d->B, C)));
            break;
        case MGA_BF16:
            MGA_LAUNCH("concat_wgrad", st, (concat_wgrad_reduce_kernel<__nv_bfloat16><<<grid, kBlock, 0, st>>>(static_cast<const __nv_bfloat16*>(Ga), static_cast<const __nv_bfloat16*>(Gb), s, beta, grad_w, d->B, C)));
            break;
        case MGA_F16:
            MGA_LAUNCH("concat_wgrad", st, (concat_wgrad_reduce_kernel<__half><<<grid, kBlock, 0, st>>>(static_cast<const __half*>(Ga), static_cast<const __half*>(Gb), s, beta, grad_w, d->B, C)));
            break;
        default: return fail(MGA_ERR_ARG, "bad GEMM result dtype %d", gemm_dtype);
    }
    return check_launch("mga_cbam_concat_wgrad_reduce");
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

#include "next_ops.cuh"  // MaskECA, MGAMaskHead tail, ProbMaskGater sampling, collate
