// cbam_cluster.cuh -- cluster-per-sample fused forward / backward kernels (the default path for samples that fit).
//
// One thread-block CLUSTER owns one sample for the whole pass.  CTA r of the cluster owns image rows
// [r*rowsPer, (r+1)*rowsPer) of that sample for ALL channels, every phase of the pass runs inside the same kernel and
// phases are ordered by the hardware cluster barrier.  Nothing of the sample is staged in shared memory: the first
// pass over x (and g) comes from HBM, the later passes re-read the same lines a few microseconds later, while only
// (resident CTAs) x (~200 KB) ~ 60 MB of feature map are in flight on the whole GPU -- they are L2 hits.  HBM therefore
// sees the algorithmic traffic only: forward 2N (+ planes), backward 3N (+ planes), instead of 4N / 6N for the
// one-kernel-per-phase path (cbam_fwd.cuh / cbam_bwd.cuh), and the six small kernels per direction disappear.
//
// Cross-CTA data is tiny and travels through distributed shared memory (DSMEM):
//   * per-channel partial sums are PUSHED into every CTA's shared memory before a cluster barrier, so after the barrier
//     every CTA recomputes the tiny shared MLP (forward or backward) from local data;
//   * the 7x7 convolution needs 3 halo rows of the 3-plane map (forward) / of dpre (backward) from the neighbouring
//     CTAs' tiles: read through DSMEM after a cluster barrier.
//
// Two thread -> data mappings (both runtime-shaped, so one instantiation per element type):
//   T1 "channel rows": a group of LPT lanes owns one channel, lane = up to K units (16 B of pixels) of the CTA's pixel
//      range -> per-CHANNEL sums are thread-private and need one shuffle reduction per channel (pool, bwd reduce 1/2);
//   T2 "pixel columns": thread = (unit, channel group), loops over the channels of its group -> per-PIXEL sums / max are
//      thread-private (channel max/mean, rescale, dx); the G channel groups merge through shared memory.
//
// Reference lines: mga_yolo/nn/modules/masked_cbam.py:87-171 (forward); SURVEY.md section 8a "Backward" (closed form).
#pragma once
#include <cooperative_groups.h>

#include "cbam_conv.cuh"
#include "cbam_fused.cuh"  // load_mask_any / store_mask_any / stamp
#include "common.cuh"

namespace mga {
namespace cgx = cooperative_groups;

constexpr int kCT = 256;        // threads per CTA (== kBlock: the conv helpers stride by kBlock)
constexpr int kCW = kCT / 32;
static_assert(kCT == kBlock, "conv helpers assume kBlock threads");

#ifndef MGA_CL_HINTS
#define MGA_CL_HINTS 1  // L2 eviction-priority hints: first touch of x/g = keep, last read = evict first
#endif
#ifndef MGA_CL_MINB
#define MGA_CL_MINB 2   // resident CTAs per SM the register allocation must allow
#endif

struct ClGeom {
    int CS;        // CTAs per cluster (= per sample)
    int rowsPer;   // image rows per CTA
    int nUmax;     // 16-byte units per CTA (rowsPer * W / VEC)
    int nPmax;     // pixels per CTA
    int G;         // T2: channel groups = kCT / nUmax
    int lsh;       // T1: log2(lanes per channel)
    int tileRows;  // rowsPer + 6
    int TWp;       // W + 8
    int planeT;    // floats per tile plane (multiple of 32)
    int smem_bytes;
};

template <int VEC> struct ClK { static constexpr int K = (VEC == 4) ? 7 : 4; };  // T1 units per lane

// ---------------------------------------------------------------- shared-memory layouts (float offsets)
struct ClFwdOff { int wk, red, avg, mx, ha, hm, sA, sB, q, mypart, part, msum, mloc, aloc, tile, mg, total; };
__host__ __device__ inline ClFwdOff cl_fwd_off(int C, int Hd, const ClGeom& g) {
    ClFwdOff o;
    int p = 0;
    auto take = [&p](int n) { const int r = p; p += (n + 3) & ~3; return r; };
    o.wk = take(3 * kMaxK * kMaxK);
    o.red = take(64);
    o.avg = take(C); o.mx = take(C); o.ha = take(Hd); o.hm = take(Hd);
    o.sA = take(C); o.sB = take(C); o.q = take(C);
    o.mypart = take(4 * C);
    o.part = take(4 * C * g.CS);
    o.msum = take(16);
    o.mloc = take(g.nPmax);
    o.aloc = take(g.nPmax);
    o.tile = take(3 * g.planeT);
    o.mg = take(3 * g.G * g.nPmax);  // T2 merge buffers [max | sum | idx][G][nPmax]; later the conv's per-plane partials [3][nPmax]
    o.total = p;
    return o;
}

struct ClBwdOff {
    int wsm, red, s, q, chA, chB, amx, dz, eloc, qloc, binloc, dha, dhm, epart, qpart, binpart, gxpart;
    int aloc, ae, mloc, idx, pmx, d0, d1s, d2, dpre, cat, tp, dwp, binw, total;
};
__host__ __device__ inline ClBwdOff cl_bwd_off(int C, int Hd, const ClGeom& g) {
    ClBwdOff o;
    int p = 0;
    auto take = [&p](int n) { const int r = p; p += (n + 3) & ~3; return r; };
    o.wsm = take(3 * kMaxK * kMaxK);
    o.red = take(64);  // 32 doubles
    o.s = take(C); o.q = take(C);
    o.chA = take(4 * C); o.chB = take(2 * C); o.amx = take(C);
    o.dz = take(C); o.eloc = take(C); o.qloc = take(C); o.binloc = take(C);
    o.dha = take(Hd); o.dhm = take(Hd);
    o.epart = take(C * g.CS); o.qpart = take(C * g.CS); o.binpart = take(C * g.CS);
    o.gxpart = take(2 * 16);  // doubles
    o.aloc = take(g.nPmax); o.ae = take(g.nPmax); o.mloc = take(g.nPmax); o.idx = take(g.nPmax); o.pmx = take(g.nPmax);
    o.d0 = take(g.nPmax); o.d1s = take(g.nPmax); o.d2 = take(g.nPmax);
    o.dpre = take(g.planeT);
    o.cat = take(3 * g.planeT);
    const int tpn = (kCW > g.G ? kCW : g.G) * g.nPmax;
    o.tp = take(tpn);          // T partials of the 8 warps, later the R partials of the G channel groups
    o.dwp = take(21 * 12 * kMaxK);
    o.binw = take(kCW * C);
    o.total = p;
    return o;
}

// ---------------------------------------------------------------- 128-bit global access with optional L2 policy
enum { kPolNone = 0, kPolKeep = 1, kPolLast = 2 };
struct ClPol { unsigned long long keep, last; };
__device__ __forceinline__ ClPol cl_policies() {
    ClPol p;
#if MGA_CL_HINTS
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p.keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p.last));
#else
    p.keep = 0; p.last = 0;
#endif
    return p;
}
template <int POL>
__device__ __forceinline__ uint4 ldg128(const void* p, const ClPol& pol) {
    uint4 t;
#if MGA_CL_HINTS
    if constexpr (POL == kPolNone) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p));
    } else {
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p), "l"(POL == kPolKeep ? pol.keep : pol.last));
    }
#else
    (void)pol;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p));
#endif
    return t;
}
template <typename T, int VEC>
__device__ __forceinline__ void unpack(const uint4& t, float (&v)[VEC]) {
    if constexpr (sizeof(T) == 4) {
        v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
    } else {
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                v[2 * i] = __uint_as_float(w[i] << 16);
                v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
            } else {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
                v[2 * i] = f.x;
                v[2 * i + 1] = f.y;
            }
        }
    }
}
template <typename T, int VEC>
__device__ __forceinline__ uint4 pack(const float (&v)[VEC]) {
    uint4 t;
    if constexpr (sizeof(T) == 4) {
        t = make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
    } else {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                w[i] = *reinterpret_cast<uint32_t*>(&h);
            } else {
                __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                w[i] = *reinterpret_cast<uint32_t*>(&h);
            }
        }
        t = make_uint4(w[0], w[1], w[2], w[3]);
    }
    return t;
}
// VEC consecutive fp32 / int32 values from shared memory (16-byte aligned)
template <int VEC>
__device__ __forceinline__ void lds_f(const float* p, float (&v)[VEC]) {
#pragma unroll
    for (int i = 0; i < VEC / 4; ++i) {
        const float4 t = reinterpret_cast<const float4*>(p)[i];
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}
template <int VEC>
__device__ __forceinline__ void lds_i(const int* p, int (&v)[VEC]) {
#pragma unroll
    for (int i = 0; i < VEC / 4; ++i) {
        const int4 t = reinterpret_cast<const int4*>(p)[i];
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}

// block-wide fp64 sum over kCT threads (result in every thread)
__device__ __forceinline__ double cl_block_sum_d(double v, double* sh /* >= 32 */) { return block_sum_d(v, sh); }

// fetch the halo rows of `nplanes` tile planes from the CTAs that own them (DSMEM); own rows and padding stay untouched
__device__ __forceinline__ void cl_fetch_halo(cgx::cluster_group& cluster, float* tile, int nplanes, const ClGeom& gm, int y0, int rows, int H,
                                              int W) {
    const int cpr = W / 4;  // float4 chunks per image row
    const int per = 6 * cpr;
    for (int i = threadIdx.x; i < nplanes * per; i += kCT) {
        const int pl = i / per, rem = i - pl * per;
        const int hr = rem / cpr, ch = rem - hr * cpr;       // halo row 0..5: 3 above, 3 below
        const int tr = hr < 3 ? hr : rows + hr;              // tile row (own rows are 3 .. 3+rows-1)
        const int yy = y0 - 3 + tr;
        if (yy < 0 || yy >= H) continue;
        const int rr = yy / gm.rowsPer;
        const int rtr = yy - rr * gm.rowsPer + 3;
        const float* rt = cluster.map_shared_rank(tile, rr);
        const float4 v = *reinterpret_cast<const float4*>(rt + (size_t)pl * gm.planeT + rtr * gm.TWp + 4 + ch * 4);
        *reinterpret_cast<float4*>(tile + (size_t)pl * gm.planeT + tr * gm.TWp + 4 + ch * 4) = v;
    }
}

// ================================================================== forward
template <typename T>
__global__ void __launch_bounds__(kCT, MGA_CL_MINB) cl_fwd_kernel(const T* __restrict__ x, const void* __restrict__ mask, int mdt,
                                                                  T* __restrict__ out, Shape sh, mga_cbam_params prm, Ctx ctx, ClGeom gm) {
    constexpr int VEC = 16 / sizeof(T);
    constexpr int K = ClK<VEC>::K;
    extern __shared__ __align__(128) float clsm[];
    float* const csm = clsm;
    cgx::cluster_group cluster = cgx::this_cluster();
    const int CS = gm.CS;
    const int r = (int)cluster.block_rank();
    const int b = blockIdx.x / CS;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int C = sh.C, S = sh.S, Hd = sh.hidden, W = sh.W, H = sh.H;
    const int y0 = r * gm.rowsPer;
    const int rows = max(0, min(gm.rowsPer, H - y0));
    const int nP = rows * W, nU = nP / VEC, p0 = y0 * W;
    const int nPmax = gm.nPmax, nUmax = gm.nUmax, TWp = gm.TWp, planeT = gm.planeT;
    const bool has_mask = sh.has_mask();
    const ClFwdOff o = cl_fwd_off(C, Hd, gm);
    float* wk = csm + o.wk;
    float* red = csm + o.red;
    float* s_avg = csm + o.avg;
    float* s_mx = csm + o.mx;
    float* s_ha = csm + o.ha;
    float* s_hm = csm + o.hm;
    float* sA = csm + o.sA;
    float* sB = csm + o.sB;
    float* s_q = csm + o.q;
    float* mypart = csm + o.mypart;
    float* part = csm + o.part;
    float* msum_part = csm + o.msum;
    float* mloc = csm + o.mloc;
    float* aloc = csm + o.aloc;
    float* tile = csm + o.tile;  // [pmax | pavg | m][tileRows][TWp]
    float* mg = csm + o.mg;
    const ClPol pol = cl_policies();
    const size_t bS = (size_t)b * S;

    // ---- phase 0: zero the tile (padding), conv weights, mask -> m (own pixels), sum(m)
    stamp(0);
    cluster.barrier_arrive();  // #0 (waited before the first push): every CTA of the cluster has started
    for (int i = tid; i < 3 * planeT; i += kCT) tile[i] = 0.0f;
    load_weights7(prm.wsam, sh.k, false, wk);
    __syncthreads();
    float mtot = 0.0f;
    {
        float macc = 0.0f;
        for (int i = tid; i < nP; i += kCT) {
            float v = 1.0f;
            if (has_mask) {
                v = load_mask_any(mask, mdt, bS + p0 + i);
                if (sh.gate_clamp()) v = fminf(fmaxf(v, 0.0f), 1.0f);
                if (sh.sigmoid_mask()) v = sigmoidf_acc(v);
                ctx.m[bS + p0 + i] = v;
                const int ry = i / W, cx = i - ry * W;
                tile[2 * planeT + (3 + ry) * TWp + 4 + cx] = v;
                macc += v;
            }
            mloc[i] = v;
        }
        mtot = block_sum(macc, red);
    }
    __syncthreads();

    // ---- phase 1 (T1): per channel over the CTA's pixels: sum x*m, sum x, masked max + arg max
    stamp(1);
    const int lsh = gm.lsh, LPT = 1 << lsh, CPW = 32 >> lsh;
    const int sub = lane >> lsh, ul = lane & (LPT - 1);
    const T* xb = x + (size_t)b * C * S + p0;  // the CTA's pixel range of channel 0
    for (int cb = w * CPW; cb < C; cb += kCW * CPW) {
        const int c = cb + sub;
        const bool cok = c < C;
        uint4 raw[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int u = ul + (k << lsh);
            if (cok && u < nU) raw[k] = ldg128<kPolKeep>(xb + (size_t)c * S + (size_t)u * VEC, pol);
        }
        float sxm = 0.0f, sx = 0.0f, best = -INFINITY;
        int bidx = -1;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int u = ul + (k << lsh);
            if (cok && u < nU) {
                float v[VEC], mv[VEC];
                unpack<T, VEC>(raw[k], v);
                lds_f<VEC>(mloc + u * VEC, mv);
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    sx += v[i];
                    sxm = fmaf(v[i], mv[i], sxm);
                    if (mv[i] > 0.5f && v[i] > best) { best = v[i]; bidx = p0 + u * VEC + i; }
                }
            }
        }
        // larger value wins; on a tie the lower pixel index (first maximum in scan order)
        for (int of = LPT >> 1; of > 0; of >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, of);
            sxm += __shfl_xor_sync(0xffffffffu, sxm, of);
            const float ob = __shfl_xor_sync(0xffffffffu, best, of);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, of);
            if ((oi >= 0) && (bidx < 0 || ob > best || (ob == best && oi < bidx))) { best = ob; bidx = oi; }
        }
        if (ul == 0 && cok) {
            mypart[c] = sxm;
            mypart[C + c] = sx;
            mypart[2 * C + c] = best;
            mypart[3 * C + c] = __int_as_float(bidx);
        }
    }
    __syncthreads();
    cluster.barrier_wait();  // #0
    // push this CTA's partials into every CTA of the cluster: part[src rank][4][C]
    if (has_mask && tid < CS) cluster.map_shared_rank(msum_part, tid)[r] = mtot;
    for (int i = tid; i < 4 * C * CS; i += kCT) {
        const int rr = i / (4 * C), j = i - rr * 4 * C;
        cluster.map_shared_rank(part, rr)[r * 4 * C + j] = mypart[j];
    }
    const float beta = __ldg(prm.beta);
    stamp(2);
    cluster.sync();  // #1: partials of all ranks are in local shared memory

    // ---- phase 2: pooled descriptors, shared MLP, sigmoid (recomputed by every CTA)
    stamp(3);
    {
        float use = 0.0f, den = 1.0f;
        if (has_mask) {
            float tot = 0.0f;
            for (int rr = 0; rr < CS; ++rr) tot += msum_part[rr];
            use = (tot / (float)S >= sh.tiny_thr) ? 1.0f : 0.0f;
            den = fmaxf(tot, sh.eps);
            if (r == 0 && tid == 0) { ctx.msum[b] = tot; ctx.use[b] = use; ctx.den[b] = den; }
        }
        const float invS = 1.0f / (float)S;
        for (int c = tid; c < C; c += kCT) {
            float sxm = 0.0f, sx = 0.0f, best = -INFINITY;
            int bidx = -1;
            for (int rr = 0; rr < CS; ++rr) {  // ranks own increasing pixel ranges: strict > keeps the first maximum
                const float* pp = part + rr * 4 * C;
                sxm += pp[c];
                sx += pp[C + c];
                const float ob = pp[2 * C + c];
                const int oi = __float_as_int(pp[3 * C + c]);
                if (oi >= 0 && (bidx < 0 || ob > best)) { best = ob; bidx = oi; }
            }
            const float G = sx * invS;
            const float A = has_mask ? sxm / den : G;
            const float avg = has_mask ? (A * use + G * (1.0f - use)) : G;
            const bool dead = bidx < 0;  // no pixel with m > 0.5 (masked_cbam.py:118-121)
            const float mx = dead ? G : best;
            s_avg[c] = avg;
            s_mx[c] = mx;
            if (r == 0) {
                const int i = b * C + c;
                ctx.avg[i] = avg; ctx.mx[i] = mx; ctx.apool[i] = A; ctx.amax[i] = dead ? -1 : bidx;
            }
        }
    }
    const float alpha = softplusf_acc(beta);
    const float k0 = sh.pyramid_multiply() ? 0.0f : 1.0f - alpha, k1 = alpha;
    if (b == 0 && r == 0 && tid == 0) { ctx.consts[0] = k0; ctx.consts[1] = k1; ctx.consts[2] = alpha; ctx.consts[3] = sigmoidf_acc(beta); }
    __syncthreads();
    for (int j = w; j < Hd; j += kCW) {
        const float* wr = prm.w1 + (size_t)j * C;
        float pa = 0.0f, pm = 0.0f;
#pragma unroll 4
        for (int c = lane; c < C; c += 32) {
            const float wv = __ldg(wr + c);
            pa = fmaf(wv, s_avg[c], pa);
            pm = fmaf(wv, s_mx[c], pm);
        }
        pa = warp_sum(pa);
        pm = warp_sum(pm);
        if (lane == 0) {
            const float bb = __ldg(prm.b1 + j);
            const float ha = fmaxf(pa + bb, 0.0f), hm = fmaxf(pm + bb, 0.0f);
            s_ha[j] = ha;
            s_hm[j] = hm;
            if (r == 0) { ctx.ha[b * Hd + j] = ha; ctx.hm[b * Hd + j] = hm; }
        }
    }
    __syncthreads();
    {
        const bool add = sh.samcam_add();
        for (int c = tid; c < C; c += kCT) {
            const float* wr = prm.w2 + (size_t)c * Hd;
            float za = 0.0f, zm = 0.0f;
#pragma unroll 4
            for (int j = 0; j < Hd; ++j) {
                const float wv = __ldg(wr + j);
                za = fmaf(wv, s_ha[j], za);
                zm = fmaf(wv, s_hm[j], zm);
            }
            const float bb = __ldg(prm.b2 + c);
            const float s = sigmoidf_acc((za + bb) + (zm + bb));  // b2 enters twice (masked_cbam.py:128)
            if (r == 0) ctx.s[b * C + c] = s;
            s_q[c] = add ? 1.0f : s;
            // out = x * (sA * a + sB):  multiply: k0 + k1*s*a ;  add: k0 + k1*(s + a)
            sA[c] = add ? k1 : k1 * s;
            sB[c] = add ? fmaf(k1, s, k0) : k0;
        }
    }
    __syncthreads();

    // ---- phase 3 (T2): per pixel max / arg max / sum over channels of x*q
    stamp(4);
    const int g2 = tid / nUmax, ul2 = tid - g2 * nUmax;
    const int G = gm.G;
    const bool act2 = g2 < G && ul2 < nU;
    constexpr int KB = (VEC == 4) ? 8 : 4;
    {
        float vmax[VEC], vsum[VEC];
        int vidx[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) { vmax[i] = -INFINITY; vsum[i] = 0.0f; vidx[i] = 0x7fffffff; }
        if (act2) {
            const T* xp = xb + (size_t)ul2 * VEC;
            for (int c0 = g2; c0 < C; c0 += G * KB) {
                uint4 raw[KB];
#pragma unroll
                for (int kc = 0; kc < KB; ++kc) {
                    const int c = c0 + kc * G;
                    if (c < C) raw[kc] = ldg128<kPolKeep>(xp + (size_t)c * S, pol);
                }
#pragma unroll
                for (int kc = 0; kc < KB; ++kc) {
                    const int c = c0 + kc * G;
                    if (c < C) {
                        float v[VEC];
                        unpack<T, VEC>(raw[kc], v);
                        const float q = s_q[c];
#pragma unroll
                        for (int i = 0; i < VEC; ++i) {
                            const float y = v[i] * q;
                            vsum[i] += y;
                            if (y > vmax[i]) { vmax[i] = y; vidx[i] = c; }
                        }
                    }
                }
            }
        }
        if (g2 < G && ul2 < nUmax) {
            float* d = mg + (size_t)g2 * nPmax + ul2 * VEC;
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                d[i] = vmax[i];
                d[(size_t)G * nPmax + i] = vsum[i];
                d[(size_t)2 * G * nPmax + i] = __int_as_float(vidx[i]);
            }
        }
    }
    __syncthreads();
    {
        const float invC = 1.0f / (float)C;
        for (int p = tid; p < nP; p += kCT) {
            float bm = mg[p], bsum = mg[(size_t)G * nPmax + p];
            int bi = __float_as_int(mg[(size_t)2 * G * nPmax + p]);
            for (int j = 1; j < G; ++j) {
                const float om = mg[(size_t)j * nPmax + p];
                const int oi = __float_as_int(mg[(size_t)(2 * G + j) * nPmax + p]);
                bsum += mg[(size_t)(G + j) * nPmax + p];
                if (om > bm || (om == bm && oi < bi)) { bm = om; bi = oi; }  // torch.max: first maximal channel
            }
            const float pavg = bsum * invC;
            const int ry = p / W, cx = p - ry * W;
            tile[(3 + ry) * TWp + 4 + cx] = bm;
            tile[planeT + (3 + ry) * TWp + 4 + cx] = pavg;
            ctx.pmax[bS + p0 + p] = bm;
            ctx.pavg[bS + p0 + p] = pavg;
            ctx.idx[bS + p0 + p] = bi;
        }
    }
    stamp(5);
    cluster.sync();  // #2: every rank's own rows of the three planes are in its tile

    // ---- phase 4: halo rows through DSMEM, then a = sigmoid(conv7x7([pmax, pavg, m])) for the own rows
    stamp(6);
    if (nP > 0) cl_fetch_halo(cluster, tile, has_mask ? 3 : 2, gm, y0, rows, H, W);
    cluster.barrier_arrive();  // #3 (waited at the end): this CTA reads no remote shared memory any more
    __syncthreads();
    {
        float* cpart = mg;  // [3][nPmax]
        const int grp = tid / kConvGroup, gl = tid - grp * kConvGroup;
        const int spr = W / 4, nStrips = rows * spr;
        if (grp < 3) {
            const float* wv = wk + grp * kMaxK * kMaxK;
            for (int s = gl; s < nStrips; s += kConvGroup) {
                const int ry = s / spr, x0 = (s - ry * spr) * 4;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                strip_conv7(tile + grp * planeT + ry * TWp + x0, TWp, wv, acc);
                *reinterpret_cast<float4*>(cpart + (size_t)grp * nPmax + s * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            }
        }
        __syncthreads();
        for (int p = tid; p < nP; p += kCT) {
            const float v = (cpart[p] + cpart[nPmax + p]) + cpart[2 * nPmax + p];
            const float a = sigmoidf_acc(v);
            aloc[p] = a;
            ctx.a[bS + p0 + p] = a;
        }
    }
    __syncthreads();

    // ---- phase 5 (T2): out = x * (sA*a + sB)
    stamp(7);
    if (act2) {
        float av[VEC];
        lds_f<VEC>(aloc + ul2 * VEC, av);
        const T* xp = xb + (size_t)ul2 * VEC;
        T* op = out + (size_t)b * C * S + p0 + (size_t)ul2 * VEC;
        for (int c0 = g2; c0 < C; c0 += G * KB) {
            uint4 raw[KB];
#pragma unroll
            for (int kc = 0; kc < KB; ++kc) {
                const int c = c0 + kc * G;
                if (c < C) raw[kc] = ldg128<kPolLast>(xp + (size_t)c * S, pol);
            }
#pragma unroll
            for (int kc = 0; kc < KB; ++kc) {
                const int c = c0 + kc * G;
                if (c < C) {
                    float v[VEC];
                    unpack<T, VEC>(raw[kc], v);
                    const float ga = sA[c], gb = sB[c];
#pragma unroll
                    for (int i = 0; i < VEC; ++i) v[i] *= fmaf(ga, av[i], gb);
                    __stcs(reinterpret_cast<uint4*>(op + (size_t)c * S), pack<T, VEC>(v));
                }
            }
        }
    }
    stamp(8);
    cluster.barrier_wait();  // #3
}

// ================================================================== backward
template <typename T>
__global__ void __launch_bounds__(kCT, MGA_CL_MINB) cl_bwd_kernel(const T* __restrict__ x, const T* __restrict__ g, const void* __restrict__ mask,
                                                                  int mdt, T* __restrict__ dx, void* __restrict__ dmask, Shape sh,
                                                                  mga_cbam_params prm, Ctx ctx, BwdScratch bs, ClGeom gm) {
    constexpr int VEC = 16 / sizeof(T);
    constexpr int K = ClK<VEC>::K;
    constexpr int kStride = 3 * kMaxK * kMaxK + 1;
    extern __shared__ __align__(128) float clsm[];
    float* const csm = clsm;
    cgx::cluster_group cluster = cgx::this_cluster();
    const int CS = gm.CS;
    const int r = (int)cluster.block_rank();
    const int b = blockIdx.x / CS;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int C = sh.C, S = sh.S, Hd = sh.hidden, W = sh.W, H = sh.H;
    const int y0 = r * gm.rowsPer;
    const int rows = max(0, min(gm.rowsPer, H - y0));
    const int nP = rows * W, nU = nP / VEC, p0 = y0 * W;
    const int nPmax = gm.nPmax, nUmax = gm.nUmax, TWp = gm.TWp, planeT = gm.planeT;
    const bool has_mask = sh.has_mask();
    const bool multiply = !sh.samcam_add();
    const ClBwdOff o = cl_bwd_off(C, Hd, gm);
    float* wsm = csm + o.wsm;
    double* redd = reinterpret_cast<double*>(csm + o.red);
    float* s_s = csm + o.s;
    float* s_q = csm + o.q;
    float4* chA = reinterpret_cast<float4*>(csm + o.chA);
    float2* chB = reinterpret_cast<float2*>(csm + o.chB);
    int* amx = reinterpret_cast<int*>(csm + o.amx);
    float* s_dz = csm + o.dz;
    float* e_loc = csm + o.eloc;
    float* q_loc = csm + o.qloc;
    float* bin_loc = csm + o.binloc;
    float* s_dha = csm + o.dha;
    float* s_dhm = csm + o.dhm;
    float* epart = csm + o.epart;
    float* qpart = csm + o.qpart;
    float* binpart = csm + o.binpart;
    double* gxpart = reinterpret_cast<double*>(csm + o.gxpart);
    float* aloc = csm + o.aloc;
    float* ae = csm + o.ae;
    float* mloc = csm + o.mloc;
    int* idxl = reinterpret_cast<int*>(csm + o.idx);
    float* pmx = csm + o.pmx;
    float* d0 = csm + o.d0;
    float* d1s = csm + o.d1s;
    float* d2 = csm + o.d2;
    float* dpre = csm + o.dpre;
    float* cat = csm + o.cat;
    float* tp = csm + o.tp;
    float* dwp = csm + o.dwp;
    float* binw = csm + o.binw;
    const ClPol pol = cl_policies();
    const size_t bS = (size_t)b * S;
    const float k0 = ctx.consts[0], k1 = ctx.consts[1];

    // ---- phase 0: per-sample vectors and planes of the saved context
    stamp(0);
    load_weights7(prm.wsam, sh.k, true, wsm);
    for (int i = tid; i < planeT; i += kCT) dpre[i] = 0.0f;
    for (int i = tid; i < kCW * C; i += kCT) binw[i] = 0.0f;
    for (int c = tid; c < C; c += kCT) {
        const float s = ctx.s[b * C + c];
        s_s[c] = s;
        s_q[c] = multiply ? s : 1.0f;
    }
    for (int p = tid; p < nP; p += kCT) {
        const float a = ctx.a[bS + p0 + p];
        aloc[p] = a;
        ae[p] = multiply ? a : 1.0f;
        mloc[p] = has_mask ? ctx.m[bS + p0 + p] : 0.0f;
        idxl[p] = ctx.idx[bS + p0 + p];
        pmx[p] = ctx.pmax[bS + p0 + p];
    }
    {
        const float* const planes[3] = {ctx.pmax + bS, ctx.pavg + bS, has_mask ? ctx.m + bS : nullptr};
        stage_three(cat, planeT, planes, y0 - kMaxK / 2, gm.tileRows, H, W, TWp);
    }
    __syncthreads();

    // ---- phase 1 (T1) over (x,g): T_p = sum_c g x q_c (per pixel), E_c = sum_p g x (a | 1) (per channel), sum g x
    stamp(1);
    const int lsh = gm.lsh, LPT = 1 << lsh, CPW = 32 >> lsh;
    const int sub = lane >> lsh, ul = lane & (LPT - 1);
    const size_t sbase = (size_t)b * C * S + p0;
    const T* xb = x + sbase;
    const T* gb = g + sbase;
    float gxs = 0.0f;
    {
        float tacc[K][VEC];
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int i = 0; i < VEC; ++i) tacc[k][i] = 0.0f;
        for (int cb = w * CPW; cb < C; cb += kCW * CPW) {
            const int c = cb + sub;
            const bool cok = c < C;
            uint4 xr[K], gr[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int u = ul + (k << lsh);
                if (cok && u < nU) {
                    xr[k] = ldg128<kPolKeep>(xb + (size_t)c * S + (size_t)u * VEC, pol);
                    gr[k] = ldg128<kPolKeep>(gb + (size_t)c * S + (size_t)u * VEC, pol);
                }
            }
            const float q = cok ? s_q[c] : 0.0f;
            float e = 0.0f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int u = ul + (k << lsh);
                if (cok && u < nU) {
                    float xv[VEC], gv[VEC], av[VEC];
                    unpack<T, VEC>(xr[k], xv);
                    unpack<T, VEC>(gr[k], gv);
                    lds_f<VEC>(ae + u * VEC, av);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const float gx = gv[i] * xv[i];
                        tacc[k][i] = fmaf(gx, q, tacc[k][i]);
                        e = fmaf(gx, av[i], e);
                        gxs += gx;
                    }
                }
            }
            for (int of = LPT >> 1; of > 0; of >>= 1) e += __shfl_xor_sync(0xffffffffu, e, of);
            if (ul == 0 && cok) e_loc[c] = e;
        }
        // merge the channel sub-groups of the warp, then the 8 warps through shared memory
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int u = ul + (k << lsh);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                float t = tacc[k][i];
                for (int of = LPT; of < 32; of <<= 1) t += __shfl_xor_sync(0xffffffffu, t, of);
                if (sub == 0 && u < nU) tp[(size_t)w * nPmax + u * VEC + i] = t;
            }
        }
    }
    __syncthreads();
    double at_acc = 0.0;
    for (int p = tid; p < nP; p += kCT) {
        float t = 0.0f;
#pragma unroll
        for (int j = 0; j < kCW; ++j) t += tp[(size_t)j * nPmax + p];
        const float a = aloc[p];
        const int ry = p / W, cx = p - ry * W;
        dpre[(3 + ry) * TWp + 4 + cx] = k1 * t * a * (1.0f - a);
        at_acc += (double)a * (double)t;
    }
    stamp(2);
    cluster.sync();  // #1: every rank's own rows of dpre are in its tile

    // ---- phase 2: dpre halo through DSMEM; dcat = conv7x7^T(dpre); dWsam partials
    stamp(3);
    if (nP > 0) cl_fetch_halo(cluster, dpre, 1, gm, y0, rows, H, W);
    __syncthreads();
    {
        const int grp = tid / kConvGroup, gl = tid - grp * kConvGroup;
        const int spr = W / 4, nStrips = rows * spr;
        const float invC = 1.0f / (float)C;
        if (grp < 3) {
            const float* wv = wsm + grp * kMaxK * kMaxK;
            float* dst = grp == 0 ? d0 : (grp == 1 ? d1s : d2);
            const float sc = grp == 1 ? invC : 1.0f;
            for (int s = gl; s < nStrips; s += kConvGroup) {
                const int ry = s / spr, x0 = (s - ry * spr) * 4;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                strip_conv7(dpre + ry * TWp + x0, TWp, wv, acc);
                *reinterpret_cast<float4*>(dst + s * 4) = make_float4(acc[0] * sc, acc[1] * sc, acc[2] * sc, acc[3] * sc);
            }
        }
        // conv2d_weight: a team of 12 threads owns one (plane, kernel row) pair, the 7 column taps live in registers
        constexpr int kTeam = 12;
        const int team = tid / kTeam, tl = tid - team * kTeam;
        if (team < 3 * kMaxK) {
            const int pl = team / kMaxK, i7 = team - pl * kMaxK;
            float acc[kMaxK] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int s = tl; s < nStrips; s += kTeam) {
                const int ry = s / spr, x0 = (s - ry * spr) * 4;
                const float4 d = *reinterpret_cast<const float4*>(dpre + (ry + kMaxK / 2) * TWp + x0 + 4);
                const float* cr = cat + pl * planeT + (ry + i7) * TWp + x0;
                const float4 a4 = *reinterpret_cast<const float4*>(cr);
                const float4 b4 = *reinterpret_cast<const float4*>(cr + 4);
                const float4 c4 = *reinterpret_cast<const float4*>(cr + 8);
                const float win[12] = {a4.x, a4.y, a4.z, a4.w, b4.x, b4.y, b4.z, b4.w, c4.x, c4.y, c4.z, c4.w};
                const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                for (int j = 0; j < kMaxK; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[j] = fmaf(win[1 + i + j], dv[i], acc[j]);
            }
#pragma unroll
            for (int j = 0; j < kMaxK; ++j) dwp[(team * kTeam + tl) * kMaxK + j] = acc[j];
        }
        __syncthreads();
        const int cta = b * CS + r;
        const int k = sh.k, koff = (kMaxK - k) / 2;
        if (tid < 3 * kMaxK * kMaxK) {
            const int tm = tid / kMaxK, j7 = tid - tm * kMaxK;  // tm = pl * 7 + i7
            const int pl = tm / kMaxK, ii = tm - pl * kMaxK - koff, jj = j7 - koff;
            if (ii >= 0 && ii < k && jj >= 0 && jj < k) {
                float t = 0.0f;
#pragma unroll
                for (int q = 0; q < kTeam; ++q) t += dwp[(tm * kTeam + q) * kMaxK + j7];
                bs.convpart[(size_t)cta * kStride + (pl * k + ii) * k + jj] = t;
            }
        }
        const double at = cl_block_sum_d(at_acc, redd);
        if (tid == 0) bs.atpart[cta] = at;
    }
    __syncthreads();

    // ---- phase 3 (T1, multiply mode) over x: Q_c = sum_p x * dcat1/C ; the arg-max term of ds goes through pmax:
    //      sum_{p: idx_p = c} dcat0_p x_cp s_c = sum_{p: idx_p = c} dcat0_p pmax_p   (binned per channel, deterministic)
    stamp(4);
    if (multiply) {
        for (int cb = w * CPW; cb < C; cb += kCW * CPW) {
            const int c = cb + sub;
            const bool cok = c < C;
            uint4 xr[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int u = ul + (k << lsh);
                if (cok && u < nU) xr[k] = ldg128<kPolKeep>(xb + (size_t)c * S + (size_t)u * VEC, pol);
            }
            float qv = 0.0f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int u = ul + (k << lsh);
                if (cok && u < nU) {
                    float xv[VEC], wv[VEC];
                    unpack<T, VEC>(xr[k], xv);
                    lds_f<VEC>(d1s + u * VEC, wv);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) qv = fmaf(xv[i], wv[i], qv);
                }
            }
            for (int of = LPT >> 1; of > 0; of >>= 1) qv += __shfl_xor_sync(0xffffffffu, qv, of);
            if (ul == 0 && cok) q_loc[c] = qv;
        }
        // warp-private bins: lanes with the same arg-max channel are summed in lane order by their lowest lane
        for (int base = w * 32; base < nP; base += kCT) {
            const int p = base + lane;
            const bool ok = p < nP;
            const int key = ok ? idxl[p] : -1 - lane;
            const float v = ok ? d0[p] * pmx[p] : 0.0f;
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            float acc = 0.0f;
#pragma unroll 8
            for (int l = 0; l < 32; ++l) {
                const float t = __shfl_sync(0xffffffffu, v, l);
                if ((peers >> l) & 1u) acc += t;
            }
            if (ok && lane == __ffs(peers) - 1 && key >= 0 && key < C) binw[w * C + key] += acc;
            __syncwarp();
        }
    }
    __syncthreads();
    for (int c = tid; c < C; c += kCT) {
        float t = 0.0f;
        if (multiply) {
#pragma unroll
            for (int j = 0; j < kCW; ++j) t += binw[j * C + c];
        } else {
            q_loc[c] = 0.0f;
        }
        bin_loc[c] = t;
    }
    const double gx_cta = cl_block_sum_d((double)gxs, redd);  // (barriers inside)
    for (int i = tid; i < C * CS; i += kCT) {
        const int rr = i / C, c = i - rr * C;
        cluster.map_shared_rank(epart, rr)[r * C + c] = e_loc[c];
        cluster.map_shared_rank(qpart, rr)[r * C + c] = q_loc[c];
        cluster.map_shared_rank(binpart, rr)[r * C + c] = bin_loc[c];
    }
    if (tid < CS) cluster.map_shared_rank(gxpart, tid)[r] = gx_cta;
    stamp(5);
    cluster.sync();  // #2: per-channel partials of all ranks are local; no remote access after this point

    // ---- phase 4: ds -> dz -> MLP backward -> per-channel coefficients (recomputed by every CTA)
    stamp(6);
    float kb = 0.0f;
    {
        double gx_tot = 0.0;
        for (int rr = 0; rr < CS; ++rr) gx_tot += gxpart[rr];
        double se_acc = 0.0;
        for (int c = tid; c < C; c += kCT) {
            float es = 0.0f, qs = 0.0f, bn = 0.0f;
            for (int rr = 0; rr < CS; ++rr) { es += epart[rr * C + c]; qs += qpart[rr * C + c]; bn += binpart[rr * C + c]; }
            const float s = s_s[c];
            const float dz = fmaf(fmaf(k1, es, qs) * s, 1.0f - s, (1.0f - s) * bn);
            s_dz[c] = dz;
            if (r == 0) bs.dz[b * C + c] = dz;
            if (!multiply) se_acc += (double)s * (double)es;  // add mode: sum_c s_c sum_p g x
        }
        const double ssum = cl_block_sum_d(se_acc, redd);
        if (r == 0 && tid == 0) bs.alphapart[b] = ssum - (sh.pyramid_multiply() ? 0.0 : gx_tot);
        __syncthreads();
        for (int j = w; j < Hd; j += kCW) {
            float acc = 0.0f;
#pragma unroll 4
            for (int c = lane; c < C; c += 32) acc = fmaf(s_dz[c], __ldg(prm.w2 + (size_t)c * Hd + j), acc);
            acc = warp_sum(acc);
            if (lane == 0) {
                const float da = ctx.ha[b * Hd + j] > 0.0f ? acc : 0.0f;
                const float dm = ctx.hm[b * Hd + j] > 0.0f ? acc : 0.0f;
                s_dha[j] = da;
                s_dhm[j] = dm;
                if (r == 0) { bs.dha[b * Hd + j] = da; bs.dhm[b * Hd + j] = dm; }
            }
        }
        __syncthreads();
        const float use = has_mask ? ctx.use[b] : 0.0f;
        const float den = has_mask ? ctx.den[b] : 1.0f;
        const float pass = (has_mask && ctx.msum[b] >= sh.eps) ? 1.0f : 0.0f;  // clamp_min backward
        const float invS = 1.0f / (float)S;
        float kacc = 0.0f;
        for (int c = tid; c < C; c += kCT) {
            float davg = 0.0f, dmx = 0.0f;
#pragma unroll 4
            for (int j = 0; j < Hd; ++j) {
                const float wv = __ldg(prm.w1 + (size_t)j * C + c);
                davg = fmaf(s_dha[j], wv, davg);
                dmx = fmaf(s_dhm[j], wv, dmx);
            }
            const int i = b * C + c;
            const int am = ctx.amax[i];
            const bool dead = has_mask && am < 0;
            const float cA = has_mask ? use * davg / den : 0.0f;
            const float cG = ((1.0f - use) * davg + (dead ? dmx : 0.0f)) * invS;
            const float cM = dead ? 0.0f : dmx;
            const float s = s_s[c];
            // dL/dout * gate: multiply: k0 + k1*s*a ; add: k0 + k1*(s + a)
            chA[c] = make_float4(multiply ? k1 * s : k1, multiply ? k0 : fmaf(k1, s, k0), multiply ? s : 1.0f, cA);
            chB[c] = make_float2(cG, cM);
            amx[c] = am;
            kacc = fmaf(cA, ctx.apool[i] * pass, kacc);
        }
        kb = block_sum(kacc, reinterpret_cast<float*>(redd));
    }
    __syncthreads();

    // ---- phase 5 (T2) over (x,g): dx (streaming store) and R_p = sum_c cA_c x -> dmask
    stamp(7);
    const int g2 = tid / nUmax, ul2 = tid - g2 * nUmax;
    const int G = gm.G;
    const bool act2 = g2 < G && ul2 < nU;
    const bool want_dmask = has_mask && dmask != nullptr;
    {
        float racc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) racc[i] = 0.0f;
        if (act2) {
            float av[VEC], wv[VEC], wdv[VEC], mv[VEC];
            int ix[VEC];
            lds_f<VEC>(aloc + ul2 * VEC, av);
            lds_f<VEC>(d1s + ul2 * VEC, wv);
            lds_f<VEC>(d0 + ul2 * VEC, wdv);
            lds_f<VEC>(mloc + ul2 * VEC, mv);
            lds_i<VEC>(idxl + ul2 * VEC, ix);
#pragma unroll
            for (int i = 0; i < VEC; ++i) wdv[i] += wv[i];
            const int pbase = p0 + ul2 * VEC;
            constexpr int KB = (VEC == 4) ? 4 : 2;
            const T* xp = xb + (size_t)ul2 * VEC;
            const T* gp = gb + (size_t)ul2 * VEC;
            T* op = dx + sbase + (size_t)ul2 * VEC;
            for (int c0 = g2; c0 < C; c0 += G * KB) {
                uint4 xr[KB], gr[KB];
#pragma unroll
                for (int kc = 0; kc < KB; ++kc) {
                    const int c = c0 + kc * G;
                    if (c < C) {
                        xr[kc] = ldg128<kPolLast>(xp + (size_t)c * S, pol);
                        gr[kc] = ldg128<kPolLast>(gp + (size_t)c * S, pol);
                    }
                }
#pragma unroll
                for (int kc = 0; kc < KB; ++kc) {
                    const int c = c0 + kc * G;
                    if (c < C) {
                        float xv[VEC], gv[VEC], ov[VEC];
                        unpack<T, VEC>(xr[kc], xv);
                        unpack<T, VEC>(gr[kc], gv);
                        const float4 ca = chA[c];
                        const float2 cb2 = chB[c];
                        const int am = amx[c];
#pragma unroll
                        for (int i = 0; i < VEC; ++i) {
                            const float gate = fmaf(ca.x, av[i], ca.y);
                            const float sel = (ix[i] == c) ? wdv[i] : wv[i];
                            float v = fmaf(ca.w, mv[i], cb2.x);
                            v = fmaf(ca.z, sel, v);
                            ov[i] = fmaf(gv[i], gate, v);
                            racc[i] = fmaf(ca.w, xv[i], racc[i]);
                        }
                        const int dlt = am - pbase;
                        if ((unsigned)dlt < (unsigned)VEC) {
#pragma unroll
                            for (int i = 0; i < VEC; ++i)
                                if (i == dlt) ov[i] += cb2.y;
                        }
                        __stcs(reinterpret_cast<uint4*>(op + (size_t)c * S), pack<T, VEC>(ov));
                    }
                }
            }
        }
        if (want_dmask && g2 < G && ul2 < nUmax) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) tp[(size_t)g2 * nPmax + ul2 * VEC + i] = racc[i];
        }
    }
    if (want_dmask) {
        __syncthreads();
        for (int p = tid; p < nP; p += kCT) {
            float rsum = 0.0f;
            for (int j = 0; j < G; ++j) rsum += tp[(size_t)j * nPmax + p];
            float dm = (rsum - kb) + d2[p];
            if (sh.sigmoid_mask()) {
                const float m = mloc[p];
                dm *= m * (1.0f - m);
            }
            if (sh.gate_clamp()) {
                const float raw = load_mask_any(mask, mdt, bS + p0 + p);
                if (!(raw >= 0.0f && raw <= 1.0f)) dm = 0.0f;
            }
            store_mask_any(dmask, mdt, bS + p0 + p, dm);
        }
    }
    stamp(8);
}

}  // namespace mga
