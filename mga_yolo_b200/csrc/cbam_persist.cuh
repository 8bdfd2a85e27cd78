// cbam_persist.cuh -- persistent, shared-memory-resident forward / backward kernels (the default path for shapes that fit).
//
// What bounds the cluster-per-sample kernels (cbam_cluster.cuh) is not HBM bandwidth but the latency of their second and third
// pass over x (L2 round trips) in a CTA that lives for one sample only (profiles/r1_cluster_kernels.md).  Here:
//
//   * the sample's slice of x (rowsPer image rows of ALL channels, <= ~84 KB) is brought into SHARED MEMORY once, by tensor-map
//     TMA (cp.async.bulk.tensor, four channel chunks, one mbarrier each: no registers, no scoreboard), and every later pass reads
//     it from there -- HBM sees x exactly once and L2 is not needed to hold anything;
//   * CTAs are PERSISTENT (2 per SM, co-resident by a cooperative launch): TS consecutive CTAs form the team of one sample and
//     walk a static list of (level, sample) items; while a CTA streams the result of item i out (pass 3 frees the slice chunk by
//     chunk) the TMA loads of item i+1 already fill the freed chunks, so reads and writes overlap and the next pooling pass finds
//     its data in place;
//   * all pyramid levels of a step go into ONE launch (teams of ST CTAs split into ST/TS sub-teams for the smaller levels), so the
//     small levels fill what would be the tail of the large one;
//   * teams synchronise through global memory (release/acquire counters, two barriers per item and direction) and exchange their
//     per-channel partial sums and the 3 halo rows of the small planes through L2 -- no cluster hardware, hence no limit on which
//     SMs can work together (8-CTA clusters could use only 264 of the 296 CTA slots).
//
// Reference lines: mga_yolo/nn/modules/masked_cbam.py:87-171 (forward); SURVEY.md section 8a "Backward" (closed form).
#pragma once
#include <cuda.h>

#include "cbam_cluster.cuh"

namespace mga {

constexpr int kPfNTF = 256;    // threads per CTA, forward  (2 CTAs per SM; few threads = little redundant per-thread setup per item)
constexpr int kPfNTB = 256;    // threads per CTA, backward (128 registers, 2 CTAs per SM)
constexpr int kPfNCH = 4;      // channel chunks of a slice: one TMA box (pair) + one mbarrier each
constexpr int kPfMaxLv = 3;    // pyramid levels per launch
constexpr int kPfMaxCtas = 2 * kSMs;
constexpr int kPfSliceBytes = 84 * 1024;  // largest x slice per CTA (two CTAs of <= 113 KB per SM)

struct PfLevel {
    const void* x; const void* mask; void* out;  // forward
    const void* g; void* dx; void* dmask;        // backward
    float* xchg;      // per-CTA exchange slots of this level: [kPfMaxCtas][xstride] floats
    unsigned* ctr;    // team barrier counters of this level: [kPfMaxCtas]
    Shape sh;
    mga_cbam_params prm;
    Ctx ctx;
    BwdScratch bs;
    int mdt;
    int TS;           // CTAs per sample
    int spi;          // samples per item = ST / TS
    int rowsPer;      // image rows per CTA
    int nP, nU;       // pixels / 16-byte units of a full slice row range
    int nsplit;       // TMA boxes per chunk along the pixel axis (box <= 256 elements): 1 or 2
    int nPbox;        // pixels per box
    int CCH;          // channels per chunk
    int chunkBytes;   // bytes per chunk (all boxes)
    int LPC;          // pooling: lanes per channel row
    int CPL;          // pooling: channels per lane (1 or 2: two channels share the loads of the mask planes)
    int UL;           // channel max/mean: units per warp task (8 / 16 / 32; the other 32/UL lane groups split the channels)
    int WS;           // channel max/mean: warps that share one group of UL units
    int G;            // rescale / dx: channel groups = NT / nU
    int TWp, planeT;  // conv tile: padded width, floats per plane
    int first_item, n_items;
    int xstride;      // floats per exchange slot
};

struct PfFwdOff { int bar, red, wk, avg, mx, ha, hm, q2, AB4, pst, mloc, mb, aloc, mraw, R, total; };

struct PfParams {
    PfLevel lv[kPfMaxLv];
    int n_levels, ST, nST, total_items;
    int xbytes;  // bytes of the x-slice region at the start of shared memory (max over levels, multiple of 128)
    int stagger_ns;  // super-team t starts t * stagger_ns / nST late: spreads the HBM phases of the teams over the item period
    PfFwdOff fo;
};
struct PfMaps { CUtensorMap x[kPfMaxLv]; };

struct PfItem { int valid, L, b, r; };

__device__ __forceinline__ PfItem pf_decode(const PfParams& P, int it, int rk) {
    PfItem x{0, 0, 0, 0};
    if (it >= P.total_items) return x;
    int L = 0;
    while (L + 1 < P.n_levels && it >= P.lv[L + 1].first_item) ++L;
    const int TS = P.lv[L].TS;
    const int sub = rk / TS;
    x.L = L;
    x.r = rk - sub * TS;
    x.b = (it - P.lv[L].first_item) * P.lv[L].spi + sub;
    x.valid = (x.b < P.lv[L].sh.B && sub < P.lv[L].spi) ? 1 : 0;
    return x;
}

// 2-D tiled TMA load, L2 evict-first (the slice is read from HBM exactly once)
__device__ __forceinline__ void tma_load_2d_once(void* dst, const void* tmap, int c0, int c1, uint64_t* bar, unsigned long long pol) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// chunk k of the slice (level lv, sample b, rank r): x viewed as a 2-D tensor (pixels of a plane, B*C planes); the slice is laid
// out [nsplit][C][nPbox] so that a (channel, unit) address does not depend on the chunking
template <int ESIZE>
__device__ __forceinline__ void pf_issue_chunk(const PfLevel& lv, const CUtensorMap* map, unsigned char* xs, uint64_t* full, int k, int b, int r,
                                               unsigned long long pol) {
    const int p0 = r * lv.rowsPer * lv.sh.W;
    const int rowB = lv.nPbox * ESIZE;
    mbar_expect_tx(&full[k], (uint32_t)lv.chunkBytes);
    for (int h = 0; h < lv.nsplit; ++h)
        tma_load_2d_once(xs + (size_t)h * lv.sh.C * rowB + (size_t)k * lv.CCH * rowB, map, p0 + h * lv.nPbox, b * lv.sh.C + k * lv.CCH, &full[k], pol);
}

// barrier among the TS CTAs of a team, split in two so that independent work can sit between the halves: every thread's global
// writes before pf_team_arrive are visible to every thread of the team after pf_team_wait
__device__ __forceinline__ void pf_team_arrive(unsigned* ctr) {
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}
__device__ __forceinline__ void pf_team_wait(unsigned* ctr, unsigned target) {
    if (threadIdx.x == 0) {
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

// ---------------------------------------------------------------- packed fp32x2 helpers (sm_100: FADD2 / FMUL2 / FFMA2 halve the issue slots)
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// one 16-byte unit of feature-map elements from shared memory as VEC/2 float pairs
template <typename T>
__device__ __forceinline__ void pf_unit(const void* p, float2 (&v)[8 / sizeof(T)]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    if constexpr (sizeof(T) == 4) {
        v[0] = make_float2(__uint_as_float(t.x), __uint_as_float(t.y));
        v[1] = make_float2(__uint_as_float(t.z), __uint_as_float(t.w));
    } else {
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if constexpr (std::is_same<T, __nv_bfloat16>::value) v[i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u));
            else v[i] = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        }
    }
}
template <typename T>
__device__ __forceinline__ uint4 pf_pack(const float2 (&v)[8 / sizeof(T)]) {
    if constexpr (sizeof(T) == 4) {
        return make_uint4(__float_as_uint(v[0].x), __float_as_uint(v[0].y), __float_as_uint(v[1].x), __float_as_uint(v[1].y));
    } else {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                __nv_bfloat162 h = __float22bfloat162_rn(v[i]);
                w[i] = *reinterpret_cast<uint32_t*>(&h);
            } else {
                __half2 h = __float22half2_rn(v[i]);
                w[i] = *reinterpret_cast<uint32_t*>(&h);
            }
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
}
// N2 float pairs (fp32 planes) from shared memory, 16-byte aligned
template <int N2>
__device__ __forceinline__ void pf_pairs(const float* p, float2 (&v)[N2]) {
#pragma unroll
    for (int i = 0; i < N2 / 2; ++i) {
        const float4 t = reinterpret_cast<const float4*>(p)[i];
        v[2 * i] = make_float2(t.x, t.y);
        v[2 * i + 1] = make_float2(t.z, t.w);
    }
}
__device__ __forceinline__ float pf_mask_val(const void* raw, int mdt, int i) {
    if (mdt == MGA_F32) return static_cast<const float*>(raw)[i];
    if (mdt == MGA_BF16) return __bfloat162float(static_cast<const __nv_bfloat16*>(raw)[i]);
    return __half2float(static_cast<const __half*>(raw)[i]);
}
// rows [max(0, y0-3), min(H, y0+rows+3)) of sample b's mask: one contiguous bulk copy into the staging buffer
__device__ __forceinline__ void pf_issue_mask(const PfLevel& lv, void* mraw, uint64_t* bar, int b, int r) {
    const int H = lv.sh.H, W = lv.sh.W;
    const int y0 = r * lv.rowsPer, rows = min(lv.rowsPer, H - y0);
    const int ys = max(0, y0 - 3), ye = min(H, y0 + rows + 3);
    const int msz = lv.mdt == MGA_F32 ? 4 : 2;
    const uint32_t bytes = (uint32_t)((ye - ys) * W * msz);
    mbar_expect_tx(bar, bytes);
    bulk_g2s(mraw, static_cast<const char*>(lv.mask) + ((size_t)b * lv.sh.S + (size_t)ys * W) * msz, bytes, bar);
}

// ================================================================== forward
// pooling of CPL channels (rows `row`, `row + rowB`) over the lane's units: sum x*m, sum x, and the running maximum of the valid
// pixels kept per UNIT (one compare per 16 bytes); the pixel inside the winning unit is resolved by the caller
template <typename T, int CPL>
__device__ __forceinline__ void pf_pool_lane(const unsigned char* row, int rowB, const float* mloc, const float* mb, int l, int LPC, int nU, int nUbox,
                                             int hoff, float2 (&sx2)[CPL], float2 (&sxm2)[CPL], float (&best)[CPL], int (&bu)[CPL]) {
    constexpr int VEC = 16 / sizeof(T), N2 = VEC / 2;
#pragma unroll
    for (int j = 0; j < CPL; ++j) { sx2[j] = make_float2(0.f, 0.f); sxm2[j] = make_float2(0.f, 0.f); best[j] = -INFINITY; bu[j] = -1; }
#pragma unroll 2
    for (int u = l; u < nU; u += LPC) {
        const unsigned char* xp = row + u * 16 + (u >= nUbox ? hoff : 0);
        float2 m2[N2], b2[N2];
        pf_pairs<N2>(mloc + u * VEC, m2);
        pf_pairs<N2>(mb + u * VEC, b2);
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            float2 v[N2];
            pf_unit<T>(xp + j * rowB, v);
            float um = -INFINITY;
#pragma unroll
            for (int i = 0; i < N2; ++i) {
                sx2[j] = f2add(sx2[j], v[i]);
                sxm2[j] = f2fma(v[i], m2[i], sxm2[j]);
                const float2 t = f2add(v[i], b2[i]);
                um = fmaxf(um, fmaxf(t.x, t.y));
            }
            if (um > best[j]) { best[j] = um; bu[j] = u; }  // ascending units: strict > keeps the first maximum
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kPfNTF, 2) pf_fwd_kernel(const __grid_constant__ PfParams P, const __grid_constant__ PfMaps maps) {
    constexpr int VEC = 16 / sizeof(T), N2 = VEC / 2;
    constexpr int NT = kPfNTF, NW = NT / 32;
    constexpr int kGrp = NT / 3;
    constexpr int kMaxHdReg = 16, kMaxCWReg = 8;  // MLP weights prefetched into registers behind the team barrier (hidden <= 16, C <= 256)
    extern __shared__ __align__(128) unsigned char pfsm[];
    unsigned char* const xs = pfsm;
    float* const fsm = reinterpret_cast<float*>(pfsm + P.xbytes);
    uint64_t* const full = reinterpret_cast<uint64_t*>(fsm + P.fo.bar);  // [kPfNCH] slice chunks, then the mask rows
    uint64_t* const mbm = full + kPfNCH;
    float* const red = fsm + P.fo.red;
    float* const wk = fsm + P.fo.wk;
    float* const s_avg = fsm + P.fo.avg;
    float* const s_mx = fsm + P.fo.mx;
    float* const s_ha = fsm + P.fo.ha;
    float* const s_hm = fsm + P.fo.hm;
    float2* const s_q2 = reinterpret_cast<float2*>(fsm + P.fo.q2);    // (q, q) per channel
    float4* const s_AB = reinterpret_cast<float4*>(fsm + P.fo.AB4);   // (A, A, B, B) per channel
    float* const pst = fsm + P.fo.pst;                                // [4][C] pooled sums of the whole sample
    float* const mloc = fsm + P.fo.mloc;
    float* const mb = fsm + P.fo.mb;      // 0 where m > 0.5, -inf elsewhere
    float* const aloc = fsm + P.fo.aloc;
    float* const mraw = fsm + P.fo.mraw;  // raw mask rows (own + halo), staged by a bulk copy one item ahead
    float* const R = fsm + P.fo.R;        // phase-local: pooling stage | conv tile + merge buffers
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int st = blockIdx.x / P.ST, rk = blockIdx.x - st * P.ST;
    if (tid == 0) {
        for (int k = 0; k < kPfNCH + 1; ++k) mbar_init(&full[k], 1);
        fence_mbar_init();
    }
    __syncthreads();
    unsigned long long pol_once;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_once));
    unsigned nb0 = 0, nb1 = 0, nb2 = 0;  // team barriers passed so far, per level
    unsigned nload = 0, nmask = 0;       // slices / mask stagings consumed so far (parities of the mbarriers)
    int wk_level = -1;

    PfItem cur = pf_decode(P, st, rk);
    if (cur.valid && tid == 0) {
        if (P.stagger_ns > 0 && st > 0) {  // de-synchronise the teams: their HBM phases should not coincide
            unsigned long long t0, t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            const unsigned long long dt = (unsigned long long)st * (unsigned)P.stagger_ns / (unsigned)P.nST;
            do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < dt);
        }
        for (int k = 0; k < kPfNCH; ++k) pf_issue_chunk<sizeof(T)>(P.lv[cur.L], &maps.x[cur.L], xs, full, k, cur.b, cur.r, pol_once);
        if (P.lv[cur.L].sh.has_mask()) pf_issue_mask(P.lv[cur.L], mraw, mbm, cur.b, cur.r);
    }

    for (int it = st; it < P.total_items; it += P.nST) {
        const PfItem nxt = pf_decode(P, it + P.nST, rk);
        if (!cur.valid) {
            if (nxt.valid && tid == 0) {
                for (int k = 0; k < kPfNCH; ++k) pf_issue_chunk<sizeof(T)>(P.lv[nxt.L], &maps.x[nxt.L], xs, full, k, nxt.b, nxt.r, pol_once);
                if (P.lv[nxt.L].sh.has_mask()) pf_issue_mask(P.lv[nxt.L], mraw, mbm, nxt.b, nxt.r);
            }
            cur = nxt;
            continue;
        }
        const PfLevel& lv = P.lv[cur.L];
        const int C = lv.sh.C, S = lv.sh.S, W = lv.sh.W, H = lv.sh.H, Hd = lv.sh.hidden;
        const int b = cur.b, r = cur.r, TS = lv.TS, TSp = (TS + 3) & ~3;
        const int y0 = r * lv.rowsPer;
        const int rows = min(lv.rowsPer, H - y0);
        const int nP = rows * W, nU = nP / VEC, p0 = y0 * W, nPl = lv.nP;
        const int TWp = lv.TWp, planeT = lv.planeT;
        const bool has_mask = lv.sh.has_mask(), save = !lv.sh.no_save();
        const size_t bS = (size_t)b * S;
        const unsigned par = nload & 1u;
        const int rowB = lv.nPbox * (int)sizeof(T), nUbox = lv.nPbox / VEC, hstride = C * rowB;
        const int hoff = hstride - nUbox * 16;  // byte offset of the second box row range relative to "unit * 16"
        unsigned* const ctr = lv.ctr + (blockIdx.x - r);
        float* const tb = lv.xchg + (size_t)(blockIdx.x - r) * lv.xstride;  // the team's exchange buffer: [4C+1][TSp], rank-minor
        float* const tile = R;                 // [3][planeT]
        float* const mg = R + 3 * planeT;      // pass-2 merge buffers [WS][3][nPl]; later the conv partials [3][nPl]
        const float invW = 1.0f / (float)W;
        const bool tl = (it == st + P.nST);    // timeline stamps: the CTA's second item (steady state)
        if (tl) stamp(0);

        // ---- phase 0: m of the own pixels from the staged mask rows; sum(m)
        if (wk_level != cur.L) {
            cl_load_weights7<NT>(lv.prm.wsam, lv.sh.k, false, wk);
            wk_level = cur.L;
        }
        float macc = 0.0f;
        if (has_mask) {
            mbar_wait(mbm, nmask & 1u);
            const int roff = (y0 - max(0, y0 - 3)) * W;  // own rows start here in the staging buffer
            for (int i = tid; i < nP; i += NT) {
                float v = pf_mask_val(mraw, lv.mdt, roff + i);
                if (lv.sh.gate_clamp()) v = fminf(fmaxf(v, 0.0f), 1.0f);
                if (lv.sh.sigmoid_mask()) v = sigmoidf_acc(v);
                mloc[i] = v;
                mb[i] = v > 0.5f ? 0.0f : -INFINITY;
                if (save) lv.ctx.m[bS + p0 + i] = v;
                macc += v;
            }
        } else {
            for (int i = tid; i < nP; i += NT) { mloc[i] = 1.0f; mb[i] = 0.0f; }
        }
        const float mtot = block_sum(macc, red);  // (barriers inside: mloc / mb are complete)

        // ---- phase 1: pooling from shared memory as the chunks land.  LPC lanes share CPL channel rows, a lane owns every LPC-th unit.
        if (tl) stamp(1);
#ifdef MGA_TUNING
        if (tl && g_timeline != nullptr) {  // timeline only: separate "slice has landed" from the pooling arithmetic
            for (int k = 0; k < kPfNCH; ++k) mbar_wait(&full[k], par);
            stamp(2);
        }
#endif
        {
            const int LPC = lv.LPC, LPCp = LPC | 1, CPL = lv.CPL, cpw = (32 / LPC) * CPL;
            const int sub = lane / LPC, l = lane - sub * LPC;
            float* const stg = R;  // [4][C][LPCp]
            const int qs = C * LPCp;
            for (int c0 = w * cpw; c0 < C; c0 += NW * cpw) {
                const int c = c0 + sub * CPL;
                if (c < C) {
                    mbar_wait(&full[c / lv.CCH], par);
                    if (CPL == 2) mbar_wait(&full[(c + 1) / lv.CCH], par);
                    const unsigned char* row = xs + (size_t)c * rowB;
                    float sxm[2], sx[2], best[2];
                    int bu[2];
                    if (CPL == 2) {
                        float2 a2[2], am2[2];
                        pf_pool_lane<T, 2>(row, rowB, mloc, mb, l, LPC, nU, nUbox, hoff, a2, am2, best, bu);
                        sx[0] = a2[0].x + a2[0].y; sx[1] = a2[1].x + a2[1].y;
                        sxm[0] = am2[0].x + am2[0].y; sxm[1] = am2[1].x + am2[1].y;
                    } else {
                        float2 a2[1], am2[1];
                        float b1[1];
                        int u1[1];
                        pf_pool_lane<T, 1>(row, rowB, mloc, mb, l, LPC, nU, nUbox, hoff, a2, am2, b1, u1);
                        sx[0] = a2[0].x + a2[0].y; sxm[0] = am2[0].x + am2[0].y; best[0] = b1[0]; bu[0] = u1[0];
                    }
                    for (int j = 0; j < CPL; ++j) {
                        int bpix = -1;
                        if (bu[j] >= 0) {  // the first pixel of that unit that holds the maximum
                            float2 v[N2], b2[N2];
                            pf_unit<T>(row + j * rowB + bu[j] * 16 + (bu[j] >= nUbox ? hoff : 0), v);
                            pf_pairs<N2>(mb + bu[j] * VEC, b2);
#pragma unroll
                            for (int i = N2 - 1; i >= 0; --i) {
                                if (v[i].y + b2[i].y == best[j]) bpix = bu[j] * VEC + 2 * i + 1;
                                if (v[i].x + b2[i].x == best[j]) bpix = bu[j] * VEC + 2 * i;
                            }
                        }
                        float* sp = stg + (c + j) * LPCp + l;
                        sp[0] = sxm[j];
                        sp[qs] = sx[j];
                        sp[2 * qs] = best[j];
                        sp[3 * qs] = __int_as_float(bpix);
                    }
                }
            }
        }
        if (tl) stamp(12);
        __syncthreads();
        if (tl) stamp(13);
        {   // per channel: combine the LPC lane partials in a fixed order and publish them rank-minor in the team buffer
            const int LPC = lv.LPC, LPCp = LPC | 1, qs = C * LPCp;
            const float* const stg = R;
            for (int t = tid; t < 3 * C; t += NT) {
                const int q = t / C, c = t - q * C;
                const float* sp = stg + q * qs + c * LPCp;
                if (q < 2) {
                    float a = 0.0f;
                    for (int j = 0; j < LPC; ++j) a += sp[j];
                    tb[(q * C + c) * TSp + r] = a;
                } else {
                    float best = -INFINITY;
                    int bpix = -1;
                    for (int j = 0; j < LPC; ++j) {
                        const float ob = sp[j];
                        const int oi = __float_as_int(sp[qs + j]);
                        if (oi >= 0 && (bpix < 0 || ob > best || (ob == best && oi < bpix))) { best = ob; bpix = oi; }
                    }
                    tb[(2 * C + c) * TSp + r] = best;
                    tb[(3 * C + c) * TSp + r] = __int_as_float(bpix < 0 ? -1 : p0 + bpix);
                }
            }
            if (tid == 0) tb[4 * C * TSp + r] = mtot;
        }
        if (tl) stamp(3);
        unsigned target;
        if (cur.L == 0) target = (++nb0) * TS; else if (cur.L == 1) target = (++nb1) * TS; else target = (++nb2) * TS;
        pf_team_arrive(ctr);  // #1 (the pooling stage is dead from here on)

        // ---- behind the barrier: conv tile (zero padding + the m plane from the staged mask rows), MLP weights into registers
        for (int i = tid; i < (3 * planeT) / 4; i += NT) reinterpret_cast<float4*>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool wreg = Hd <= kMaxHdReg && C <= 32 * kMaxCWReg;
        float w1r[kMaxCWReg], w2r[kMaxHdReg], b1r = 0.0f, b2r = 0.0f;
        if (wreg) {
            if (w < Hd) {
#pragma unroll
                for (int k = 0; k < kMaxCWReg; ++k) w1r[k] = (lane + 32 * k < C) ? __ldg(lv.prm.w1 + (size_t)w * C + lane + 32 * k) : 0.0f;
                b1r = __ldg(lv.prm.b1 + w);
            }
            if (tid < C) {
#pragma unroll
                for (int j = 0; j < kMaxHdReg; ++j) w2r[j] = (j < Hd) ? __ldg(lv.prm.w2 + (size_t)tid * Hd + j) : 0.0f;
                b2r = __ldg(lv.prm.b2 + tid);
            }
        }
        const float beta = __ldg(lv.prm.beta);
        __syncthreads();
        if (has_mask) {
            const int ys = max(0, y0 - 3), ye = min(H, y0 + rows + 3);
            const int n = (ye - ys) * W;
            for (int i = tid; i < n; i += NT) {
                const int ry = __float2int_rz(((float)i + 0.5f) * invW), cx = i - ry * W;
                float v = pf_mask_val(mraw, lv.mdt, i);
                if (lv.sh.gate_clamp()) v = fminf(fmaxf(v, 0.0f), 1.0f);
                if (lv.sh.sigmoid_mask()) v = sigmoidf_acc(v);
                tile[2 * planeT + (ys + ry - (y0 - 3)) * TWp + 4 + cx] = v;
            }
        }
        const float alpha = softplusf_acc(beta);
        const float k0 = lv.sh.pyramid_multiply() ? 0.0f : 1.0f - alpha, k1 = alpha;
        if (b == 0 && r == 0 && tid == 0) { lv.ctx.consts[0] = k0; lv.ctx.consts[1] = k1; lv.ctx.consts[2] = alpha; lv.ctx.consts[3] = sigmoidf_acc(beta); }
        pf_team_wait(ctr, target);  // #1: the partials of every rank are in L2 (and the m plane is complete)
        // the staged mask rows are consumed: the next item's rows can land (a whole item ahead of their use)
        if (tid == 0 && nxt.valid && P.lv[nxt.L].sh.has_mask()) pf_issue_mask(P.lv[nxt.L], mraw, mbm, nxt.b, nxt.r);

        // ---- phase 2: sums over the ranks (a thread reads one contiguous row of TS values: independent 16-byte loads), descriptors, MLP
        if (tl) stamp(4);
        for (int t = tid; t <= 3 * C; t += NT) {
            const int q = t / C, c = t - q * C;
            if (t == 3 * C) {  // sum(m) of the sample
                float tot = 0.0f;
                const float* rowp = tb + (size_t)4 * C * TSp;
                for (int rr = 0; rr < TS; ++rr) tot += __ldcg(rowp + rr);
                red[32] = tot;
            } else {
                const float4* rb = reinterpret_cast<const float4*>(tb + (size_t)(q * C + c) * TSp);
                const float4* ri = rb + (size_t)C * (TSp / 4);  // (q == 2: the row of pixel indices)
                float a = 0.0f, best = -INFINITY;
                int bidx = -1;
                for (int j0 = 0; j0 < TSp / 4; j0 += 4) {
                    float4 vb[4], vi[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        vb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        vi[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (j0 + j < TSp / 4) {
                            vb[j] = __ldcg(rb + j0 + j);
                            if (q == 2) vi[j] = __ldcg(ri + j0 + j);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float ob[4] = {vb[j].x, vb[j].y, vb[j].z, vb[j].w};
                        const int oi[4] = {__float_as_int(vi[j].x), __float_as_int(vi[j].y), __float_as_int(vi[j].z), __float_as_int(vi[j].w)};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (4 * (j0 + j) + e < TS) {
                                a += ob[e];
                                // ranks own increasing pixel ranges: strict > keeps the first maximum
                                if (oi[e] >= 0 && (bidx < 0 || ob[e] > best)) { best = ob[e]; bidx = oi[e]; }
                            }
                        }
                    }
                }
                if (q < 2) pst[q * C + c] = a;
                else { pst[2 * C + c] = best; pst[3 * C + c] = __int_as_float(bidx); }
            }
        }
        __syncthreads();
        if (tl) stamp(5);
        {
            float use = 0.0f, den = 1.0f;
            if (has_mask) {
                const float tot = red[32];
                use = (tot / (float)S >= lv.sh.tiny_thr) ? 1.0f : 0.0f;
                den = fmaxf(tot, lv.sh.eps);
                if (r == 0 && tid == 0) { lv.ctx.msum[b] = tot; lv.ctx.use[b] = use; lv.ctx.den[b] = den; }
            }
            const float invS = 1.0f / (float)S;
            for (int c = tid; c < C; c += NT) {
                const float sxm = pst[c], sx = pst[C + c], best = pst[2 * C + c];
                const int bidx = __float_as_int(pst[3 * C + c]);
                const float Gm = sx * invS;
                const float A = has_mask ? sxm / den : Gm;
                const float avg = has_mask ? (A * use + Gm * (1.0f - use)) : Gm;
                const bool dead = bidx < 0;  // no pixel with m > 0.5 (masked_cbam.py:118-121)
                const float mx = dead ? Gm : best;
                s_avg[c] = avg;
                s_mx[c] = mx;
                if (r == 0) {
                    const int i = b * C + c;
                    lv.ctx.avg[i] = avg; lv.ctx.mx[i] = mx; lv.ctx.apool[i] = A; lv.ctx.amax[i] = dead ? -1 : bidx;
                }
            }
        }
        __syncthreads();
        for (int j = w; j < Hd; j += NW) {
            float pa = 0.0f, pm = 0.0f;
            if (wreg) {
#pragma unroll
                for (int k = 0; k < kMaxCWReg; ++k) {
                    const int c = lane + 32 * k;
                    if (c < C) { pa = fmaf(w1r[k], s_avg[c], pa); pm = fmaf(w1r[k], s_mx[c], pm); }
                }
            } else {
                const float* wr = lv.prm.w1 + (size_t)j * C;
                for (int c = lane; c < C; c += 32) {
                    const float wv = __ldg(wr + c);
                    pa = fmaf(wv, s_avg[c], pa);
                    pm = fmaf(wv, s_mx[c], pm);
                }
            }
            pa = warp_sum(pa);
            pm = warp_sum(pm);
            if (lane == 0) {
                const float bb = wreg ? b1r : __ldg(lv.prm.b1 + j);
                const float ha = fmaxf(pa + bb, 0.0f), hm = fmaxf(pm + bb, 0.0f);
                s_ha[j] = ha;
                s_hm[j] = hm;
                if (r == 0) { lv.ctx.ha[b * Hd + j] = ha; lv.ctx.hm[b * Hd + j] = hm; }
            }
        }
        __syncthreads();
        {
            const bool add = lv.sh.samcam_add();
            for (int c = tid; c < C; c += NT) {
                float za = 0.0f, zm = 0.0f;
                if (wreg) {
#pragma unroll
                    for (int j = 0; j < kMaxHdReg; ++j)
                        if (j < Hd) { za = fmaf(w2r[j], s_ha[j], za); zm = fmaf(w2r[j], s_hm[j], zm); }
                } else {
                    const float* wr = lv.prm.w2 + (size_t)c * Hd;
                    for (int j = 0; j < Hd; ++j) {
                        const float wv = __ldg(wr + j);
                        za = fmaf(wv, s_ha[j], za);
                        zm = fmaf(wv, s_hm[j], zm);
                    }
                }
                const float bb = wreg ? b2r : __ldg(lv.prm.b2 + c);
                const float s = sigmoidf_acc((za + bb) + (zm + bb));  // b2 enters twice (masked_cbam.py:128)
                if (r == 0) lv.ctx.s[b * C + c] = s;
                const float q = add ? 1.0f : s;
                s_q2[c] = make_float2(q, q);
                // out = x * (A * a + B):  multiply: k0 + k1*s*a ;  add: k0 + k1*(s + a)
                const float A = add ? k1 : k1 * s, Bc = add ? fmaf(k1, s, k0) : k0;
                s_AB[c] = make_float4(A, A, Bc, Bc);
            }
        }
#pragma unroll
        for (int k = 0; k < kPfNCH; ++k) mbar_wait(&full[k], par);  // every thread observes every chunk before it reads across channels
        __syncthreads();

        // ---- phase 3: per pixel max / arg max / sum over channels of x*q, from shared memory.
        //      A warp task = UL consecutive units x (32/UL) channel ranges; WS warps split the channels further.
        if (tl) stamp(6);
        {
            const int UL = lv.UL, CQ = 32 / UL, WS = lv.WS;
            const int nUG = (nU + UL - 1) / UL;
            const int CR = (C + CQ * WS - 1) / (CQ * WS);
            const int cq = lane / UL, ul = lane - cq * UL;
            for (int t = w; t < nUG * WS; t += NW) {
                const int ws = t / nUG, ug = t - ws * nUG;
                const int u = ug * UL + ul;
                const int c_lo = (ws * CQ + cq) * CR, c_hi = min(C, c_lo + CR);
                float2 vs[N2];
                float vmax[VEC];
                int vidx[VEC];
#pragma unroll
                for (int i = 0; i < N2; ++i) vs[i] = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < VEC; ++i) { vmax[i] = -INFINITY; vidx[i] = 0x7fffffff; }
                if (u < nU) {
                    const unsigned char* p = xs + u * 16 + (u >= nUbox ? hoff : 0) + (size_t)c_lo * rowB;
#pragma unroll 4
                    for (int c = c_lo; c < c_hi; ++c, p += rowB) {
                        float2 v[N2];
                        pf_unit<T>(p, v);
                        const float2 q2 = s_q2[c];
#pragma unroll
                        for (int i = 0; i < N2; ++i) {
                            const float2 y = f2mul(v[i], q2);
                            vs[i] = f2add(vs[i], y);
                            if (y.x > vmax[2 * i]) { vmax[2 * i] = y.x; vidx[2 * i] = c; }
                            if (y.y > vmax[2 * i + 1]) { vmax[2 * i + 1] = y.y; vidx[2 * i + 1] = c; }
                        }
                    }
                }
                __syncwarp();
                for (int of = UL; of < 32; of <<= 1) {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const float om = __shfl_xor_sync(0xffffffffu, vmax[i], of);
                        const int oi = __shfl_xor_sync(0xffffffffu, vidx[i], of);
                        if (om > vmax[i] || (om == vmax[i] && oi < vidx[i])) { vmax[i] = om; vidx[i] = oi; }  // torch.max: first maximal channel
                    }
#pragma unroll
                    for (int i = 0; i < N2; ++i) {
                        vs[i].x += __shfl_xor_sync(0xffffffffu, vs[i].x, of);
                        vs[i].y += __shfl_xor_sync(0xffffffffu, vs[i].y, of);
                    }
                }
                if (cq == 0 && u < nU) {
                    float* d = mg + (size_t)ws * 3 * nPl + u * VEC;
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        d[i] = vmax[i];
                        d[nPl + i] = (i & 1) ? vs[i >> 1].y : vs[i >> 1].x;
                        d[2 * nPl + i] = __int_as_float(vidx[i]);
                    }
                }
            }
            __syncthreads();
            const float invC = 1.0f / (float)C;
            for (int p = tid; p < nP; p += NT) {
                float bm = mg[p], bsum = mg[nPl + p];
                int bi = __float_as_int(mg[2 * nPl + p]);
                for (int j = 1; j < WS; ++j) {
                    const float* d = mg + (size_t)j * 3 * nPl + p;
                    const float om = d[0];
                    const int oi = __float_as_int(d[2 * nPl]);
                    bsum += d[nPl];
                    if (om > bm || (om == bm && oi < bi)) { bm = om; bi = oi; }
                }
                const float pavg = bsum * invC;
                const int ry = __float2int_rz(((float)p + 0.5f) * invW), cx = p - ry * W;
                tile[(3 + ry) * TWp + 4 + cx] = bm;
                tile[planeT + (3 + ry) * TWp + 4 + cx] = pavg;
                lv.ctx.pmax[bS + p0 + p] = bm;  // (also the halo exchange with the neighbouring ranks)
                lv.ctx.pavg[bS + p0 + p] = pavg;
                if (save) lv.ctx.idx[bS + p0 + p] = bi;
            }
        }
        if (tl) stamp(7);
        if (cur.L == 0) target = (++nb0) * TS; else if (cur.L == 1) target = (++nb1) * TS; else target = (++nb2) * TS;
        pf_team_arrive(ctr);
        pf_team_wait(ctr, target);  // #2: every rank's rows of the two planes are in L2

        // ---- phase 4: halo rows of [pmax, pavg] from L2, then a = sigmoid(conv7x7([pmax, pavg, m])) for the own rows
        if (tl) stamp(8);
        {
            const int cpr = W / 4, per = 6 * cpr;
            const float icpr = 1.0f / (float)cpr;
            for (int i = tid; i < 2 * per; i += NT) {
                const int pl = i >= per ? 1 : 0, rem = i - pl * per;
                const int hr = __float2int_rz(((float)rem + 0.5f) * icpr), ch = rem - hr * cpr;
                const int tr = hr < 3 ? hr : rows + hr;
                const int yy = y0 - 3 + tr;
                if (yy < 0 || yy >= H) continue;
                const float* src = (pl == 0 ? lv.ctx.pmax : lv.ctx.pavg) + bS + (size_t)yy * W + ch * 4;
                *reinterpret_cast<float4*>(tile + pl * planeT + tr * TWp + 4 + ch * 4) = __ldcg(reinterpret_cast<const float4*>(src));
            }
        }
        __syncthreads();
        if (tl) stamp(9);
        {
            float* cpart = mg;  // [3][nPl]
            const int grp = tid / kGrp, gl = tid - grp * kGrp;
            const int spr = W / 4, nStrips = rows * spr;
            const float ispr = 1.0f / (float)spr;
            if (grp < 3) {
                const float* wv = wk + grp * kMaxK * kMaxK;
                for (int s = gl; s < nStrips; s += kGrp) {
                    const int ry = __float2int_rz(((float)s + 0.5f) * ispr), x0 = (s - ry * spr) * 4;
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
                    strip_conv7(tile + grp * planeT + ry * TWp + x0, TWp, wv, acc);
                    *reinterpret_cast<float4*>(cpart + (size_t)grp * nPl + s * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                }
            }
            __syncthreads();
            for (int p = tid; p < nP; p += NT) {
                const float v = (cpart[p] + cpart[nPl + p]) + cpart[2 * nPl + p];
                const float a = sigmoidf_acc(v);
                aloc[p] = a;
                lv.ctx.a[bS + p0 + p] = a;
            }
        }
        __syncthreads();

        // ---- phase 5: out = x * (A*a + B) from shared memory, chunk by chunk (G divides the channels of a chunk: every thread does
        //      the same number of channels per chunk); a finished chunk is refilled with the next item's slice
        if (tl) stamp(10);
        {
            const int G = lv.G, nUl = lv.nU, jpc = lv.CCH / G;
            const int cg = tid / nUl, u = tid - cg * nUl;
            const bool act = cg < G && u < nU;
            const bool pf_ok = nxt.valid && nxt.L == cur.L;
            float2 av[N2];
#pragma unroll
            for (int i = 0; i < N2; ++i) av[i] = make_float2(0.f, 0.f);
            if (act) pf_pairs<N2>(aloc + u * VEC, av);
            const unsigned char* xp = xs + u * 16 + (u >= nUbox ? hoff : 0) + cg * rowB;
            char* op = reinterpret_cast<char*>(lv.out) + ((size_t)b * C * S + p0 + (size_t)cg * S) * sizeof(T) + (size_t)u * 16;
            const unsigned xstep = (unsigned)(G * rowB);
            const size_t ostep = (size_t)G * S * sizeof(T);
            const float4* abp = s_AB + cg;
            for (int k = 0; k < kPfNCH; ++k) {
                if (act) {
#pragma unroll 4
                    for (int j = 0; j < jpc; ++j) {
                        float2 v[N2];
                        pf_unit<T>(xp, v);
                        const float4 ab = *abp;
                        const float2 A2 = make_float2(ab.x, ab.y), B2 = make_float2(ab.z, ab.w);
#pragma unroll
                        for (int i = 0; i < N2; ++i) v[i] = f2mul(v[i], f2fma(A2, av[i], B2));
                        __stcs(reinterpret_cast<uint4*>(op), pf_pack<T>(v));
                        xp += xstep;
                        op += ostep;
                        abp += G;
                    }
                }
                __syncthreads();
                if (tid == 0) {
                    if (pf_ok) pf_issue_chunk<sizeof(T)>(lv, &maps.x[cur.L], xs, full, k, nxt.b, nxt.r, pol_once);
                    else if (nxt.valid && k == kPfNCH - 1)
                        for (int kk = 0; kk < kPfNCH; ++kk) pf_issue_chunk<sizeof(T)>(P.lv[nxt.L], &maps.x[nxt.L], xs, full, kk, nxt.b, nxt.r, pol_once);
                }
            }
        }
        if (tl) stamp(11);
        ++nload;
        if (has_mask) ++nmask;
        cur = nxt;
    }
}

}  // namespace mga
