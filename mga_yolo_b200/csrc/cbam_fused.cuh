// cbam_fused.cuh -- cluster-resident fused forward / backward kernels.
//
// One thread-block CLUSTER per sample.  CTA r of the cluster owns a contiguous slice of the
// sample's pixels for ALL channels and keeps that slice of x in shared memory for the whole
// kernel (up to ~205 KB per CTA, loaded once with 1-D bulk TMA copies, one per channel row).
// Every later pass over x reads shared memory, so HBM sees x exactly once per direction:
//
//   forward : HBM reads x, mask; writes out          (algorithmic 2N + BS)
//   backward: HBM reads x, g, mask; writes dx, dmask (algorithmic 3N + 2BS; g is read a second
//             time, from L2, for the dx pass)
//
// Cross-CTA traffic is tiny: per-channel partial sums through distributed shared memory,
// and the 2/3-plane spatial maps (with conv halo) through global memory/L2, ordered by the
// hardware cluster barrier.  The tiny shared MLP is recomputed by every CTA of the cluster.
//
// Two thread->data mappings are used so that reductions stay thread-private:
//   channel-mapped: thread = (channel, unit-part)  -> per-channel sums need no shuffles
//   unit-mapped   : warp = (unit tile, channel group), lane = one 16-byte unit of pixels
//                   -> per-pixel sums / max are thread-private; channel groups merge via smem
// The row stride of the slice in shared memory is an odd number of 16-byte units, which makes
// both mappings bank-conflict free.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace mga {
namespace cg = cooperative_groups;

constexpr int kFusedMaxGroups = 8;
constexpr int kFB = 512;             // threads per CTA of the fused kernels (1 CTA / SM: shared memory bound)
constexpr int kFW = kFB / 32;
constexpr int kSmemLimit = 232448;  // 227 KB opt-in maximum per CTA on sm_100

struct FusedGeom {
    int CS;            // CTAs per cluster (= per sample)
    int nUmax;         // 16-byte units of pixels per CTA slice (ceil(U / CS))
    int rsU;           // row stride of the slice in 16-byte units (odd)
    int UT, CG;        // unit-mapped passes: unit tiles x channel groups (UT * CG = 16 warps)
    int NJ;            // channel-mapped passes: unit parts per channel (threads = NJ * C when C < 512)
    int NG, rowsPerGroup;  // bulk-load groups (one mbarrier each)
    int tileRows;      // rows of the conv tile incl. halo
    int off_misc, off_chan, off_pix, off_tile, off_xs;
    int plane_floats;  // floats per tile plane (tileRows * (W + 2*pad)); tile order in smem: [mask | pmax | pavg]
    int scratch_floats;  // capacity of the scratch region that aliases the pmax/pavg planes (and may extend past them)
    int smem_bytes;
};

// optional in-kernel timeline (debug): thread 0 of each CTA stamps %globaltimer at phase boundaries
__device__ unsigned long long* g_timeline = nullptr;
__device__ __forceinline__ void stamp(int k) {
    if (g_timeline != nullptr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_timeline[(size_t)blockIdx.x * 16 + k] = t;
        if (k == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            g_timeline[(size_t)blockIdx.x * 16 + 15] = smid;
        }
    }
}

// 16-byte shared-memory read of one unit into fp32 registers
template <typename T, int VEC>
__device__ __forceinline__ void lds_unit(const unsigned char* p, float (&v)[VEC]) {
    if constexpr (sizeof(T) == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        const uint4 t = *reinterpret_cast<const uint4*>(p);
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

__device__ __forceinline__ float load_mask_any(const void* mask, int mdt, size_t i) {
    if (mdt == MGA_F32) return __ldg(static_cast<const float*>(mask) + i);
    if (mdt == MGA_BF16) return __bfloat162float(static_cast<const __nv_bfloat16*>(mask)[i]);
    return __half2float(static_cast<const __half*>(mask)[i]);
}
__device__ __forceinline__ void store_mask_any(void* mask, int mdt, size_t i, float v) {
    if (mdt == MGA_F32) static_cast<float*>(mask)[i] = v;
    else if (mdt == MGA_BF16) static_cast<__nv_bfloat16*>(mask)[i] = __float2bfloat16_rn(v);
    else static_cast<__half*>(mask)[i] = __float2half_rn(v);
}

// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only)
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// slice load with per-thread 16-byte async copies: warp w takes rows w, w+16, ...; lanes stride over the row's units.
// (1-D bulk TMA copies were measured at ~30 ns per copy per SM: fine for 3 KB rows, far too slow for 800 B rows.)
template <typename T>
__device__ __forceinline__ void issue_slice_cp_async(const T* __restrict__ xb, unsigned char* xs, const FusedGeom& gm, int C, int S, int p0,
                                                     int nU) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = warp; c < C; c += kFW) {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(xb + (size_t)c * S + p0);
        unsigned char* dst = xs + (size_t)c * gm.rsU * 16;
        for (int u = lane; u < nU; u += 32) cp_async16(dst + u * 16, src + u * 16);
    }
    cp_async_commit();
}

// issue the bulk loads of this CTA's slice: one row (channel) per copy, NG groups / mbarriers
template <typename T>
__device__ __forceinline__ void issue_slice_loads(const T* __restrict__ xb /* sample base */, unsigned char* xs, uint64_t* bars,
                                                  const FusedGeom& gm, int C, int S, int p0, int nU) {
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32 && nU > 0) {
        if (lane == 0) {
            for (int g = 0; g < gm.NG; ++g) {
                const int rows = min(gm.rowsPerGroup, C - g * gm.rowsPerGroup);
                if (rows > 0) mbar_expect_tx(&bars[g], (uint32_t)rows * nU * 16u);
            }
        }
        __syncwarp();
        for (int c = lane; c < C; c += 32)
            bulk_g2s(xs + (size_t)c * gm.rsU * 16, xb + (size_t)c * S + p0, (uint32_t)nU * 16u, &bars[c / gm.rowsPerGroup]);
    }
}

// block-wide sum over kFB threads; result valid in every thread
__device__ __forceinline__ float fused_block_sum(float v, float* sh /* >= 32 floats */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    float r = (lane < kFW) ? sh[lane] : 0.0f;
    return warp_sum(r);
}

// gather per-channel partials of all ranks: remote (DSMEM) reads are issued as independent loads by
// (channel, rank) threads and staged in local shared memory `stage` [CS][nq][C] when it is large enough.
template <int NQ>
__device__ __forceinline__ const float* stage_remote(cg::cluster_group& cluster, float* xbuf /* local [NQ][C] exchange buffer */,
                                                     float* stage, int stage_cap, int C, int CS) {
    if (NQ * C * CS > stage_cap) return nullptr;
    for (int i = threadIdx.x; i < C * CS; i += kFB) {
        const int c = i % C, rr = i / C;
        const float* rp = cluster.map_shared_rank(xbuf, rr);
        float v[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) v[q] = rp[q * C + c];
#pragma unroll
        for (int q = 0; q < NQ; ++q) stage[(rr * NQ + q) * C + c] = v[q];
    }
    __syncthreads();
    return stage;
}

// The tiny shared MLP (masked_cbam.py:128-129), recomputed by every CTA of the cluster.  Its weights are
// fetched into registers EARLY (before the cluster barrier) so their L2 latency hides behind the barrier.
constexpr int kW1R = 16, kW2R = 32;
struct MlpRegs {
    float w1[kW1R], w2[kW2R], b1[2], b2;
    bool ok;
};
__device__ __forceinline__ void mlp_prefetch(const mga_cbam_params& prm, int C, int Hd, MlpRegs& m) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nJ = (Hd + kFW - 1) / kFW, nC = (C + 31) / 32;
    m.ok = nJ <= 2 && nJ * nC <= kW1R && Hd <= kW2R && C <= kFB;
    if (!m.ok) return;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
        const int j = warp + jj * kFW;
        m.b1[jj] = (jj < nJ && j < Hd) ? __ldg(prm.b1 + j) : 0.0f;
    }
#pragma unroll
    for (int q = 0; q < kW1R; ++q) {
        const int jj = q / nC, cc = q % nC;  // nC is warp-uniform
        const int j = warp + jj * kFW, c = lane + 32 * cc;
        m.w1[q] = (q < nJ * nC && j < Hd && c < C) ? __ldg(prm.w1 + (size_t)j * C + c) : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < kW2R; ++j) m.w2[j] = (tid < C && j < Hd) ? __ldg(prm.w2 + (size_t)tid * Hd + j) : 0.0f;
    m.b2 = tid < C ? __ldg(prm.b2 + tid) : 0.0f;
}
__device__ __forceinline__ void mlp_forward(const mga_cbam_params& prm, const MlpRegs& m, int C, int Hd, const float* s_avg,
                                            const float* s_mx, float* s_ha, float* s_hm, float* s_s) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (m.ok) {
        const int nJ = (Hd + kFW - 1) / kFW, nC = (C + 31) / 32;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            if (jj >= nJ) break;
            const int j = warp + jj * kFW;
            float pa = 0.0f, pm = 0.0f;
#pragma unroll
            for (int q = 0; q < kW1R; ++q) {
                if (q / nC != jj || q >= nJ * nC) continue;
                const int c = lane + 32 * (q % nC);
                if (c < C) { pa = fmaf(m.w1[q], s_avg[c], pa); pm = fmaf(m.w1[q], s_mx[c], pm); }
            }
            pa = warp_sum(pa);
            pm = warp_sum(pm);
            if (lane == 0 && j < Hd) { s_ha[j] = fmaxf(pa + m.b1[jj], 0.0f); s_hm[j] = fmaxf(pm + m.b1[jj], 0.0f); }
        }
        __syncthreads();
        if (tid < C) {
            float za = 0.0f, zm = 0.0f;
#pragma unroll
            for (int j = 0; j < kW2R; ++j)
                if (j < Hd) { za = fmaf(m.w2[j], s_ha[j], za); zm = fmaf(m.w2[j], s_hm[j], zm); }
            s_s[tid] = sigmoidf_acc((za + m.b2) + (zm + m.b2));  // b2 enters twice
        }
        __syncthreads();
        return;
    }
    for (int j = warp; j < Hd; j += kFW) {
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
            s_ha[j] = fmaxf(pa + bb, 0.0f);
            s_hm[j] = fmaxf(pm + bb, 0.0f);
        }
    }
    __syncthreads();
    for (int c = tid; c < C; c += kFB) {
        const float* wr = prm.w2 + (size_t)c * Hd;
        float za = 0.0f, zm = 0.0f;
#pragma unroll 4
        for (int j = 0; j < Hd; ++j) {
            const float wv = __ldg(wr + j);
            za = fmaf(wv, s_ha[j], za);
            zm = fmaf(wv, s_hm[j], zm);
        }
        const float bb = __ldg(prm.b2 + c);
        s_s[c] = sigmoidf_acc((za + bb) + (zm + bb));
    }
    __syncthreads();
}

// geometry of a rank's plane tile: rows [y_lo, y_lo + rows) of the image, TW = W + 2*pad columns (pad zero columns each side)
struct TileGeo {
    int y_lo, rows, TW;
    __device__ __forceinline__ int index(int p, int W) const {
        const int y = p / W;
        return (y - y_lo) * TW + (p - y * W) + kMaxK / 2;
    }
};
__device__ __forceinline__ TileGeo tile_of_rank(int rr, int nPmax, int S, int W) {
    constexpr int pad = kMaxK / 2;
    const int q0 = rr * nPmax, q1 = min(S, q0 + nPmax);
    TileGeo t;
    t.TW = W + 2 * pad;
    if (q1 <= q0) { t.y_lo = 0; t.rows = 0; return t; }
    t.y_lo = q0 / W - pad;
    t.rows = (q1 - 1) / W - q0 / W + 1 + 2 * pad;
    return t;
}

// fill the halo of `nplanes` tile planes (every in-image pixel of the tile that another rank owns) through DSMEM
__device__ __forceinline__ void fill_halo(cg::cluster_group& cluster, float* tile, int plane_floats, int first_plane, int nplanes,
                                          const TileGeo& tg, int p0, int nP, int nPmax, int S, int H, int W) {
    constexpr int pad = kMaxK / 2;
    const int per = tg.rows * W;
    for (int i = threadIdx.x; i < nplanes * per; i += kFB) {
        const int pl = first_plane + i / per, rem = i % per;
        const int yy = tg.y_lo + rem / W, xx = rem % W;
        if (yy < 0 || yy >= H) continue;
        const int p = yy * W + xx;
        if (p >= p0 && p < p0 + nP) continue;
        const int rr = p / nPmax;
        const TileGeo rg = tile_of_rank(rr, nPmax, S, W);
        const float* rt = cluster.map_shared_rank(tile, rr);
        tile[(size_t)pl * plane_floats + (yy - tg.y_lo) * tg.TW + xx + pad] = rt[(size_t)pl * plane_floats + rg.index(p, W)];
    }
}

// ================================================================== forward
template <typename T>
__global__ void __launch_bounds__(kFB, 1) fused_fwd_kernel(const T* __restrict__ x, const void* __restrict__ mask, int mdt,
                                                           T* __restrict__ out, Shape sh, mga_cbam_params prm, Ctx ctx, FusedGeom gm) {
    constexpr int VEC = 16 / sizeof(T);
    extern __shared__ __align__(128) unsigned char fsm[];
    unsigned char* smem = fsm;
    cg::cluster_group cluster = cg::this_cluster();
    const int r = (int)cluster.block_rank();
    const int b = blockIdx.x / gm.CS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = sh.C, S = sh.S, Hd = sh.hidden, W = sh.W, H = sh.H;
    const int U = S / VEC;
    const int u0 = r * gm.nUmax;
    const int nU = max(0, min(gm.nUmax, U - u0));
    const int nP = nU * VEC, p0 = u0 * VEC, nPmax = gm.nUmax * VEC;
    const bool has_mask = sh.has_mask();

    float* red = reinterpret_cast<float*>(smem + gm.off_misc);  // 32
    float* xmsum = red + 32;                                    // 1 (exchange)
    float* wk = red + 40;                                       // 3*7*7 (zero padded to 7x7)
    float* xpool = reinterpret_cast<float*>(smem + gm.off_chan);  // [4][C] exchange: sxm, sx, best, bidx
    float* s_avg = xpool + 4 * C;
    float* s_mx = s_avg + C;
    float* s_s = s_mx + C;
    float* s_ha = s_s + C;
    float* s_hm = s_ha + Hd;
    float* pix = reinterpret_cast<float*>(smem + gm.off_pix);    // [nPmax]: mask (pool) -> arg-max channel -> attention map a
    float* tile = reinterpret_cast<float*>(smem + gm.off_tile);  // planes [mask | pmax | pavg], own pixels in place, halo filled later
    float* scratch = tile + gm.plane_floats;                     // pool partials / gather stage alias the pmax,pavg planes
    unsigned char* xs = smem + gm.off_xs;
    const TileGeo tg = tile_of_rank(r, nPmax, S, W);

    // ---- phase 0: start the slice load, zero the mask plane, mask
    stamp(0);
    issue_slice_cp_async<T>(x + (size_t)b * C * S, xs, gm, C, S, p0, nU);
    stamp(1);
    {   // spatial kernel, zero padded to 7x7 so the conv below has compile-time bounds
        const int k = sh.k, off = (kMaxK - k) / 2;
        for (int i = tid; i < 3 * kMaxK * kMaxK; i += kFB) {
            const int pl = i / (kMaxK * kMaxK), ii = (i / kMaxK) % kMaxK - off, jj = i % kMaxK - off;
            wk[i] = (ii >= 0 && ii < k && jj >= 0 && jj < k) ? __ldg(prm.wsam + (pl * k + ii) * k + jj) : 0.0f;
        }
    }
    float raw[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int i = tid + q * kFB;
        raw[q] = (has_mask && i < nP) ? load_mask_any(mask, mdt, (size_t)b * S + p0 + i) : 0.0f;
    }
    for (int i = tid; i < gm.plane_floats; i += kFB) tile[i] = 0.0f;
    __syncthreads();
    float macc = 0.0f;
    for (int i0 = tid; i0 < nP; i0 += 2 * kFB) {
        if (i0 != tid) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int i = i0 + q * kFB;
                raw[q] = (has_mask && i < nP) ? load_mask_any(mask, mdt, (size_t)b * S + p0 + i) : 0.0f;
            }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int i = i0 + q * kFB;
            if (i >= nP) continue;
            float v = 1.0f;
            if (has_mask) {
                v = raw[q];
                if (sh.gate_clamp()) v = fminf(fmaxf(v, 0.0f), 1.0f);
                if (sh.sigmoid_mask()) v = sigmoidf_acc(v);
                ctx.m[(size_t)b * S + p0 + i] = v;
                tile[tg.index(p0 + i, W)] = v;
                macc += v;
            }
            pix[i] = v;
        }
    }
    {
        const float tot = fused_block_sum(macc, red);
        if (tid == 0) *xmsum = tot;
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- phase 1: channel-mapped pooling over the slice (thread-private sums, no shuffles)
    stamp(2);
    {
        const int NJ = gm.NJ;
        int j = 0, c_first = tid;
        bool active = true;
        if (C < kFB) { j = tid / C; c_first = tid % C; active = j < NJ; }
        const int uj0 = (int)((long long)j * nU / NJ), uj1 = (int)((long long)(j + 1) * nU / NJ);
        for (int c = c_first; c < C && active; c += kFB) {
            float sxm = 0.0f, sx = 0.0f, best = -INFINITY;
            int bidx = -1;
            const unsigned char* row = xs + (size_t)c * gm.rsU * 16;
#pragma unroll 2
            for (int u = uj0; u < uj1; ++u) {
                float v[VEC], mv[VEC];
                lds_unit<T, VEC>(row + u * 16, v);
#pragma unroll
                for (int q = 0; q < VEC / 4; ++q) {
                    const float4 t = *reinterpret_cast<const float4*>(pix + u * VEC + 4 * q);
                    mv[4 * q] = t.x; mv[4 * q + 1] = t.y; mv[4 * q + 2] = t.z; mv[4 * q + 3] = t.w;
                }
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    sx += v[i];
                    sxm = fmaf(v[i], mv[i], sxm);
                    if ((!has_mask || mv[i] > 0.5f) && v[i] > best) { best = v[i]; bidx = p0 + u * VEC + i; }
                }
            }
            scratch[(j * 4 + 0) * C + c] = sxm;
            scratch[(j * 4 + 1) * C + c] = sx;
            scratch[(j * 4 + 2) * C + c] = best;
            scratch[(j * 4 + 3) * C + c] = __int_as_float(bidx);
        }
        __syncthreads();
        for (int c = tid; c < C; c += kFB) {
            float sxm = 0.0f, sx = 0.0f, best = -INFINITY;
            int bidx = -1;
            for (int jj = 0; jj < NJ; ++jj) {  // parts own increasing pixel ranges: first maximum wins ties
                sxm += scratch[(jj * 4 + 0) * C + c];
                sx += scratch[(jj * 4 + 1) * C + c];
                const float ob = scratch[(jj * 4 + 2) * C + c];
                const int oi = __float_as_int(scratch[(jj * 4 + 3) * C + c]);
                if (oi >= 0 && (bidx < 0 || ob > best)) { best = ob; bidx = oi; }
            }
            xpool[0 * C + c] = sxm;
            xpool[1 * C + c] = sx;
            xpool[2 * C + c] = best;
            xpool[3 * C + c] = __int_as_float(bidx);
        }
    }
    MlpRegs mr;
    mlp_prefetch(prm, C, Hd, mr);  // L2 latency of the weights hides behind the cluster barrier
    const float beta = __ldg(prm.beta);
    stamp(3);
    cluster.sync();  // #1: every CTA's xpool / xmsum is complete and visible cluster-wide

    // ---- phase 2: gather through distributed shared memory, then the (redundant) shared MLP
    stamp(4);
    float use = 0.0f, den = 1.0f;
    if (has_mask) {
        float part = (lane < gm.CS) ? *cluster.map_shared_rank(xmsum, lane) : 0.0f;  // CS <= 16 remote loads in parallel
        float tot = 0.0f;
        for (int rr = 0; rr < gm.CS; ++rr) tot += __shfl_sync(0xffffffffu, part, rr);  // rank order: identical in every thread
        use = (tot / (float)S >= sh.tiny_thr) ? 1.0f : 0.0f;
        den = fmaxf(tot, sh.eps);
        if (r == 0 && tid == 0) { ctx.msum[b] = tot; ctx.use[b] = use; ctx.den[b] = den; }
    }
    {
        const float* st = stage_remote<4>(cluster, xpool, scratch, gm.scratch_floats, C, gm.CS);
        const float invS = 1.0f / (float)S;
        for (int c = tid; c < C; c += kFB) {
            float sxm = 0.0f, sx = 0.0f, best = -INFINITY;
            int bidx = -1;
            for (int rr = 0; rr < gm.CS; ++rr) {  // ranks own increasing pixel ranges
                const float* rp = st ? st + (size_t)rr * 4 * C : cluster.map_shared_rank(xpool, rr);
                sxm += rp[0 * C + c];
                sx += rp[1 * C + c];
                const float ob = rp[2 * C + c];
                const int oi = __float_as_int(rp[3 * C + c]);
                if (oi >= 0 && (bidx < 0 || ob > best)) { best = ob; bidx = oi; }
            }
            const float G = sx * invS;
            const float A = has_mask ? sxm / den : G;
            const float avg = has_mask ? (A * use + G * (1.0f - use)) : G;
            const bool dead = bidx < 0;
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
    if (b == 0 && r == 0 && tid == 0) {
        ctx.consts[0] = k0; ctx.consts[1] = k1; ctx.consts[2] = alpha; ctx.consts[3] = sigmoidf_acc(beta);
    }
    __syncthreads();
    mlp_forward(prm, mr, C, Hd, s_avg, s_mx, s_ha, s_hm, s_s);
    if (r == 0) {
        for (int c = tid; c < C; c += kFB) ctx.s[b * C + c] = s_s[c];
        for (int j = tid; j < Hd; j += kFB) { ctx.ha[b * Hd + j] = s_ha[j]; ctx.hm[b * Hd + j] = s_hm[j]; }
    }

    // the scratch that aliased the pmax/pavg planes is dead now: clear them (pad columns / out-of-image rows must read 0)
    for (int i = tid; i < 2 * gm.plane_floats; i += kFB) tile[gm.plane_floats + i] = 0.0f;
    __syncthreads();

    // ---- phase 3: unit-mapped channel max / mean of x*q per pixel, merged in place in the tile planes
    stamp(5);
    const int ut = warp % gm.UT, cgp = warp / gm.UT;
    const int u = ut * 32 + lane;
    const bool act = u < nU;
    const bool use_q = !sh.samcam_add();
    float* tmax = tile + gm.plane_floats;
    float* tavg = tile + 2 * gm.plane_floats;
    int* pidx = reinterpret_cast<int*>(pix);
    {
        float vmax[VEC], vsum[VEC];
        int vidx[VEC], tix[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) { vmax[i] = -INFINITY; vsum[i] = 0.0f; vidx[i] = 0; tix[i] = 0; }
        if (act) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) tix[i] = tg.index(p0 + u * VEC + i, W);
            const unsigned char* col = xs + (size_t)u * 16;
#pragma unroll 4
            for (int c = cgp; c < C; c += gm.CG) {
                float v[VEC];
                lds_unit<T, VEC>(col + (size_t)c * gm.rsU * 16, v);
                const float q = use_q ? s_s[c] : 1.0f;
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const float y = v[i] * q;
                    vsum[i] += y;
                    if (y > vmax[i]) { vmax[i] = y; vidx[i] = c; }
                }
            }
        }
        const float invC = 1.0f / (float)C;
        for (int gsel = 0; gsel < gm.CG; ++gsel) {
            if (cgp == gsel && act) {
                const bool last = gsel == gm.CG - 1;
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const int e = u * VEC + i;
                    float bm = vmax[i], bs = vsum[i];
                    int bi = vidx[i];
                    if (gsel > 0) {
                        const float om = tmax[tix[i]];
                        const int oi = pidx[e];
                        bs += tavg[tix[i]];
                        if (!(bm > om || (bm == om && bi < oi))) { bm = om; bi = oi; }  // torch.max: first maximal channel
                    }
                    if (last) {
                        bs = bs / (float)C;
                        const size_t o = (size_t)b * S + p0 + e;
                        ctx.pmax[o] = bm; ctx.pavg[o] = bs; ctx.idx[o] = bi;  // saved for backward
                    }
                    tmax[tix[i]] = bm; tavg[tix[i]] = bs; pidx[e] = bi;
                }
            }
            __syncthreads();
        }
        (void)invC;
    }
    stamp(6);
    cluster.sync();  // #2: every rank's own pixels of the three planes are in its tile

    // ---- phase 4: halo through DSMEM, then the 7x7 conv over [pmax, pavg, m] for the own pixels
    stamp(7);
    if (nP > 0) {
        fill_halo(cluster, tile, gm.plane_floats, has_mask ? 0 : 1, has_mask ? 3 : 2, tg, p0, nP, nPmax, S, H, W);
    }
    cluster.barrier_arrive();  // #3 (waited at the very end): nobody reads this CTA's tile after its own arrive
    __syncthreads();
    if (nP > 0) {
        // work item = (pixel, plane): 3 consecutive lanes share a pixel, 10 pixels per warp, weights of the lane's plane in registers
        const int pl = lane % 3, slot = lane / 3;
        const int tp = (pl == 0) ? 1 : (pl == 1 ? 2 : 0);  // reference plane order [max, avg, mask] -> tile order [mask | max | avg]
        float wreg[kMaxK * kMaxK];
#pragma unroll
        for (int t = 0; t < kMaxK * kMaxK; ++t) wreg[t] = wk[pl * kMaxK * kMaxK + t];
        constexpr int pad = kMaxK / 2;
        for (int base = warp * 10; base < nP; base += kFW * 10) {
            const int i = base + slot;
            const bool ok = lane < 30 && i < nP;
            float acc = 0.0f;
            if (ok) {
                const float* trow = tile + (size_t)tp * gm.plane_floats + tg.index(p0 + i, W) - pad * tg.TW - pad;
                float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
                for (int ii = 0; ii < kMaxK; ++ii) {
#pragma unroll
                    for (int jj = 0; jj < kMaxK; ++jj) {
                        const float tv = trow[ii * tg.TW + jj];
                        if ((ii * kMaxK + jj) & 1) a1 = fmaf(tv, wreg[ii * kMaxK + jj], a1);
                        else a0 = fmaf(tv, wreg[ii * kMaxK + jj], a0);
                    }
                }
                acc = a0 + a1;
            }
            const float s1 = __shfl_down_sync(0xffffffffu, acc, 1);
            const float s2 = __shfl_down_sync(0xffffffffu, acc, 2);
            if (ok && pl == 0) {
                const float a = sigmoidf_acc((acc + s1) + s2);
                pix[i] = a;
                ctx.a[(size_t)b * S + p0 + i] = a;
            }
        }
    }
    __syncthreads();

    // ---- phase 5: out = x * (k0 + k1 * gate), straight from shared memory
    stamp(8);
    if (act) {
        float av[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) av[i] = pix[u * VEC + i];
        const bool add = sh.samcam_add();
        const unsigned char* col = xs + (size_t)u * 16;
        T* ob = out + ((size_t)b * C) * S + p0 + (size_t)u * VEC;
#pragma unroll 4
        for (int c = cgp; c < C; c += gm.CG) {
            float v[VEC];
            lds_unit<T, VEC>(col + (size_t)c * gm.rsU * 16, v);
            const float sc = s_s[c];
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float gate = add ? (sc + av[i]) : (sc * av[i]);
                v[i] *= fmaf(k1, gate, k0);
            }
            stv<T, VEC, true>(ob + (size_t)c * S, v);
        }
    }
    stamp(9);
    cluster.barrier_wait();  // #3
}

}  // namespace mga
