// cbam_conv.cuh -- the 7x7 spatial-attention convolution (forward) and its two gradients (backward),
// strip-mined: a work item is (4 consecutive output pixels of one image row, one input plane), the 49
// weights of that plane live in registers and every (plane,row) needs three 128-bit shared-memory reads.
// Needs W % 4 == 0 (true for every YOLO pyramid level); other widths use the generic kernels.
//
//   forward : a = sigmoid(conv7x7([pmax, pavg, m]))                       masked_cbam.py:146-147
//   backward: dpre = k1*T*a*(1-a); dcat_k = conv7x7_T(dpre, W_k); dW_k = corr(cat_k, dpre); sum a*T
#pragma once
#include <cuda.h>  // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "common.cuh"

namespace mga {

constexpr int kConvGroup = 85;  // threads per input plane (3 * 85 = 255 of 256 threads)

struct ConvGeom {
    int RB;      // image rows per CTA
    int TWp;     // tile row pitch in floats = W + 8 (4 zero columns each side keeps 16-byte alignment)
    int rowsT;   // RB + 6
    int nStrips; // RB * W / 4
    int planeT;  // floats per staged plane, rounded up to 128 bytes (TMA destination alignment)
    int use_tma; // 1: the [pmax, pavg, m] tiles (+halo, zero padded) arrive by one 3-D TMA box load per plane
};
inline ConvGeom conv_geom(int W) {
    ConvGeom g;
    g.RB = std::max(1, kConvGroup / (W / 4));
    g.TWp = W + 8;
    g.rowsT = g.RB + kMaxK - 1;
    g.nStrips = g.RB * (W / 4);
    g.planeT = (g.rowsT * g.TWp + 31) & ~31;
    g.use_tma = (g.TWp <= 256 && g.rowsT <= 256) ? 1 : 0;
    return g;
}
struct PlaneMaps {  // tensor maps of the three (B,H,W) fp32 planes, box = (W + 8, RB + 6, 1)
    CUtensorMap m[3];
};

// the three planes of rows [y0-3, y0-3+rowsT) x columns [-4, W+4): one elected thread issues one box load per plane; out-of-image
// rows / columns are zero-filled by the TMA unit -- exactly the conv's zero padding.  Every thread then waits on the mbarrier.
__device__ __forceinline__ void stage_three_tma(float* tile, const ConvGeom& cg, const PlaneMaps& maps, bool has_mask, int b, int y0, uint64_t* bar) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (!has_mask)
        for (int i = threadIdx.x; i < cg.planeT; i += kBlock) tile[2 * cg.planeT + i] = 0.0f;
    __syncthreads();
    if (threadIdx.x == 0) {
        const int np = has_mask ? 3 : 2;
        mbar_expect_tx(bar, (uint32_t)(np * cg.rowsT * cg.TWp * 4));
        for (int pl = 0; pl < np; ++pl) tma_load_3d(tile + pl * cg.planeT, &maps.m[pl], -4, y0 - kMaxK / 2, b, bar);
    }
    mbar_wait(bar, 0);
}

// Stage image rows [y_lo, y_lo+rows) of one (H,W) plane into shared memory as float4 chunks; chunk 0 and the
// last chunk of every row, and rows outside the image, are zero (= the conv zero padding).  `f` maps the
// loaded float4 (and its pixel offset) to the stored value.  Loads are issued four at a time before any store.
template <typename F>
__device__ __forceinline__ void stage_plane4(float* dst, const float* __restrict__ src, int y_lo, int rows, int H, int W, int TWp, F f) {
    const int cpr = TWp / 4;  // chunks per row
    const int total = rows * cpr;
    for (int i0 = threadIdx.x; i0 < total; i0 += 4 * kBlock) {
        float4 v[4];
        int off[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + q * kBlock;
            v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            off[q] = -1;
            if (i < total) {
                const int r = i / cpr, c = i - r * cpr;
                const int yy = y_lo + r;
                if (src != nullptr && yy >= 0 && yy < H && c >= 1 && c < cpr - 1) {
                    off[q] = yy * W + (c - 1) * 4;
                    v[q] = __ldg(reinterpret_cast<const float4*>(src + off[q]));
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + q * kBlock;
            if (i < total) reinterpret_cast<float4*>(dst)[i] = off[q] >= 0 ? f(v[q], off[q]) : v[q];
        }
    }
}

// 4 outputs x 49 taps from a staged plane: out[i] += sum_{r,j} tile[(row0 + r)][x0 + 1 + i + j] * w[r][j]
// (w lives in shared memory: every thread of a plane group reads the same address -> broadcast)
__device__ __forceinline__ void strip_conv7(const float* tile_row0 /* &tile[row0][x0] (16-byte aligned) */, int TWp,
                                            const float* __restrict__ w, float (&acc)[4]) {
#pragma unroll
    for (int r = 0; r < kMaxK; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(tile_row0 + r * TWp);
        const float4 b = *reinterpret_cast<const float4*>(tile_row0 + r * TWp + 4);
        const float4 c = *reinterpret_cast<const float4*>(tile_row0 + r * TWp + 8);
        const float win[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
        for (int j = 0; j < kMaxK; ++j) {
            const float wv = w[r * kMaxK + j];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fmaf(win[1 + i + j], wv, acc[i]);
        }
    }
}

// all three planes' weights into shared memory, zero padded (centred) from k x k to 7 x 7;
// flip = true gives the transposed-conv kernel
__device__ __forceinline__ void load_weights7(const float* __restrict__ wsam, int k, bool flip, float* w /* [3][49] */) {
    const int off = (kMaxK - k) / 2;
    for (int t = threadIdx.x; t < 3 * kMaxK * kMaxK; t += kBlock) {
        const int pl = t / (kMaxK * kMaxK);
        int ii = (t / kMaxK) % kMaxK, jj = t % kMaxK;
        if (flip) { ii = kMaxK - 1 - ii; jj = kMaxK - 1 - jj; }
        ii -= off; jj -= off;
        w[t] = (ii >= 0 && ii < k && jj >= 0 && jj < k) ? __ldg(wsam + (pl * k + ii) * k + jj) : 0.0f;
    }
}

// stage the three planes [pmax, pavg, m] of rows [y_lo, y_lo + rows) with ALL global loads issued before any store
__device__ __forceinline__ void stage_three(float* tile, int planeT, const float* const (&planes)[3], int y_lo, int rows, int H, int W, int TWp) {
    const int cpr = TWp / 4, total = rows * cpr;
    for (int i0 = threadIdx.x; i0 < total; i0 += 2 * kBlock) {
        float4 v[2][3];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int i = i0 + q * kBlock;
#pragma unroll
            for (int pl = 0; pl < 3; ++pl) v[q][pl] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < total) {
                const int r = i / cpr, c = i - r * cpr, yy = y_lo + r;
                if (yy >= 0 && yy < H && c >= 1 && c < cpr - 1) {
                    const int o = yy * W + (c - 1) * 4;
#pragma unroll
                    for (int pl = 0; pl < 3; ++pl)
                        if (planes[pl] != nullptr) v[q][pl] = __ldcg(reinterpret_cast<const float4*>(planes[pl] + o));
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int i = i0 + q * kBlock;
            if (i < total) {
#pragma unroll
                for (int pl = 0; pl < 3; ++pl) reinterpret_cast<float4*>(tile + pl * planeT)[i] = v[q][pl];
            }
        }
    }
}

// ------------------------------------------------------------------ forward: grid (ceil(H/RB), B)
__device__ __forceinline__ void sam_conv4_body(const Shape& sh, const float* __restrict__ wsam, const Ctx& ctx, const ConvGeom& cg,
                                               const PlaneMaps* maps /* nullptr: stage with plain loads */, const Blk blk, float* csm) {
    __shared__ __align__(8) uint64_t bar;
    const int H = sh.H, W = sh.W, S = sh.S, b = blk.y, y0 = blk.x * cg.RB;
    const int planeT = cg.planeT;
    float* tile = csm;                      // [3][rowsT][TWp]
    float* part = csm + 3 * planeT;         // [3][nStrips][4]
    float* wsm = part + 3 * cg.nStrips * 4; // [3][49]
    const int grp = threadIdx.x / kConvGroup, gl = threadIdx.x % kConvGroup;
    const float* const planes[3] = {ctx.pmax + (size_t)b * S, ctx.pavg + (size_t)b * S, sh.has_mask() ? ctx.m + (size_t)b * S : nullptr};
    load_weights7(wsam, sh.k, false, wsm);
    if (cg.use_tma && maps != nullptr) stage_three_tma(tile, cg, *maps, sh.has_mask(), b, y0, &bar);
    else stage_three(tile, planeT, planes, y0 - kMaxK / 2, cg.rowsT, H, W, cg.TWp);
    __syncthreads();
    const float* w = wsm + (grp < 3 ? grp : 0) * kMaxK * kMaxK;
    const int spr = W / 4;
    if (grp < 3) {
        for (int s = gl; s < cg.nStrips; s += kConvGroup) {
            const int ry = s / spr, x0 = (s - ry * spr) * 4;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            if (y0 + ry < H) strip_conv7(tile + grp * planeT + ry * cg.TWp + x0, cg.TWp, w, acc);
            *reinterpret_cast<float4*>(part + ((size_t)grp * cg.nStrips + s) * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < cg.nStrips * 4; e += kBlock) {
        const int s = e / 4, ry = s / spr, xx = (s - ry * spr) * 4 + (e & 3);
        if (y0 + ry >= H) continue;
        const float v = (part[e] + part[cg.nStrips * 4 + e]) + part[2 * cg.nStrips * 4 + e];
        ctx.a[(size_t)b * S + (y0 + ry) * W + xx] = sigmoidf_acc(v);
    }
}
__global__ void __launch_bounds__(kBlock, 4) sam_conv4_kernel(Shape sh, const float* __restrict__ wsam, Ctx ctx, ConvGeom cg,
                                                              const __grid_constant__ PlaneMaps maps) {
    extern __shared__ __align__(128) float csm_fwd[];
    sam_conv4_body(sh, wsam, ctx, cg, &maps, this_block(), csm_fwd);
}

// ------------------------------------------------------------------ backward: grid (ceil(H/RB), B)
__global__ void __launch_bounds__(kBlock, 4) bwd_conv4_kernel(Shape sh, const float* __restrict__ wsam, Ctx ctx, BwdScratch bs, ConvGeom cg,
                                                              const __grid_constant__ PlaneMaps maps) {
    extern __shared__ __align__(128) float csm[];
    __shared__ double red[32];
    __shared__ __align__(8) uint64_t bar;
    const int H = sh.H, W = sh.W, S = sh.S, b = blockIdx.y, y0 = blockIdx.x * cg.RB;
    const int cta = blockIdx.y * gridDim.x + blockIdx.x;
    const int planeT = cg.planeT;
    float* dpre = csm;              // [rowsT][TWp]
    float* cat = csm + planeT;      // [3][rowsT][TWp]
    float* wsm = cat + 3 * planeT;  // [3][49] flipped kernels
    float* dwp = wsm + 3 * kMaxK * kMaxK;  // [21][12][7] dW partials of the (plane,row) thread teams
    const int grp = threadIdx.x / kConvGroup, gl = threadIdx.x % kConvGroup;
    load_weights7(wsam, sh.k, true, wsm);
    const float k1 = sh.gates_only() ? 1.0f : ctx.consts[1];  // gates mode: bs.T already holds dL/da
    const float* ap = ctx.a + (size_t)b * S;
    const float* Tp = bs.T + (size_t)b * S;
    // dpre = k1 * T * a * (1 - a), and the CTA's share of sum_p a_p T_p (own rows only)
    double at_acc = 0.0;
    {
        const int cpr = cg.TWp / 4, total = cg.rowsT * cpr;
        for (int i0 = threadIdx.x; i0 < total; i0 += 2 * kBlock) {
            float4 av[2], tv[2];
            bool in[2], own[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int i = i0 + q * kBlock;
                in[q] = false; own[q] = false;
                av[q] = tv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < total) {
                    const int r = i / cpr, c = i - r * cpr, yy = y0 - kMaxK / 2 + r;
                    if (yy >= 0 && yy < H && c >= 1 && c < cpr - 1) {
                        const int o = yy * W + (c - 1) * 4;
                        av[q] = __ldg(reinterpret_cast<const float4*>(ap + o));
                        tv[q] = __ldg(reinterpret_cast<const float4*>(Tp + o));
                        in[q] = true;
                        own[q] = r >= kMaxK / 2 && r < kMaxK / 2 + cg.RB;
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int i = i0 + q * kBlock;
                if (i >= total) continue;
                float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
                if (in[q]) {
                    d.x = k1 * tv[q].x * av[q].x * (1.0f - av[q].x);
                    d.y = k1 * tv[q].y * av[q].y * (1.0f - av[q].y);
                    d.z = k1 * tv[q].z * av[q].z * (1.0f - av[q].z);
                    d.w = k1 * tv[q].w * av[q].w * (1.0f - av[q].w);
                    if (own[q]) at_acc += ((double)av[q].x * tv[q].x + (double)av[q].y * tv[q].y) + ((double)av[q].z * tv[q].z + (double)av[q].w * tv[q].w);
                }
                reinterpret_cast<float4*>(dpre)[i] = d;
            }
        }
    }
    const float* const planes[3] = {ctx.pmax + (size_t)b * S, ctx.pavg + (size_t)b * S, sh.has_mask() ? ctx.m + (size_t)b * S : nullptr};
    if (cg.use_tma) stage_three_tma(cat, cg, maps, sh.has_mask(), b, y0, &bar);
    else stage_three(cat, planeT, planes, y0 - kMaxK / 2, cg.rowsT, H, W, cg.TWp);
    __syncthreads();
    const float* w = wsm + (grp < 3 ? grp : 0) * kMaxK * kMaxK;

    // conv2d_input: dcat_k = correlation of dpre with the flipped kernel of plane k
    const int spr = W / 4;
    const size_t plane = (size_t)sh.B * S;
    if (grp < 3) {
        for (int s = gl; s < cg.nStrips; s += kConvGroup) {
            const int ry = s / spr, x0 = (s - ry * spr) * 4;
            if (y0 + ry >= H) continue;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            strip_conv7(dpre + ry * cg.TWp + x0, cg.TWp, w, acc);
            *reinterpret_cast<float4*>(bs.dcat + grp * plane + (size_t)b * S + (y0 + ry) * W + x0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
    }
    // conv2d_weight: dW[pl][i][j] = sum_{own pixels} cat[pl][y + i - 3][x + j - 3] * dpre[y][x].
    // A team of 12 threads owns one (plane, kernel row i) pair and keeps the 7 column taps in registers: per strip of 4
    // pixels it reads a 12-float window of cat and 4 values of dpre (4 x 128-bit) for 28 FMAs.
    constexpr int kStride = 3 * kMaxK * kMaxK + 1;
    constexpr int kTeam = 12;
    float* part = bs.convpart + (size_t)cta * kStride;
    {
        const int team = threadIdx.x / kTeam, tl = threadIdx.x % kTeam;
        if (team < 3 * kMaxK) {
            const int pl = team / kMaxK, i7 = team % kMaxK;
            float acc[kMaxK] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int s = tl; s < cg.nStrips; s += kTeam) {
                const int ry = s / spr, x0 = (s - ry * spr) * 4;
                const float4 d = *reinterpret_cast<const float4*>(dpre + (ry + kMaxK / 2) * cg.TWp + x0 + 4);
                const float* cr = cat + pl * planeT + (ry + i7) * cg.TWp + x0;
                const float4 a = *reinterpret_cast<const float4*>(cr);
                const float4 bq = *reinterpret_cast<const float4*>(cr + 4);
                const float4 c = *reinterpret_cast<const float4*>(cr + 8);
                const float win[12] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w, c.x, c.y, c.z, c.w};
                const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                for (int j = 0; j < kMaxK; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[j] = fmaf(win[1 + i + j], dv[i], acc[j]);
            }
#pragma unroll
            for (int j = 0; j < kMaxK; ++j) dwp[(team * kTeam + tl) * kMaxK + j] = acc[j];
        }
    }
    __syncthreads();
    {
        const int k = sh.k, koff = (kMaxK - k) / 2;
        if ((int)threadIdx.x < 3 * kMaxK * kMaxK) {
            const int team = threadIdx.x / kMaxK, j7 = threadIdx.x % kMaxK;  // team = pl * 7 + i7
            const int pl = team / kMaxK, ii = team % kMaxK - koff, jj = j7 - koff;
            if (ii >= 0 && ii < k && jj >= 0 && jj < k) {
                float t = 0.0f;
#pragma unroll
                for (int q = 0; q < kTeam; ++q) t += dwp[(team * kTeam + q) * kMaxK + j7];
                part[(pl * k + ii) * k + jj] = t;
            }
        }
    }
    const double at = block_sum_d(at_acc, red);
    if (threadIdx.x == 0) bs.atpart[cta] = at;
}

}  // namespace mga
