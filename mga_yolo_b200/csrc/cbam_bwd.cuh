// cbam_bwd.cuh -- backward kernels of the split path (closed form, SURVEY.md section 8a "Backward").
//
//   B1 bwd_reduce1   over (x,g): T_p = sum_c g x q_c ; E_c = sum_p g x (a_p | 1) ; Gx_c = sum_p g x
//   B2 bwd_conv      dpre = k1 T a(1-a); dcat = convT(dpre, Wsam); dWsam partials; sum_p a_p T_p
//   B3 bwd_reduce2   over x      : Q_c = sum_p x (dcat1/C + [idx==c] dcat0)        (multiply mode only)
//   B4 bwd_mlp       ds -> dz -> MLP backward -> per-channel coefficients for B5
//   B5 bwd_dx        over (x,g)  : dx, and R_p = sum_c cA_c x -> dmask
//   B6 bwd_wgrad     dW1 db1 dW2 db2 (batch sums), dWsam / dbeta (partial sums)
#pragma once
#include "common.cuh"

namespace mga {

// ------------------------------------------------------------------ B1
// over (x,g): T_p (per pixel) and E_c, Gx_c (per channel, per-tile partials).  Thread mapping: TileMap (common.cuh).
template <typename T, int VEC, int LPT, int UPT>
__global__ void __launch_bounds__(kBlock, VEC == 8 ? MGA_TILE_MINB_16 : MGA_TILE_MINB) bwd_reduce1_kernel(const T* __restrict__ x, const T* __restrict__ g, Shape sh, Ctx ctx,
                                                                            BwdScratch bs, int nT) {
    using TM_ = TileMap<LPT, UPT, VEC>;
    constexpr int TP = TM_::TP;
    __shared__ float sh_t[kWarpsPerBlock][TP];
    const TM_ tm;
    const int b = blockIdx.y, tile = blockIdx.x;
    const int U = sh.S / VEC, C = sh.C;
    const bool multiply = !sh.samcam_add();
    const float* sp = ctx.s + (size_t)b * C;

    float av[UPT][VEC], tacc[UPT][VEC];
#pragma unroll
    for (int k = 0; k < UPT; ++k) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) { av[k][i] = 1.0f; tacc[k][i] = 0.0f; }
        const int u = tm.unit(tile, k);
        if (u < U && multiply) ldf<VEC>(ctx.a + (size_t)b * sh.S + (size_t)u * VEC, av[k]);
    }
    const size_t base = ((size_t)b * C) * sh.S;
    // channels are taken KB at a time: all loads of a batch are issued before anything consumes them
    constexpr int KB0 = VEC == 8 ? MGA_KB2_16 : MGA_KB2;
    constexpr int KB = KB0 / UPT > 0 ? KB0 / UPT : 1;
    // (warp-uniform bound: the shuffles inside need every lane, also when C is not a multiple of the channels per warp load)
    for (int cw = tm.w * TM_::CPW; cw < C; cw += TM_::kChanStep * KB) {
        const int c0 = cw + tm.sub;
        RawV xr[KB][UPT], gr[KB][UPT];  // held raw (zero bits = 0.0 in every element type), unpacked at use
        float e[KB], gxs[KB];
#pragma unroll
        for (int kc = 0; kc < KB; ++kc) {
            const int c = c0 + kc * TM_::kChanStep;
#pragma unroll
            for (int k = 0; k < UPT; ++k) {
                const int u = tm.unit(tile, k);
                if (u < U && c < C) {
                    xr[kc][k] = ldraw<T, VEC, kLdKeepL2>(x + base + (size_t)c * sh.S + (size_t)u * VEC);
                    gr[kc][k] = ldraw<T, VEC, kLdKeepL2>(g + base + (size_t)c * sh.S + (size_t)u * VEC);
                } else {
                    xr[kc][k].t = make_uint4(0u, 0u, 0u, 0u);
                    gr[kc][k].t = make_uint4(0u, 0u, 0u, 0u);
                }
            }
        }
#pragma unroll
        for (int kc = 0; kc < KB; ++kc) {
            const int c = c0 + kc * TM_::kChanStep;
            const float q = (multiply && c < C) ? __ldg(sp + c) : 1.0f;
            e[kc] = 0.0f; gxs[kc] = 0.0f;
#pragma unroll
            for (int k = 0; k < UPT; ++k) {
                float xv[VEC], gv[VEC];
                unpackv<T, VEC>(xr[kc][k], xv);
                unpackv<T, VEC>(gr[kc][k], gv);
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const float gx = gv[i] * xv[i];
                    tacc[k][i] = fmaf(gx, q, tacc[k][i]);
                    e[kc] = fmaf(gx, av[k][i], e[kc]);
                    gxs[kc] += gx;
                }
            }
        }
#pragma unroll
        for (int kc = 0; kc < KB; ++kc) {
            e[kc] = group_sum<LPT>(e[kc]);
            gxs[kc] = group_sum<LPT>(gxs[kc]);
        }
        if (tm.ul == 0) {
#pragma unroll
            for (int kc = 0; kc < KB; ++kc) {
                const int c = c0 + kc * TM_::kChanStep;
                if (c < C) {
                    const size_t o = ((size_t)b * nT + tile) * C + c;
                    bs.epart[o] = e[kc];
                    bs.gxpart[o] = gxs[kc];
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < UPT; ++k)
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float t = cross_group_sum<LPT>(tacc[k][i]);
            if (tm.sub == 0) sh_t[tm.w][tm.slot(k, i)] = t;
        }
    __syncthreads();
    for (int e = threadIdx.x; e < TP; e += kBlock) {
        const int p = tile * TP + e;
        if (p >= sh.S) continue;
        float t = 0.0f;
#pragma unroll
        for (int j = 0; j < kWarpsPerBlock; ++j) t += sh_t[j][e];
        bs.T[(size_t)b * sh.S + p] = t;
    }
}

// ------------------------------------------------------------------ B2
constexpr int kBT_W = 32, kBT_H = 8;
__global__ void __launch_bounds__(kBlock) bwd_conv_kernel(Shape sh, const float* __restrict__ wsam, Ctx ctx, BwdScratch bs) {
    constexpr int TW = kBT_W + kMaxK - 1, TH = kBT_H + kMaxK - 1;
    __shared__ float dpre[TH][TW + 1];
    __shared__ float cat[3][TH][TW + 1];
    __shared__ float wk[3 * kMaxK * kMaxK];
    __shared__ double red[32];
    const int k = sh.k, pad = k / 2, H = sh.H, W = sh.W, S = sh.S;
    const int b = blockIdx.z, x0 = blockIdx.x * kBT_W, y0 = blockIdx.y * kBT_H;
    const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const float k1 = sh.gates_only() ? 1.0f : ctx.consts[1];
    const bool has_mask = sh.has_mask();
    const float* planes[3] = {ctx.pmax + (size_t)b * S, ctx.pavg + (size_t)b * S, ctx.m + (size_t)b * S};
    const float* Tp = bs.T + (size_t)b * S;
    const float* ap = ctx.a + (size_t)b * S;

    for (int i = threadIdx.x; i < 3 * k * k; i += kBlock) wk[i] = wsam[i];
    const int tw = kBT_W + k - 1, th = kBT_H + k - 1;
    double at_acc = 0.0;
    for (int i = threadIdx.x; i < tw * th; i += kBlock) {
        const int r = i / tw, c = i % tw;
        const int yy = y0 + r - pad, xx = x0 + c - pad;
        const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
        float dp = 0.0f;
        if (in) {
            const float a = ap[yy * W + xx], t = Tp[yy * W + xx];
            dp = k1 * t * a * (1.0f - a);
            const bool own = r >= pad && r < pad + kBT_H && c >= pad && c < pad + kBT_W;
            if (own) at_acc += (double)a * (double)t;
        }
        dpre[r][c] = dp;
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) cat[pl][r][c] = (in && (pl < 2 || has_mask)) ? planes[pl][yy * W + xx] : 0.0f;
    }
    __syncthreads();

    // conv2d_input: dcat[pl][y][x] = sum_ij dpre[y - i + pad][x - j + pad] * W[pl][i][j]
    const int tx = threadIdx.x % kBT_W, ty = threadIdx.x / kBT_W;
    const int ox = x0 + tx, oy = y0 + ty;
    if (ox < W && oy < H) {
        float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) {
                const float dp = dpre[ty + 2 * pad - i][tx + 2 * pad - j];
                d0 = fmaf(dp, wk[(0 * k + i) * k + j], d0);
                d1 = fmaf(dp, wk[(1 * k + i) * k + j], d1);
                d2 = fmaf(dp, wk[(2 * k + i) * k + j], d2);
            }
        const size_t o = (size_t)b * S + oy * W + ox;
        const size_t plane = (size_t)sh.B * S;
        bs.dcat[o] = d0;
        bs.dcat[plane + o] = d1;
        bs.dcat[2 * plane + o] = d2;
    }
    // conv2d_weight: dW[pl][i][j] = sum_{y,x in tile} cat[pl][y + i - pad][x + j - pad] * dpre[y][x]
    const int ntap = 3 * k * k;
    float* part = bs.convpart + (size_t)cta * (3 * kMaxK * kMaxK + 1);
    if ((int)threadIdx.x < ntap) {
        const int pl = threadIdx.x / (k * k), i = (threadIdx.x / k) % k, j = threadIdx.x % k;
        float acc = 0.0f;
        for (int yy = 0; yy < kBT_H; ++yy)
            for (int xx = 0; xx < kBT_W; ++xx) acc = fmaf(cat[pl][yy + i][xx + j], dpre[yy + pad][xx + pad], acc);
        part[threadIdx.x] = acc;
    }
    const double at = block_sum_d(at_acc, red);
    if (threadIdx.x == 0) bs.atpart[cta] = at;
}

// ------------------------------------------------------------------ B3 (multiply mode)
// over x: Q_c = sum_p x (dcat1/C + [idx == c] dcat0), per-tile partials.  Thread mapping: TileMap.
template <typename T, int VEC, int LPT, int UPT>
__global__ void __launch_bounds__(kBlock, VEC == 8 ? MGA_TILE_MINB_16 : MGA_TILE_MINB) bwd_reduce2_kernel(const T* __restrict__ x, Shape sh, Ctx ctx, BwdScratch bs, int nT) {
    using TM_ = TileMap<LPT, UPT, VEC>;
    const TM_ tm;
    const int b = blockIdx.y, tile = blockIdx.x;
    const int U = sh.S / VEC, C = sh.C;
    const size_t plane = (size_t)sh.B * sh.S;
    float d0[UPT][VEC], d1[UPT][VEC];
    int ix[UPT][VEC];
    const float invC = 1.0f / (float)C;
#pragma unroll
    for (int k = 0; k < UPT; ++k) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) { d0[k][i] = 0.0f; d1[k][i] = 0.0f; ix[k][i] = -1; }
        const int u = tm.unit(tile, k);
        if (u < U) {
            const size_t o = (size_t)b * sh.S + (size_t)u * VEC;
            ldf<VEC>(bs.dcat + o, d0[k]);
            ldf<VEC>(bs.dcat + plane + o, d1[k]);
            ldi<VEC>(ctx.idx + o, ix[k]);
#pragma unroll
            for (int i = 0; i < VEC; ++i) d1[k][i] *= invC;
        }
    }
    const size_t base = ((size_t)b * C) * sh.S;
    constexpr int KB0 = VEC == 8 ? MGA_KB1_16 : MGA_KB1;
    constexpr int KB = KB0 / UPT > 0 ? KB0 / UPT : 1;
    // (warp-uniform bound: the shuffles inside need every lane, also when C is not a multiple of the channels per warp load)
    for (int cw = tm.w * TM_::CPW; cw < C; cw += TM_::kChanStep * KB) {
        const int c0 = cw + tm.sub;
        RawV xr[KB][UPT];  // held raw, unpacked at use
        float qv[KB];
#pragma unroll
        for (int kc = 0; kc < KB; ++kc) {
            const int c = c0 + kc * TM_::kChanStep;
#pragma unroll
            for (int k = 0; k < UPT; ++k) {
                const int u = tm.unit(tile, k);
                if (u < U && c < C) xr[kc][k] = ldraw<T, VEC, kLdKeepL2>(x + base + (size_t)c * sh.S + (size_t)u * VEC);
                else xr[kc][k].t = make_uint4(0u, 0u, 0u, 0u);
            }
        }
#pragma unroll
        for (int kc = 0; kc < KB; ++kc) {
            const int c = c0 + kc * TM_::kChanStep;
            qv[kc] = 0.0f;
#pragma unroll
            for (int k = 0; k < UPT; ++k) {
                float xv[VEC];
                unpackv<T, VEC>(xr[kc][k], xv);
#pragma unroll
                for (int i = 0; i < VEC; ++i) qv[kc] = fmaf(xv[i], d1[k][i] + (ix[k][i] == c ? d0[k][i] : 0.0f), qv[kc]);
            }
        }
#pragma unroll
        for (int kc = 0; kc < KB; ++kc) qv[kc] = group_sum<LPT>(qv[kc]);
        if (tm.ul == 0) {
#pragma unroll
            for (int kc = 0; kc < KB; ++kc) {
                const int c = c0 + kc * TM_::kChanStep;
                if (c < C) bs.qpart[((size_t)b * nT + tile) * C + c] = qv[kc];
            }
        }
    }
}

// ------------------------------------------------------------------ B3' (large planes): sum the per-tile partial rows
// grid (ceil(C/32), B): lanes = 32 consecutive channels (coalesced rows of the (B,nT,C) partial arrays), the 8 warps split the
// tiles, their sums meet in shared memory and are added in warp order (deterministic).  Output psum (3,B,C) = [e | gx | q].
__global__ void __launch_bounds__(kBlock) bwd_partsum_kernel(Shape sh, BwdScratch bs, int nT, int nT2) {
    __shared__ float part[3][kWarpsPerBlock][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane, b = blockIdx.y, C = sh.C;
    float e = 0.0f, gx = 0.0f, q = 0.0f;
    if (c < C) {
#pragma unroll 4
        for (int t = w; t < nT; t += kWarpsPerBlock) {
            const size_t o = ((size_t)b * nT + t) * C + c;
            e += __ldg(bs.epart + o);
            gx += __ldg(bs.gxpart + o);
        }
#pragma unroll 4
        for (int t = w; t < nT2; t += kWarpsPerBlock) q += __ldg(bs.qpart + ((size_t)b * nT2 + t) * C + c);
    }
    part[0][w][lane] = e; part[1][w][lane] = gx; part[2][w][lane] = q;
    __syncthreads();
    if (w < 3 && c < C) {
        float t = 0.0f;
#pragma unroll
        for (int j = 0; j < kWarpsPerBlock; ++j) t += part[w][j][lane];
        bs.psum[(size_t)w * sh.B * C + (size_t)b * C + c] = t;
    }
}

// ------------------------------------------------------------------ B4 (one CTA per sample)
__global__ void __launch_bounds__(kBlock) bwd_mlp_kernel(Shape sh, mga_cbam_params prm, Ctx ctx, BwdScratch bs, int nT, int nT2) {
    extern __shared__ float smem[];
    __shared__ float red[32];
    __shared__ double redd[32];
    const int C = sh.C, Hd = sh.hidden, b = blockIdx.x;
    float* s_dz = smem;          // C
    float* s_dha = s_dz + C;     // Hd
    float* s_dhm = s_dha + Hd;   // Hd
    const bool multiply = !sh.samcam_add();
    const bool has_mask = sh.has_mask();
    const float k1 = sh.gates_only() ? 1.0f : ctx.consts[1];  // gates mode: epart (one tile) already holds dL/ds
    double gx_tot = 0.0, se_tot = 0.0;
    {
        // sum the per-tile partials: thread = (channel, tile part) so that every thread has independent loads in flight
        float* s_e = s_dhm + Hd;       // [parts][C]
        float* s_g = s_e + kBlock;     // [parts][C]
        float* s_q = s_g + kBlock;     // [parts][C]
        const int parts = C < kBlock ? kBlock / C : 1;
        for (int c0 = 0; c0 < C; c0 += kBlock) {
            const int c = c0 + (int)threadIdx.x % (C < kBlock ? C : kBlock), part = C < kBlock ? (int)threadIdx.x / C : 0;
            float e = 0.0f, q = 0.0f, gxs = 0.0f;
            if (c < C && part < parts) {
#pragma unroll 4
                for (int t = part; t < nT; t += parts) {  // partials of bwd_reduce1 (nT tiles)
                    const size_t o = ((size_t)b * nT + t) * C + c;
                    e += __ldg(bs.epart + o);
                    gxs += __ldg(bs.gxpart + o);
                }
                if (multiply) {
#pragma unroll 4
                    for (int t = part; t < nT2; t += parts) q += __ldg(bs.qpart + ((size_t)b * nT2 + t) * C + c);  // bwd_reduce2 (nT2 tiles)
                }
            }
            if (part < parts && c < C) { s_e[part * (C < kBlock ? C : kBlock) + (c - c0)] = e; s_g[part * (C < kBlock ? C : kBlock) + (c - c0)] = gxs; s_q[part * (C < kBlock ? C : kBlock) + (c - c0)] = q; }
            __syncthreads();
            const int cc = c0 + (int)threadIdx.x;
            if ((int)threadIdx.x < (C < kBlock ? C : kBlock) && cc < C) {
                const int pitch = C < kBlock ? C : kBlock;
                float es = 0.0f, qs = 0.0f, gs = 0.0f;
                for (int pp = 0; pp < parts; ++pp) { es += s_e[pp * pitch + threadIdx.x]; gs += s_g[pp * pitch + threadIdx.x]; qs += s_q[pp * pitch + threadIdx.x]; }
                const float s = ctx.s[b * C + cc];
                const float ds = k1 * es + qs;
                const float dz = ds * s * (1.0f - s);
                s_dz[cc] = dz;
                bs.dz[b * C + cc] = dz;
                gx_tot += (double)gs;
                if (!multiply) se_tot += (double)s * (double)es;  // add mode: sum_c s_c sum_p g x
            }
            __syncthreads();
        }
    }
    // per-sample piece of d alpha: sum g x gate - [pyramid add] sum g x, minus the part B2 owns (sum_p a_p T_p)
    const double gsum = block_sum_d(gx_tot, redd);
    const double ssum = block_sum_d(se_tot, redd);
    if (threadIdx.x == 0) bs.alphapart[b] = ssum - (sh.pyramid_multiply() ? 0.0 : gsum);
    __syncthreads();

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int j = w; j < Hd; j += kWarpsPerBlock) {
        float acc = 0.0f;
#pragma unroll 8
        for (int c = lane; c < C; c += 32) acc = fmaf(s_dz[c], __ldg(prm.w2 + (size_t)c * Hd + j), acc);
        acc = warp_sum(acc);
        if (lane == 0) {
            const float da = ctx.ha[b * Hd + j] > 0.0f ? acc : 0.0f;
            const float dm = ctx.hm[b * Hd + j] > 0.0f ? acc : 0.0f;
            s_dha[j] = da;
            s_dhm[j] = dm;
            bs.dha[b * Hd + j] = da;
            bs.dhm[b * Hd + j] = dm;
        }
    }
    __syncthreads();
    const float use = has_mask ? ctx.use[b] : 0.0f;
    const float den = has_mask ? ctx.den[b] : 1.0f;
    const float pass = (has_mask && ctx.msum[b] >= sh.eps) ? 1.0f : 0.0f;  // clamp_min backward
    const float invS = 1.0f / (float)sh.S;
    float kacc = 0.0f;
    for (int c = threadIdx.x; c < C; c += kBlock) {
        float davg = 0.0f, dmx = 0.0f;
#pragma unroll 8
        for (int j = 0; j < Hd; ++j) {
            const float wv = __ldg(prm.w1 + (size_t)j * C + c);
            davg = fmaf(s_dha[j], wv, davg);
            dmx = fmaf(s_dhm[j], wv, dmx);
        }
        const int i = b * C + c;
        const bool dead = has_mask && ctx.amax[i] < 0;
        const float cA = has_mask ? use * davg / den : 0.0f;
        bs.cA[i] = cA;
        bs.cG[i] = ((1.0f - use) * davg + (dead ? dmx : 0.0f)) * invS;
        bs.cM[i] = dead ? 0.0f : dmx;
        kacc = fmaf(cA, ctx.apool[i] * pass, kacc);
    }
    const float ksum = block_sum(kacc, red);
    if (threadIdx.x == 0) bs.kb[b] = ksum;
}

// ------------------------------------------------------------------ B5
// over (x,g): dx (streaming store) and R_p = sum_c cA_c x -> dmask.  Thread mapping: TileMap with UPT = 1.
template <typename T, int VEC, int LPT>
// (16-bit: 48 per-pixel registers + two channels of raw (x, g) loads need ~80 registers: three CTAs per SM instead of four)
__global__ void __launch_bounds__(kBlock, VEC == 8 ? 3 : MGA_TILE_MINB) bwd_dx_kernel(const T* __restrict__ x, const T* __restrict__ g, const void* __restrict__ mask,
                                                                       int mdt, T* __restrict__ dx, void* __restrict__ dmask, Shape sh, Ctx ctx,
                                                                       BwdScratch bs) {
    using TM_ = TileMap<LPT, 1, VEC>;
    constexpr int TP = TM_::TP;
    __shared__ float sh_r[kWarpsPerBlock][TP];
    const TM_ tm;
    const int b = blockIdx.y, tile = blockIdx.x;
    const int U = sh.S / VEC, C = sh.C;
    const int u = tm.unit(tile, 0);
    const bool act = u < U;
    const bool add = sh.samcam_add();
    const bool has_mask = sh.has_mask();
    // gates mode: no direct g term; with an upstream feature gradient in place of g (MGA_GATES_ACC) it passes through with weight 1
    const float k0 = sh.gates_only() ? (sh.gates_acc() ? 1.0f : 0.0f) : ctx.consts[0], k1 = sh.gates_only() ? 0.0f : ctx.consts[1];
    const size_t plane = (size_t)sh.B * sh.S;

    float av[VEC], d0[VEC], d1[VEC], mv[VEC], racc[VEC];
    int ix[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) { av[i] = 0.0f; d0[i] = 0.0f; d1[i] = 0.0f; mv[i] = 0.0f; racc[i] = 0.0f; ix[i] = -1; }
    if (act) {
        const size_t o = (size_t)b * sh.S + (size_t)u * VEC;
        ldf<VEC>(ctx.a + o, av);
        ldf<VEC>(bs.dcat + o, d0);
        ldf<VEC>(bs.dcat + plane + o, d1);
        ldi<VEC>(ctx.idx + o, ix);
        if (has_mask) ldf<VEC>(ctx.m + o, mv);
        const float invC = 1.0f / (float)C;
#pragma unroll
        for (int i = 0; i < VEC; ++i) d1[i] *= invC;
    }
    const size_t base = ((size_t)b * C) * sh.S + (size_t)u * VEC;
    if (act) {
        constexpr int KB = VEC == 8 ? MGA_KB2_16 : MGA_KB2;  // channels per batch: 2*KB independent 128-bit loads in flight per thread, held raw
        for (int c0 = tm.chan0(); c0 < C; c0 += TM_::kChanStep * KB) {
            RawV xr[KB], gr[KB];
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                const int c = c0 + k * TM_::kChanStep;
                if (c < C) {
                    xr[k] = ldraw<T, VEC, kLdLastUse>(x + base + (size_t)c * sh.S);
                    gr[k] = ldraw<T, VEC, kLdLastUse>(g + base + (size_t)c * sh.S);
                }
            }
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                const int c = c0 + k * TM_::kChanStep;
                if (c >= C) continue;
                const int bc = b * C + c;
                float ov[VEC], xvk[VEC], gvk[VEC];
                unpackv<T, VEC>(xr[k], xvk);
                unpackv<T, VEC>(gr[k], gvk);
                const float s = __ldg(ctx.s + bc), cA = __ldg(bs.cA + bc), cG = __ldg(bs.cG + bc), cM = __ldg(bs.cM + bc);
                const int am = __ldg(ctx.amax + bc);
                const float q = add ? 1.0f : s;
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const float gate = add ? (s + av[i]) : (s * av[i]);
                    float v = gvk[i] * fmaf(k1, gate, k0);
                    v = fmaf(q, d1[i] + (ix[i] == c ? d0[i] : 0.0f), v);
                    v = fmaf(cA, mv[i], v) + cG;
                    if (u * VEC + i == am) v += cM;
                    ov[i] = v;
                    racc[i] = fmaf(cA, xvk[i], racc[i]);
                }
                stv<T, VEC, true>(dx + base + (size_t)c * sh.S, ov);
            }
        }
    }
    if (!has_mask || dmask == nullptr) return;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float r = cross_group_sum<LPT>(racc[i]);
        if (tm.sub == 0) sh_r[tm.w][tm.slot(0, i)] = r;
    }
    __syncthreads();
    const float kb = bs.kb[b];
    for (int e = threadIdx.x; e < TP; e += kBlock) {
        const int p = tile * TP + e;
        if (p >= sh.S) continue;
        float r = 0.0f;
#pragma unroll
        for (int j = 0; j < kWarpsPerBlock; ++j) r += sh_r[j][e];
        const size_t o = (size_t)b * sh.S + p;
        float dm = (r - kb) + bs.dcat[2 * plane + o];
        if (sh.sigmoid_mask()) {
            const float m = ctx.m[o];
            dm *= m * (1.0f - m);
        }
        if (sh.gate_clamp()) {
            const float raw = mdt == MGA_F32 ? static_cast<const float*>(mask)[o]
                            : mdt == MGA_BF16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(mask)[o])
                                              : __half2float(static_cast<const __half*>(mask)[o]);
            if (!(raw >= 0.0f && raw <= 1.0f)) dm = 0.0f;
        }
        if (mdt == MGA_F32) static_cast<float*>(dmask)[o] = dm;
        else if (mdt == MGA_BF16) static_cast<__nv_bfloat16*>(dmask)[o] = __float2bfloat16_rn(dm);
        else static_cast<__half*>(dmask)[o] = __float2half_rn(dm);
    }
}

// ------------------------------------------------------------------ B6
// blocks [0,nMlpBlocks): 32 MLP-gradient elements per block, summing over the batch;
// blocks after that: one block per spatial-conv tap (and one for d beta), summing the per-CTA partials.
__global__ void __launch_bounds__(kBlock) bwd_wgrad_kernel(Shape sh, Ctx ctx, BwdScratch bs, mga_cbam_grads gp, int nConvCta, int nMlpBlocks,
                                                           int nAlphaPart) {
    __shared__ double red[32];
    const int C = sh.C, Hd = sh.hidden, B = sh.B;
    constexpr int kStride = 3 * kMaxK * kMaxK + 1;
    if ((int)blockIdx.x >= nMlpBlocks) {
        const int t = blockIdx.x - nMlpBlocks;  // tap index, or 3*k*k for d beta
        const int n_sam = 3 * sh.k * sh.k;
        double acc = 0.0;
        if (t < n_sam) {
            for (int r = threadIdx.x; r < nConvCta; r += kBlock) acc += (double)bs.convpart[(size_t)r * kStride + t];
        } else {
            for (int r = threadIdx.x; r < nConvCta; r += kBlock) acc += bs.atpart[r];
            for (int r = threadIdx.x; r < nAlphaPart; r += kBlock) acc += bs.alphapart[r];
        }
        const double tot = block_sum_d(acc, red);
        if (threadIdx.x == 0) {
            if (t < n_sam) gp.wsam[t] = (float)tot;
            else gp.beta[0] = (float)((double)ctx.consts[3] * tot);  // d beta = sigmoid(beta) * d alpha
        }
        return;
    }
    // 32 consecutive MLP-gradient elements per block (lanes: coalesced along the fast index of avg / mx / dz / ha / hm);
    // the 8 warps split the batch, their partial sums meet in shared memory and are added in warp order (deterministic).
    __shared__ float part[kWarpsPerBlock][32];
    const int n_w1 = Hd * C, n_b1 = Hd, n_w2 = C * Hd, n_b2 = C;
    const int total = n_w1 + n_b1 + n_w2 + n_b2;
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    int i = blockIdx.x * 32 + lane;
    float acc = 0.0f;
    if (i < total) {
        if (i < n_w1) {  // dW1[j][c] = sum_b dha[b][j] avg[b][c] + dhm[b][j] mx[b][c]
            const int j = i / C, c = i - j * C;
            for (int b = wq; b < B; b += kWarpsPerBlock)
                acc = fmaf(__ldg(bs.dha + b * Hd + j), __ldg(ctx.avg + b * C + c), fmaf(__ldg(bs.dhm + b * Hd + j), __ldg(ctx.mx + b * C + c), acc));
        } else if (i < n_w1 + n_b1) {
            const int j = i - n_w1;
            for (int b = wq; b < B; b += kWarpsPerBlock) acc += __ldg(bs.dha + b * Hd + j) + __ldg(bs.dhm + b * Hd + j);
        } else if (i < n_w1 + n_b1 + n_w2) {  // dW2[c][j] = sum_b dz[b][c] (ha + hm)[b][j]
            const int q = i - n_w1 - n_b1, c = q / Hd, j = q - c * Hd;
            for (int b = wq; b < B; b += kWarpsPerBlock)
                acc = fmaf(__ldg(bs.dz + b * C + c), __ldg(ctx.ha + b * Hd + j) + __ldg(ctx.hm + b * Hd + j), acc);
        } else {
            const int c = i - n_w1 - n_b1 - n_w2;
            for (int b = wq; b < B; b += kWarpsPerBlock) acc += __ldg(bs.dz + b * C + c);
            acc *= 2.0f;  // b2 enters the forward twice (masked_cbam.py:128)
        }
    }
    part[wq][lane] = acc;
    __syncthreads();
    if (wq == 0 && i < total) {
        float t = 0.0f;
#pragma unroll
        for (int q = 0; q < kWarpsPerBlock; ++q) t += part[q][lane];
        float* dst = i < n_w1 ? gp.w1 + i : (i < n_w1 + n_b1 ? gp.b1 + (i - n_w1) : (i < n_w1 + n_b1 + n_w2 ? gp.w2 + (i - n_w1 - n_b1) : gp.b2 + (i - n_w1 - n_b1 - n_w2)));
        *dst = t;
    }
}

}  // namespace mga
