// cbam_fwd.cuh -- forward kernels of the split (one kernel per phase) path.
//
// Phases (reference lines in mga_yolo/nn/modules/masked_cbam.py):
//   F0 mask_prep     m = sigmoid(clamp(mask)), sum(m), use/den        :93-99 (+ probmaskgater.py:77)
//   F1 cam_pool      per (b,c): sum x*m, sum x, masked max + arg max  :100-101,116-117
//   F2 cam_mlp       blend/fall-back, shared MLP, sigmoid -> s        :102,118-121,128-129
//   F3 sam_reduce    per pixel: max_c / mean_c of x*q                 :135-136
//   F4 sam_conv      a = sigmoid(conv7x7([pmax,pavg,m]))              :146-147
//   F5 rescale       out = x*(k0 + k1*gate)                           :130,148,166-171
#pragma once
#include "common.cuh"

namespace mga {

// ------------------------------------------------------------------ F0
// grid (tiles of kMaskTile pixels, B): writes m and one partial sum per tile; cam_mlp finalises use/den.
constexpr int kMaskTile = 4 * kBlock;
template <typename TM>
__device__ __forceinline__ void mask_prep_body(const TM* __restrict__ mask, const Shape& sh, const Ctx& ctx, const FwdScratch& fs, const Blk blk) {
    __shared__ float red[32];
    const int b = blk.y;
    const int S = sh.S;
    const TM* mp = mask + (size_t)b * S;
    float* mo = ctx.m + (size_t)b * S;
    float acc = 0.0f;
    const int p0 = blk.x * kMaskTile;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int p = p0 + i * kBlock + threadIdx.x;
        if (p < S) {
            float v = to_f<TM>(mp[p]);
            if (sh.gate_clamp()) v = fminf(fmaxf(v, 0.0f), 1.0f);
            if (sh.sigmoid_mask()) v = sigmoidf_acc(v);
            mo[p] = v;
            acc += v;
        }
    }
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) fs.mpart[(size_t)b * blk.gx + blk.x] = tot;
}
template <typename TM>
__global__ void __launch_bounds__(kBlock) mask_prep_kernel(const TM* __restrict__ mask, Shape sh, Ctx ctx, FwdScratch fs) {
    mask_prep_body<TM>(mask, sh, ctx, fs, this_block());
}

// ------------------------------------------------------------------ F1
// TPP threads cooperate on one (b,c) plane (TPP = 32..256, chosen so that a thread owns ~6 units):
// every thread issues its loads four units at a time, so a whole plane is in flight at once.
template <typename T, int VEC, int TPP>
__device__ __forceinline__ void cam_pool_body(const T* __restrict__ x, const Shape& sh, const Ctx& ctx, const FwdScratch& fs, const Blk blk) {
    constexpr int kGroups = kBlock / TPP;
    constexpr int kWarpsPerPlane = TPP / 32;
    __shared__ float red[4][kWarpsPerBlock];
    const int grp = threadIdx.x / TPP, lt = threadIdx.x % TPP;
    const int planes = sh.B * sh.C;
    int pl = blk.x * kGroups + grp;
    const bool active = pl < planes;
    if (!active) pl = planes - 1;  // keep every lane alive for the shuffles
    const int b = pl / sh.C;
    const int U = sh.S / VEC;
    const T* xp = x + (size_t)pl * sh.S;
    const float* mp = ctx.m + (size_t)b * sh.S;
    const bool has_mask = sh.has_mask();

    float sxm = 0.0f, sx = 0.0f, best = -INFINITY;
    int bidx = -1;
    for (int u0 = lt; u0 < U; u0 += 4 * TPP) {
        RawV raw[4];
        float mv[4][VEC];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int u = u0 + q * TPP;
            if (u < U) {
                raw[q] = ldraw<T, VEC, kLdKeepL2>(xp + (size_t)u * VEC);
                if (has_mask) ldf<VEC>(mp + (size_t)u * VEC, mv[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int u = u0 + q * TPP;
            if (u < U) {
                float v[VEC];
                unpackv<T, VEC>(raw[q], v);
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const float m = has_mask ? mv[q][i] : 1.0f;
                    sx += v[i];
                    sxm = fmaf(v[i], m, sxm);
                    if ((!has_mask || m > 0.5f) && v[i] > best) { best = v[i]; bidx = u * VEC + i; }
                }
            }
        }
    }
    // larger value wins; on a tie the lower pixel index wins (first maximum in scan order)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sxm += __shfl_xor_sync(0xffffffffu, sxm, o);
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if ((oi >= 0) && (bidx < 0 || ob > best || (ob == best && oi < bidx))) { best = ob; bidx = oi; }
    }
    if constexpr (kWarpsPerPlane > 1) {
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) { red[0][w] = sx; red[1][w] = sxm; red[2][w] = best; red[3][w] = __int_as_float(bidx); }
        __syncthreads();
        if (lt == 0) {
            const int w0 = grp * kWarpsPerPlane;
            for (int j = 1; j < kWarpsPerPlane; ++j) {
                sx += red[0][w0 + j];
                sxm += red[1][w0 + j];
                const float ob = red[2][w0 + j];
                const int oi = __float_as_int(red[3][w0 + j]);
                if ((oi >= 0) && (bidx < 0 || ob > best || (ob == best && oi < bidx))) { best = ob; bidx = oi; }
            }
        }
    }
    if (active && lt == 0) {
        fs.sxm[pl] = sxm;
        fs.sx[pl] = sx;
        fs.best[pl] = best;
        fs.bidx[pl] = bidx;
    }
}
template <typename T, int VEC, int TPP>
__global__ void __launch_bounds__(kBlock) cam_pool_kernel(const T* __restrict__ x, Shape sh, Ctx ctx, FwdScratch fs) {
    cam_pool_body<T, VEC, TPP>(x, sh, ctx, fs, this_block());
}


// F1, four channels per CTA: the (B,S) mask plane m is fp32 (twice the bytes of a 16-bit feature plane) and cam_pool re-reads it for
// every channel -- here one CTA owns channels c0..c0+3 of a sample and every unit of m is loaded once for the four of them
// (measured at 256x80x80 bf16, batch 128: 279 us -> see profiles/r2_concat_kernel.md).  Needs C % 4 == 0; same results / tie rules.
template <typename T, int VEC>
__global__ void __launch_bounds__(kBlock) cam_pool4_kernel(const T* __restrict__ x, Shape sh, Ctx ctx, FwdScratch fs) {
    __shared__ float red[4][4][kWarpsPerBlock];
    const int b = blockIdx.y, c0 = blockIdx.x * 4;
    const int U = sh.S / VEC;
    const T* xp = x + ((size_t)b * sh.C + c0) * sh.S;
    const float* mp = ctx.m + (size_t)b * sh.S;
    const bool has_mask = sh.has_mask();
    float sxm[4], sx[4], best[4];
    int bidx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { sxm[j] = 0.0f; sx[j] = 0.0f; best[j] = -INFINITY; bidx[j] = -1; }
#pragma unroll 2
    for (int u = threadIdx.x; u < U; u += kBlock) {
        float mv[VEC];
        RawV raw[4];  // held raw, unpacked channel by channel (a 16-bit unit would otherwise occupy 8 registers while in flight)
#pragma unroll
        for (int j = 0; j < 4; ++j) raw[j] = ldraw<T, VEC, kLdStream>(xp + (size_t)j * sh.S + (size_t)u * VEC);
        if (has_mask) ldf<VEC>(mp + (size_t)u * VEC, mv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v[VEC];
            unpackv<T, VEC>(raw[j], v);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float m = has_mask ? mv[i] : 1.0f;
                sx[j] += v[i];
                sxm[j] = fmaf(v[i], m, sxm[j]);
                if ((!has_mask || m > 0.5f) && v[i] > best[j]) { best[j] = v[i]; bidx[j] = u * VEC + i; }  // ascending pixels per thread
            }
        }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sx[j] += __shfl_xor_sync(0xffffffffu, sx[j], o);
            sxm[j] += __shfl_xor_sync(0xffffffffu, sxm[j], o);
            const float ob = __shfl_xor_sync(0xffffffffu, best[j], o);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx[j], o);
            if ((oi >= 0) && (bidx[j] < 0 || ob > best[j] || (ob == best[j] && oi < bidx[j]))) { best[j] = ob; bidx[j] = oi; }
        }
        if (lane == 0) { red[j][0][w] = sx[j]; red[j][1][w] = sxm[j]; red[j][2][w] = best[j]; red[j][3][w] = __int_as_float(bidx[j]); }
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        const int j = threadIdx.x;
        float tsx = 0.0f, tsxm = 0.0f, tb = -INFINITY;
        int ti = -1;
        for (int q = 0; q < kWarpsPerBlock; ++q) {
            tsx += red[j][0][q];
            tsxm += red[j][1][q];
            const float ob = red[j][2][q];
            const int oi = __float_as_int(red[j][3][q]);
            if ((oi >= 0) && (ti < 0 || ob > tb || (ob == tb && oi < ti))) { tb = ob; ti = oi; }
        }
        const int pl = b * sh.C + c0 + j;
        fs.sxm[pl] = tsxm;
        fs.sx[pl] = tsx;
        fs.best[pl] = tb;
        fs.bidx[pl] = ti;
    }
}

// ------------------------------------------------------------------ F2 (one CTA per sample)
__device__ __forceinline__ void cam_mlp_body(const Shape& sh, const mga_cbam_params& prm, const Ctx& ctx, const FwdScratch& fs, int nMaskTiles,
                                             const Blk blk, float* smem /* 2C + 2h floats */) {
    const int C = sh.C, Hd = sh.hidden, b = blk.x;
    float* s_avg = smem;            // C
    float* s_mx = s_avg + C;        // C
    float* s_ha = s_mx + C;         // Hd
    float* s_hm = s_ha + Hd;        // Hd
    const bool has_mask = sh.has_mask();
    float use = 0.0f, den = 1.0f;
    if (has_mask) {
        float tot = 0.0f;
#pragma unroll 8
        for (int t = 0; t < nMaskTiles; ++t) tot += ldc(fs.mpart + (size_t)b * nMaskTiles + t);  // same order in every thread
        use = (tot / (float)sh.S >= sh.tiny_thr) ? 1.0f : 0.0f;
        den = fmaxf(tot, sh.eps);
        if (threadIdx.x == 0) {
            ctx.msum[b] = tot;
            ctx.use[b] = use;
            ctx.den[b] = den;
        }
    }
    const float invS = 1.0f / (float)sh.S;

    for (int c = threadIdx.x; c < C; c += kBlock) {
        const int i = b * C + c;
        const float G = ldc(fs.sx + i) * invS;
        const float A = has_mask ? ldc(fs.sxm + i) / den : G;
        const float avg = has_mask ? (A * use + G * (1.0f - use)) : G;
        const int bi = ldc(fs.bidx + i);
        const bool dead = bi < 0;  // no pixel with m > 0.5 (masked_cbam.py:118-121)
        const float mx = dead ? G : ldc(fs.best + i);
        s_avg[c] = avg;
        s_mx[c] = mx;
        ctx.avg[i] = avg;
        ctx.mx[i] = mx;
        ctx.apool[i] = A;
        ctx.amax[i] = dead ? -1 : bi;
    }
    if (b == 0 && threadIdx.x == 0) {
        const float beta = prm.beta[0];
        const float alpha = softplusf_acc(beta);
        ctx.consts[0] = sh.pyramid_multiply() ? 0.0f : 1.0f - alpha;
        ctx.consts[1] = alpha;
        ctx.consts[2] = alpha;
        ctx.consts[3] = sigmoidf_acc(beta);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // 128-bit weight loads when the rows allow it: a warp's whole row of W1 (a thread's whole row of W2) is then one batch of
    // independent loads = one L2 round trip instead of C/256 (Hd/8) of them -- with one CTA per sample this kernel is pure latency
    const bool vec4 = (C % 4 == 0) && (Hd % 4 == 0) && (((reinterpret_cast<uintptr_t>(prm.w1) | reinterpret_cast<uintptr_t>(prm.w2)) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(smem) & 15) == 0);
    for (int j = w; j < Hd; j += kWarpsPerBlock) {
        const float* wr = prm.w1 + (size_t)j * C;
        float pa = 0.0f, pm = 0.0f;
        if (vec4) {
#pragma unroll 8
            for (int c = 4 * lane; c < C; c += 128) {
                const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + c));
                const float4 av = *reinterpret_cast<const float4*>(s_avg + c);
                const float4 mv = *reinterpret_cast<const float4*>(s_mx + c);
                pa = fmaf(wv.x, av.x, fmaf(wv.y, av.y, fmaf(wv.z, av.z, fmaf(wv.w, av.w, pa))));
                pm = fmaf(wv.x, mv.x, fmaf(wv.y, mv.y, fmaf(wv.z, mv.z, fmaf(wv.w, mv.w, pm))));
            }
        } else {
#pragma unroll 8
            for (int c = lane; c < C; c += 32) {
                const float wv = __ldg(wr + c);
                pa = fmaf(wv, s_avg[c], pa);
                pm = fmaf(wv, s_mx[c], pm);
            }
        }
        pa = warp_sum(pa);
        pm = warp_sum(pm);
        if (lane == 0) {
            const float bb = prm.b1[j];
            const float ha = fmaxf(pa + bb, 0.0f), hm = fmaxf(pm + bb, 0.0f);
            s_ha[j] = ha;
            s_hm[j] = hm;
            ctx.ha[b * Hd + j] = ha;
            ctx.hm[b * Hd + j] = hm;
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kBlock) {
        const float* wr = prm.w2 + (size_t)c * Hd;
        float za = 0.0f, zm = 0.0f;
        if (vec4) {
#pragma unroll 4
            for (int j = 0; j < Hd; j += 4) {
                const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + j));
                za = fmaf(wv.x, s_ha[j], fmaf(wv.y, s_ha[j + 1], fmaf(wv.z, s_ha[j + 2], fmaf(wv.w, s_ha[j + 3], za))));
                zm = fmaf(wv.x, s_hm[j], fmaf(wv.y, s_hm[j + 1], fmaf(wv.z, s_hm[j + 2], fmaf(wv.w, s_hm[j + 3], zm))));
            }
        } else {
#pragma unroll 8
            for (int j = 0; j < Hd; ++j) {
                const float wv = __ldg(wr + j);
                za = fmaf(wv, s_ha[j], za);
                zm = fmaf(wv, s_hm[j], zm);
            }
        }
        const float bb = prm.b2[c];
        const float z = (za + bb) + (zm + bb);  // b2 enters twice (masked_cbam.py:128)
        ctx.s[b * C + c] = sigmoidf_acc(z);
    }
}
__global__ void __launch_bounds__(kBlock) cam_mlp_kernel(Shape sh, mga_cbam_params prm, Ctx ctx, FwdScratch fs, int nMaskTiles) {
    extern __shared__ float dsm_mlp[];
    cam_mlp_body(sh, prm, ctx, fs, nMaskTiles, this_block(), dsm_mlp);
}

// ------------------------------------------------------------------ F3
// per pixel: max / arg max / mean over channels of x*q.  Thread mapping: TileMap (common.cuh).
template <typename T, int VEC, int LPT, int UPT>
__device__ __forceinline__ void sam_reduce_body(const T* __restrict__ x, const Shape& sh, const Ctx& ctx, const Blk blk) {
    using TM_ = TileMap<LPT, UPT, VEC>;
    constexpr int TP = TM_::TP;
    __shared__ float sh_max[kWarpsPerBlock][TP];
    __shared__ float sh_sum[kWarpsPerBlock][TP];
    __shared__ int sh_idx[kWarpsPerBlock][TP];
    const TM_ tm;
    const int b = blk.y, tile = blk.x;
    const int U = sh.S / VEC, C = sh.C;
    const bool use_q = !sh.samcam_add();
    const float* sp = ctx.s + (size_t)b * C;

    float vmax[UPT][VEC], vsum[UPT][VEC];
    int vidx[UPT][VEC];
#pragma unroll
    for (int k = 0; k < UPT; ++k)
#pragma unroll
        for (int i = 0; i < VEC; ++i) { vmax[k][i] = -INFINITY; vsum[k][i] = 0.0f; vidx[k][i] = 0x7fffffff; }
    const T* xp = x + ((size_t)b * C) * sh.S;
    constexpr int KB0 = VEC == 8 ? MGA_KB1_16 : MGA_KB1;
    constexpr int KB = KB0 / UPT > 0 ? KB0 / UPT : 1;
    for (int c0 = tm.chan0(); c0 < C; c0 += TM_::kChanStep * KB) {
        RawV raw[KB][UPT];  // held raw, unpacked at use: KB * UPT loads of 16 bytes in flight per thread for every element type
#pragma unroll
        for (int kc = 0; kc < KB; ++kc) {
            const int c = c0 + kc * TM_::kChanStep;
#pragma unroll
            for (int k = 0; k < UPT; ++k) {
                const int u = tm.unit(tile, k);
                if (c < C && u < U) raw[kc][k] = ldraw<T, VEC, kLdKeepL2>(xp + (size_t)c * sh.S + (size_t)u * VEC);
            }
        }
#pragma unroll
        for (int kc = 0; kc < KB; ++kc) {
            const int c = c0 + kc * TM_::kChanStep;
            if (c >= C) continue;
            const float q = use_q ? ldc(sp + c) : 1.0f;
#pragma unroll
            for (int k = 0; k < UPT; ++k) {
                if (tm.unit(tile, k) >= U) continue;
                float v[VEC];
                unpackv<T, VEC>(raw[kc][k], v);
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const float y = v[i] * q;
                    vsum[k][i] += y;
                    if (y > vmax[k][i]) { vmax[k][i] = y; vidx[k][i] = c; }
                }
            }
        }
    }
    // merge the 32/LPT channel groups of the warp (lanes with the same unit), then the 8 warps through shared memory
#pragma unroll
    for (int k = 0; k < UPT; ++k)
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float bm = vmax[k][i], bsum = vsum[k][i];
            int bi = vidx[k][i];
#pragma unroll
            for (int o = LPT; o < 32; o <<= 1) {
                const float om = __shfl_xor_sync(0xffffffffu, bm, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
                if (om > bm || (om == bm && oi < bi)) { bm = om; bi = oi; }
            }
            if (tm.sub == 0) {
                const int e = tm.slot(k, i);
                sh_max[tm.w][e] = bm;
                sh_sum[tm.w][e] = bsum;
                sh_idx[tm.w][e] = bi;
            }
        }
    __syncthreads();
    for (int e = threadIdx.x; e < TP; e += kBlock) {
        const int p = tile * TP + e;
        if (p >= sh.S) continue;
        float bm = sh_max[0][e], bsum = sh_sum[0][e];
        int bi = sh_idx[0][e];
#pragma unroll
        for (int j = 1; j < kWarpsPerBlock; ++j) {
            const float om = sh_max[j][e];
            const int oi = sh_idx[j][e];
            bsum += sh_sum[j][e];
            if (om > bm || (om == bm && oi < bi)) { bm = om; bi = oi; }  // torch.max: first maximal channel
        }
        const size_t o = (size_t)b * sh.S + p;
        ctx.pmax[o] = bm;
        ctx.pavg[o] = bsum / (float)C;
        ctx.idx[o] = bi;
    }
}
template <typename T, int VEC, int LPT, int UPT>
__global__ void __launch_bounds__(kBlock, VEC == 8 ? MGA_TILE_MINB_16 : MGA_TILE_MINB) sam_reduce_kernel(const T* __restrict__ x, Shape sh, Ctx ctx) {
    sam_reduce_body<T, VEC, LPT, UPT>(x, sh, ctx, this_block());
}

// ------------------------------------------------------------------ F4
// 32x8 output tile per CTA, the three input planes (+halo, zero padded) staged in shared memory.
__global__ void __launch_bounds__(kBlock) sam_conv_kernel(Shape sh, const float* __restrict__ wsam, Ctx ctx) {
    __shared__ float tile[3][kConvTH + kMaxK - 1][kConvTW + kMaxK - 1 + 1];
    __shared__ float wk[3 * kMaxK * kMaxK];
    const int k = sh.k, pad = k / 2, H = sh.H, W = sh.W;
    const int b = blockIdx.z, x0 = blockIdx.x * kConvTW, y0 = blockIdx.y * kConvTH;
    const float* planes[3] = {ctx.pmax + (size_t)b * sh.S, ctx.pavg + (size_t)b * sh.S, ctx.m + (size_t)b * sh.S};
    const bool has_mask = sh.has_mask();
    for (int i = threadIdx.x; i < 3 * k * k; i += kBlock) wk[i] = wsam[i];
    const int tw = kConvTW + k - 1, th = kConvTH + k - 1;
    for (int i = threadIdx.x; i < 3 * tw * th; i += kBlock) {
        const int pl = i / (tw * th), r = (i / tw) % th, c = i % tw;
        const int yy = y0 + r - pad, xx = x0 + c - pad;
        float v = 0.0f;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W && (pl < 2 || has_mask)) v = planes[pl][yy * W + xx];
        tile[pl][r][c] = v;
    }
    __syncthreads();
    const int tx = threadIdx.x % kConvTW, ty = threadIdx.x / kConvTW;
    const int ox = x0 + tx, oy = y0 + ty;
    if (ox < W && oy < H) {
        float acc = 0.0f;
        for (int pl = 0; pl < 3; ++pl)
            for (int i = 0; i < k; ++i)
#pragma unroll 7
                for (int j = 0; j < k; ++j) acc = fmaf(tile[pl][ty + i][tx + j], wk[(pl * k + i) * k + j], acc);
        ctx.a[(size_t)b * sh.S + oy * W + ox] = sigmoidf_acc(acc);
    }
}

// ------------------------------------------------------------------ F5
template <typename T, int VEC>
__global__ void __launch_bounds__(kBlock) rescale_kernel(const T* __restrict__ x, T* __restrict__ out, Shape sh, Ctx ctx) {
    const int U = sh.S / VEC;
    const size_t total = (size_t)sh.B * sh.C * U;
    const float k0 = ctx.consts[0], k1 = ctx.consts[1];
    const bool add = sh.samcam_add();
    for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < total; i += (size_t)gridDim.x * kBlock) {
        const int pl = (int)(i / U), u = (int)(i % U);
        const int b = pl / sh.C;
        float v[VEC], av[VEC];
        ldv<T, VEC, kLdLastUse>(x + i * VEC, v);
        ldf<VEC>(ctx.a + (size_t)b * sh.S + (size_t)u * VEC, av);
        const float s = __ldg(ctx.s + pl);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const float gate = add ? (s + av[e]) : (s * av[e]);
            v[e] = v[e] * fmaf(k1, gate, k0);
        }
        stv<T, VEC, true>(out + i * VEC, v);
    }
}

// per-sample form for the dataflow kernel: CTA `tile` of sample b rescales 8 units per thread (2048 units per CTA)
constexpr int kRescaleUnits = 8 * kBlock;
template <typename T, int VEC>
__device__ __forceinline__ void rescale_tile_body(const T* __restrict__ x, T* __restrict__ out, const Shape& sh, const Ctx& ctx, float k0, float k1,
                                                  int b, int tile) {
    const int U = sh.S / VEC;
    const int total = sh.C * U;  // units of one sample
    const bool add = sh.samcam_add();
    const size_t sbase = (size_t)b * sh.C * sh.S;
#pragma unroll 2
    for (int k = 0; k < 8; k += 4) {
        float v[4][VEC], av[4][VEC], s[4];
        int idx[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            idx[q] = tile * kRescaleUnits + (k + q) * kBlock + threadIdx.x;
            if (idx[q] < total) {
                const int c = idx[q] / U, u = idx[q] - c * U;
                ldv<T, VEC, kLdLastUse>(x + sbase + (size_t)idx[q] * VEC, v[q]);
                ldf<VEC>(ctx.a + (size_t)b * sh.S + (size_t)u * VEC, av[q]);
                s[q] = ldc(ctx.s + b * sh.C + c);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (idx[q] >= total) continue;
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const float gate = add ? (s[q] + av[q][e]) : (s[q] * av[q][e]);
                v[q][e] *= fmaf(k1, gate, k0);
            }
            stv<T, VEC, true>(out + sbase + (size_t)idx[q] * VEC, v[q]);
        }
    }
}

}  // namespace mga
