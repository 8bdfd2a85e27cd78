// spade_ops.cu -- the feature side of MaskSPADE, C ABI in include/mga_cbam.h:
//
//   mga_spade_forward    y = gamma * IN(x) + beta           /root/reference/mga_yolo/nn/modules/masked_spade.py:73,126,143
//                        (affine-free InstanceNorm2d: per (sample, channel) mean and biased variance over H*W, eps inside the root;
//                        gamma == NULL: plain normalisation, masked_spade.py:128-130)
//   mga_spade_backward   closed form of the same: d gamma = g * xhat, d beta = g (the caller's own tensor),
//                        dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)) with dxhat = g * gamma
//
// HBM-bound streaming kernels; one (sample, channel) row is always owned by ONE CTA (or warp), so both statistics are fixed-order sums
// with no atomics, x / gamma / beta are read from HBM once and y written once (forward 4 N elements of traffic; backward x, g, gamma in,
// dx, d gamma out = 5 N).  Three forms by row length (16-byte units = 4 fp32 / 8 16-bit pixels), measured in profiles/r2g_spade.md:
//   * long rows (P3: 80 x 80):  one CTA per row, the row staged in shared memory as fp32 (loads issued in batches and held raw; gamma /
//     beta announced to L2 by a bulk prefetch once the x row is in); rows above the staging budget are re-read through L2 instead;
//   * rows of <= two units per thread (P4: 40 x 40): one CTA per row, row AND gamma / beta in registers, one load batch, no staging;
//   * rows of <= 512 pixels (P5: 20 x 20): one WARP per row, eight rows per CTA, shuffles only.
// The mask branch that PRODUCES gamma / beta (3x3 convolutions 1 -> hidden -> C, masked_spade.py:78-84) is dense convolution work and
// stays with the caller's library.
#include <cstdint>
#include <cstdlib>

#include "common.cuh"
#include "mga_cbam.h"

namespace mga {

constexpr int kSpNT = 256;
constexpr size_t kSpStageBytes = 112 * 1024;  // two CTAs per SM keep their rows in shared memory

template <int VEC>
__device__ __forceinline__ void sp_lds(const float* p, float (&v)[VEC]) {  // shared-memory row, 128-bit reads
    if constexpr (VEC == 1) {
        v[0] = p[0];
    } else {
#pragma unroll
        for (int i = 0; i < VEC / 4; ++i) {
            const float4 t = reinterpret_cast<const float4*>(p)[i];
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
    }
}
template <typename TG, int VEC>
__device__ __forceinline__ void sp_ld_mod(const TG* __restrict__ p, float (&v)[VEC]) {
    if constexpr (VEC == 1) v[0] = to_f<TG>(p[0]);
    else if constexpr (sizeof(TG) == 4) ldf<VEC>(reinterpret_cast<const float*>(p), v);
    else ldv<TG, VEC, kLdStream>(p, v);
}
template <typename TG, int VEC>
__device__ __forceinline__ void sp_st_mod(TG* __restrict__ p, const float (&v)[VEC]) {
    if constexpr (VEC == 1) p[0] = from_f<TG>(v[0]);
    else if constexpr (sizeof(TG) == 4) stf<VEC>(reinterpret_cast<float*>(p), v);
    else stv<TG, VEC, true>(p, v);
}

// x: (R, S) rows of T; gamma / beta: (R, S) rows of TG or NULL; stats: (R, 2) = mean, rstd (saved for backward)
template <typename T, typename TG, int VEC, bool STAGED>
__global__ void __launch_bounds__(kSpNT) spade_fwd_kernel(const T* __restrict__ x, const TG* __restrict__ gamma, const TG* __restrict__ beta,
                                                          T* __restrict__ out, float* __restrict__ stats, int S, float eps, int pf) {
    extern __shared__ __align__(16) unsigned char sp_smem[];
    __shared__ float red[32];
    float* const rowf = reinterpret_cast<float*>(sp_smem);
    const int tid = threadIdx.x, nt = blockDim.x;
    const size_t base = (size_t)blockIdx.x * S;
    const T* const xr = x + base;
    const int nU = S / VEC;
    const float invS = 1.0f / (float)S;
    const bool mod = gamma != nullptr;
    // gamma / beta are needed only after the two statistics: bulk L2 prefetches of the two rows (copy engine, no registers, no
    // scoreboard) start their HBM fetch early, so the modulation pass reads them from L2 instead of paying a second serialized HBM
    // round trip per CTA.  pf = 1: at kernel entry; pf = 2: once the x row is in (shorter stay in L2); 0: none.
    auto prefetch_mod = [&]() {
        if constexpr (VEC > 1) {
            if (mod && tid == 0) {
                const size_t bytes = (size_t)S * sizeof(TG);
                const char* const gp = reinterpret_cast<const char*>(gamma + base);
                const char* const bp = reinterpret_cast<const char*>(beta + base);
                constexpr size_t kChunk = 32768;
                for (size_t o = 0; o < bytes; o += kChunk) {
                    const unsigned n = (unsigned)(bytes - o < kChunk ? bytes - o : kChunk);
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gp + o), "r"(n) : "memory");
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(bp + o), "r"(n) : "memory");
                }
            }
        }
    };
    if (pf == 1) prefetch_mod();

    float sum = 0.0f;
    constexpr int KU1 = (VEC == 1) ? 1 : 4;  // 16-byte loads in flight per thread
    for (int u0 = tid; u0 < nU; u0 += KU1 * nt) {
        RawV raw[KU1];
#pragma unroll
        for (int k = 0; k < KU1; ++k)
            if (u0 + k * nt < nU) raw[k] = ldraw<T, VEC, STAGED ? kLdStream : kLdDefault>(xr + (size_t)(u0 + k * nt) * VEC);
#pragma unroll
        for (int k = 0; k < KU1; ++k) {
            const int u = u0 + k * nt;
            if (u < nU) {
                float v[VEC];
                unpackv<T, VEC>(raw[k], v);
#pragma unroll
                for (int i = 0; i < VEC; ++i) sum += v[i];
                if constexpr (STAGED) stf<VEC>(rowf + (size_t)u * VEC, v);  // every thread re-reads only what it wrote: no barrier needed
            }
        }
    }
    if (pf == 2) prefetch_mod();
    const float mean = block_sum(sum, red) * invS;
    float sq = 0.0f;
    for (int u = tid; u < nU; u += nt) {
        float v[VEC];
        if constexpr (STAGED) {
            sp_lds<VEC>(rowf + (size_t)u * VEC, v);
        } else {
            ldv<T, VEC, kLdDefault>(xr + (size_t)u * VEC, v);
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float d = v[i] - mean;
            sq = fmaf(d, d, sq);
        }
    }
    const float var = block_sum(sq, red) * invS;
    const float rstd = 1.0f / sqrtf(var + eps);
    if (tid == 0) {
        stats[2 * (size_t)blockIdx.x] = mean;
        stats[2 * (size_t)blockIdx.x + 1] = rstd;
    }
    constexpr int KU3 = (VEC == 1) ? 1 : 2;
    for (int u0 = tid; u0 < nU; u0 += KU3 * nt) {
        float gm[KU3][VEC], bt[KU3][VEC], v[KU3][VEC];
#pragma unroll
        for (int k = 0; k < KU3; ++k) {
            const int u = u0 + k * nt;
            if (u < nU) {
                if (mod) {
                    sp_ld_mod<TG, VEC>(gamma + base + (size_t)u * VEC, gm[k]);
                    sp_ld_mod<TG, VEC>(beta + base + (size_t)u * VEC, bt[k]);
                }
                if constexpr (!STAGED) ldv<T, VEC, kLdStream>(xr + (size_t)u * VEC, v[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < KU3; ++k) {
            const int u = u0 + k * nt;
            if (u < nU) {
                float o[VEC];
                if constexpr (STAGED) sp_lds<VEC>(rowf + (size_t)u * VEC, v[k]);
                if (mod) {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) o[i] = fmaf(gm[k][i], (v[k][i] - mean) * rstd, bt[k][i]);
                } else {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) o[i] = (v[k][i] - mean) * rstd;
                }
                stv<T, VEC, true>(out + base + (size_t)u * VEC, o);
            }
        }
    }
}

template <typename T, typename TG, int VEC, bool STAGED>
__global__ void __launch_bounds__(kSpNT) spade_bwd_kernel(const T* __restrict__ x, const T* __restrict__ g, const TG* __restrict__ gamma,
                                                          const float* __restrict__ stats, T* __restrict__ dx, TG* __restrict__ dgamma, int S) {
    extern __shared__ __align__(16) unsigned char sp_smem[];
    __shared__ float red[32];
    float* const xh_s = reinterpret_cast<float*>(sp_smem);  // STAGED: xhat row, then dxhat row
    float* const dh_s = xh_s + S;
    const int tid = threadIdx.x, nt = blockDim.x;
    const size_t base = (size_t)blockIdx.x * S;
    const int nU = S / VEC;
    const float invS = 1.0f / (float)S;
    const float mean = __ldg(stats + 2 * (size_t)blockIdx.x), rstd = __ldg(stats + 2 * (size_t)blockIdx.x + 1);
    const bool mod = gamma != nullptr;

    float s1 = 0.0f, s2 = 0.0f;
    constexpr int KU = (VEC == 1) ? 1 : 2;  // units per batch: 6 loads of 16 bytes in flight per thread
    for (int u0 = tid; u0 < nU; u0 += KU * nt) {
        RawV rx[KU], rg[KU];
        float gm[KU][VEC];
#pragma unroll
        for (int k = 0; k < KU; ++k) {
            const int u = u0 + k * nt;
            if (u < nU) {
                const size_t o = base + (size_t)u * VEC;
                rx[k] = ldraw<T, VEC, STAGED ? kLdStream : kLdDefault>(x + o);
                rg[k] = ldraw<T, VEC, STAGED ? kLdStream : kLdDefault>(g + o);
                if (mod) sp_ld_mod<TG, VEC>(gamma + o, gm[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < KU; ++k) {
            const int u = u0 + k * nt;
            if (u < nU) {
                const size_t o = base + (size_t)u * VEC;
                float xv[VEC], gv[VEC], dh[VEC];
                unpackv<T, VEC>(rx[k], xv);
                unpackv<T, VEC>(rg[k], gv);
#pragma unroll
                for (int i = 0; i < VEC; ++i) xv[i] = (xv[i] - mean) * rstd;
                if (mod) {
                    float dg[VEC];
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        dh[i] = gv[i] * gm[k][i];
                        dg[i] = gv[i] * xv[i];
                    }
                    if (dgamma != nullptr) sp_st_mod<TG, VEC>(dgamma + o, dg);
                } else {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) dh[i] = gv[i];
                }
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    s1 += dh[i];
                    s2 = fmaf(dh[i], xv[i], s2);
                }
                if constexpr (STAGED) {
                    stf<VEC>(xh_s + (size_t)u * VEC, xv);
                    stf<VEC>(dh_s + (size_t)u * VEC, dh);
                }
            }
        }
    }
    const float m1 = block_sum(s1, red) * invS;
    const float m2 = block_sum(s2, red) * invS;
    for (int u = tid; u < nU; u += nt) {
        const size_t o = base + (size_t)u * VEC;
        float xv[VEC], dh[VEC], r[VEC];
        if constexpr (STAGED) {
            sp_lds<VEC>(xh_s + (size_t)u * VEC, xv);
            sp_lds<VEC>(dh_s + (size_t)u * VEC, dh);
        } else {
            float gv[VEC];
            ldv<T, VEC, kLdStream>(x + o, xv);
            ldv<T, VEC, kLdStream>(g + o, gv);
#pragma unroll
            for (int i = 0; i < VEC; ++i) xv[i] = (xv[i] - mean) * rstd;
            if (mod) {
                float gm[VEC];
                sp_ld_mod<TG, VEC>(gamma + o, gm);
#pragma unroll
                for (int i = 0; i < VEC; ++i) dh[i] = gv[i] * gm[i];
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) dh[i] = gv[i];
            }
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) r[i] = rstd * (dh[i] - m1 - xv[i] * m2);
        stv<T, VEC, true>(dx + o, r);
    }
}

// Short rows (P5: 20 x 20): ONE WARP per row, the row lives in registers (up to sp_kw<VEC>() 16-byte units per lane), the statistics are
// warp shuffles -- no shared memory, no block barrier, eight rows per CTA instead of one small CTA per row.
template <int VEC> __host__ __device__ constexpr int sp_kw() { return VEC == 8 ? 2 : 4; }  // 512 pixels per row either way; 16-bit rows keep the register count of fp32 ones
// ROWS = 1: one warp per row (P5), rows of <= 32 * sp_kw units; the CTA-wide form below (`short`) is the same idea for rows of up to
// two units per THREAD (P4): the whole row and its gamma / beta live in registers, every operand is requested in ONE batch at kernel
// entry (no second HBM round trip, no shared-memory staging), the statistics are shuffles (+ one block sum in the CTA-wide form).
template <typename T, typename TG, int VEC>
__global__ void __launch_bounds__(kSpNT) spade_fwd_warp_kernel(const T* __restrict__ x, const TG* __restrict__ gamma, const TG* __restrict__ beta,
                                                               T* __restrict__ out, float* __restrict__ stats, int R, int S, float eps) {
    constexpr int KW = sp_kw<VEC>();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= R) return;
    const size_t base = (size_t)row * S;
    const int nU = S / VEC;
    const float invS = 1.0f / (float)S;
    const bool mod = gamma != nullptr;
    RawV rx[KW];
    float gm[KW][VEC], bt[KW][VEC];
#pragma unroll
    for (int k = 0; k < KW; ++k) {
        const int u = lane + 32 * k;
        if (u < nU) {
            const size_t o = base + (size_t)u * VEC;
            rx[k] = ldraw<T, VEC, kLdStream>(x + o);
            if (mod) {
                sp_ld_mod<TG, VEC>(gamma + o, gm[k]);
                sp_ld_mod<TG, VEC>(beta + o, bt[k]);
            }
        }
    }
    float v[KW][VEC];
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < KW; ++k)
        if (lane + 32 * k < nU) {
            unpackv<T, VEC>(rx[k], v[k]);
#pragma unroll
            for (int i = 0; i < VEC; ++i) sum += v[k][i];
        }
    const float mean = warp_sum(sum) * invS;
    float sq = 0.0f;
#pragma unroll
    for (int k = 0; k < KW; ++k)
        if (lane + 32 * k < nU) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float d = v[k][i] - mean;
                sq = fmaf(d, d, sq);
            }
        }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) * invS + eps);
    if (lane == 0) {
        stats[2 * (size_t)row] = mean;
        stats[2 * (size_t)row + 1] = rstd;
    }
#pragma unroll
    for (int k = 0; k < KW; ++k) {
        const int u = lane + 32 * k;
        if (u < nU) {
            float r[VEC];
            if (mod) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) r[i] = fmaf(gm[k][i], (v[k][i] - mean) * rstd, bt[k][i]);
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) r[i] = (v[k][i] - mean) * rstd;
            }
            stv<T, VEC, true>(out + base + (size_t)u * VEC, r);
        }
    }
}

// one CTA (128 / 256 threads) per row of at most two units per thread; see above
template <typename T, typename TG, int VEC>
__global__ void __launch_bounds__(kSpNT) spade_fwd_short_kernel(const T* __restrict__ x, const TG* __restrict__ gamma, const TG* __restrict__ beta,
                                                                T* __restrict__ out, float* __restrict__ stats, int S, float eps) {
    __shared__ float red[32];
    constexpr int KS = 2;
    const int tid = threadIdx.x, nt = blockDim.x;
    const size_t base = (size_t)blockIdx.x * S;
    const int nU = S / VEC;
    const float invS = 1.0f / (float)S;
    const bool mod = gamma != nullptr;
    RawV rx[KS];
    float gm[KS][VEC], bt[KS][VEC];
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        const int u = tid + k * nt;
        if (u < nU) {
            const size_t o = base + (size_t)u * VEC;
            rx[k] = ldraw<T, VEC, kLdStream>(x + o);
            if (mod) {
                sp_ld_mod<TG, VEC>(gamma + o, gm[k]);
                sp_ld_mod<TG, VEC>(beta + o, bt[k]);
            }
        }
    }
    float v[KS][VEC];
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < KS; ++k)
        if (tid + k * nt < nU) {
            unpackv<T, VEC>(rx[k], v[k]);
#pragma unroll
            for (int i = 0; i < VEC; ++i) sum += v[k][i];
        }
    const float mean = block_sum(sum, red) * invS;
    float sq = 0.0f;
#pragma unroll
    for (int k = 0; k < KS; ++k)
        if (tid + k * nt < nU) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float d = v[k][i] - mean;
                sq = fmaf(d, d, sq);
            }
        }
    const float rstd = 1.0f / sqrtf(block_sum(sq, red) * invS + eps);
    if (tid == 0) {
        stats[2 * (size_t)blockIdx.x] = mean;
        stats[2 * (size_t)blockIdx.x + 1] = rstd;
    }
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        const int u = tid + k * nt;
        if (u < nU) {
            float r[VEC];
            if (mod) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) r[i] = fmaf(gm[k][i], (v[k][i] - mean) * rstd, bt[k][i]);
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) r[i] = (v[k][i] - mean) * rstd;
            }
            stv<T, VEC, true>(out + base + (size_t)u * VEC, r);
        }
    }
}

template <typename T, typename TG, int VEC>
__global__ void __launch_bounds__(kSpNT) spade_bwd_warp_kernel(const T* __restrict__ x, const T* __restrict__ g, const TG* __restrict__ gamma,
                                                               const float* __restrict__ stats, T* __restrict__ dx, TG* __restrict__ dgamma, int R, int S) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= R) return;
    const size_t base = (size_t)row * S;
    const int nU = S / VEC;
    const float invS = 1.0f / (float)S;
    const float mean = __ldg(stats + 2 * (size_t)row), rstd = __ldg(stats + 2 * (size_t)row + 1);
    const bool mod = gamma != nullptr;
    constexpr int kSpKW = sp_kw<VEC>();
    float xh[kSpKW][VEC], dh[kSpKW][VEC];
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int k = 0; k < kSpKW; ++k) {
        const int u = lane + 32 * k;
        if (u < nU) {
            const size_t o = base + (size_t)u * VEC;
            float gv[VEC];
            ldv<T, VEC, kLdStream>(x + o, xh[k]);
            ldv<T, VEC, kLdStream>(g + o, gv);
#pragma unroll
            for (int i = 0; i < VEC; ++i) xh[k][i] = (xh[k][i] - mean) * rstd;
            if (mod) {
                float gm[VEC], dg[VEC];
                sp_ld_mod<TG, VEC>(gamma + o, gm);
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    dh[k][i] = gv[i] * gm[i];
                    dg[i] = gv[i] * xh[k][i];
                }
                if (dgamma != nullptr) sp_st_mod<TG, VEC>(dgamma + o, dg);
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) dh[k][i] = gv[i];
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                s1 += dh[k][i];
                s2 = fmaf(dh[k][i], xh[k][i], s2);
            }
        }
    }
    const float m1 = warp_sum(s1) * invS, m2 = warp_sum(s2) * invS;
#pragma unroll
    for (int k = 0; k < kSpKW; ++k) {
        const int u = lane + 32 * k;
        if (u < nU) {
            float r[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) r[i] = rstd * (dh[k][i] - m1 - xh[k][i] * m2);
            stv<T, VEC, true>(dx + base + (size_t)u * VEC, r);
        }
    }
}

// backward of a row of at most two units per thread: xhat and dxhat stay in registers, no shared-memory staging
template <typename T, typename TG, int VEC>
__global__ void __launch_bounds__(kSpNT) spade_bwd_short_kernel(const T* __restrict__ x, const T* __restrict__ g, const TG* __restrict__ gamma,
                                                                const float* __restrict__ stats, T* __restrict__ dx, TG* __restrict__ dgamma, int S) {
    __shared__ float red[32];
    constexpr int KS = 2;
    const int tid = threadIdx.x, nt = blockDim.x;
    const size_t base = (size_t)blockIdx.x * S;
    const int nU = S / VEC;
    const float invS = 1.0f / (float)S;
    const float mean = __ldg(stats + 2 * (size_t)blockIdx.x), rstd = __ldg(stats + 2 * (size_t)blockIdx.x + 1);
    const bool mod = gamma != nullptr;
    RawV rx[KS], rg[KS];
    float gm[KS][VEC];
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        const int u = tid + k * nt;
        if (u < nU) {
            const size_t o = base + (size_t)u * VEC;
            rx[k] = ldraw<T, VEC, kLdStream>(x + o);
            rg[k] = ldraw<T, VEC, kLdStream>(g + o);
            if (mod) sp_ld_mod<TG, VEC>(gamma + o, gm[k]);
        }
    }
    float xh[KS][VEC], dh[KS][VEC];
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        const int u = tid + k * nt;
        if (u < nU) {
            float gv[VEC];
            unpackv<T, VEC>(rx[k], xh[k]);
            unpackv<T, VEC>(rg[k], gv);
#pragma unroll
            for (int i = 0; i < VEC; ++i) xh[k][i] = (xh[k][i] - mean) * rstd;
            if (mod) {
                float dg[VEC];
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    dh[k][i] = gv[i] * gm[k][i];
                    dg[i] = gv[i] * xh[k][i];
                }
                if (dgamma != nullptr) sp_st_mod<TG, VEC>(dgamma + base + (size_t)u * VEC, dg);
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) dh[k][i] = gv[i];
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                s1 += dh[k][i];
                s2 = fmaf(dh[k][i], xh[k][i], s2);
            }
        }
    }
    const float m1 = block_sum(s1, red) * invS;
    const float m2 = block_sum(s2, red) * invS;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        const int u = tid + k * nt;
        if (u < nU) {
            float r[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) r[i] = rstd * (dh[k][i] - m1 - xh[k][i] * m2);
            stv<T, VEC, true>(dx + base + (size_t)u * VEC, r);
        }
    }
}

// tuning builds read MGA_SP_PREFETCH (0 none, 1 at kernel entry, 2 after the first pass); the product library has no knobs
static int sp_prefetch_mode() {
#ifdef MGA_TUNING
    static const int m = [] { const char* e = getenv("MGA_SP_PREFETCH"); return (e && *e) ? atoi(e) : 2; }();
    return m;
#else
    return 2;  // measured on B200 (profiles/r2g_spade.md): 74 us at P3 fp32 against 79 (none) and 84 (at entry: the rows are evicted again before use)
#endif
}
// tuning builds: MGA_SP_SHORT=0 sends rows of <= two units per thread through the general (staged) kernels
static bool sp_short_rows() {
#ifdef MGA_TUNING
    static const int m = [] { const char* e = getenv("MGA_SP_SHORT"); return (e && *e) ? atoi(e) : 1; }();
    return m != 0;
#else
    return true;
#endif
}
// rows up to this many bytes of fp32 staging live in shared memory (tuning builds: MGA_SP_STAGE_KB_F / _B)
static size_t sp_stage_limit(bool bwd) {
#ifdef MGA_TUNING
    static const long f = [] { const char* e = getenv("MGA_SP_STAGE_KB_F"); return (e && *e) ? atol(e) * 1024 : (long)kSpStageBytes; }();
    static const long b = [] { const char* e = getenv("MGA_SP_STAGE_KB_B"); return (e && *e) ? atol(e) * 1024 : (long)kSpStageBytes; }();
    return (size_t)(bwd ? b : f);
#else
    (void)bwd;
    return kSpStageBytes;
#endif
}
// small rows (P5: 20 x 20) take small CTAs so that the SM still holds many rows
static int sp_threads(int units) { return units <= 128 ? 64 : (units <= 512 ? 128 : kSpNT); }
static bool sp_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename Kern>
static int sp_prepare(Kern kern, size_t smem) {
    // per-device function state: set on every launch that needs it (cheap), never cached across devices
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return fail(MGA_ERR_CUDA, "cudaFuncSetAttribute(%zu bytes of shared memory) failed", smem);
    }
    return MGA_OK;
}

template <typename T, typename TG>
static int spade_fwd_launch(const void* x, const void* gamma, const void* beta, void* out, float* stats, int R, int S, float eps, cudaStream_t st) {
    constexpr int V = 16 / (int)sizeof(T);
    const bool vec = S % V == 0 && sp_aligned16(x) && sp_aligned16(out) && sp_aligned16(gamma) && sp_aligned16(beta);
    const size_t stage = (size_t)S * sizeof(float);
    const bool staged = stage <= sp_stage_limit(false);
    const size_t smem = staged ? stage : 0;
#define MGA_SP_F(VEC, STG)                                                                                                              \
    do {                                                                                                                                \
        auto kern = spade_fwd_kernel<T, TG, VEC, STG>;                                                                                  \
        if (int rc = sp_prepare(kern, smem)) return rc;                                                                                 \
        kern<<<R, sp_threads(S / VEC), smem, st>>>(static_cast<const T*>(x), static_cast<const TG*>(gamma), static_cast<const TG*>(beta), static_cast<T*>(out), \
                                     stats, S, eps, sp_prefetch_mode());                                                                \
    } while (0)
    if (vec && S / V <= 32 * sp_kw<V>()) {  // short rows: one warp per row, eight rows per CTA
        spade_fwd_warp_kernel<T, TG, V><<<(R + kSpNT / 32 - 1) / (kSpNT / 32), kSpNT, 0, st>>>(
            static_cast<const T*>(x), static_cast<const TG*>(gamma), static_cast<const TG*>(beta), static_cast<T*>(out), stats, R, S, eps);
        return MGA_OK;
    }
    if (vec && S / V <= 2 * kSpNT && sp_short_rows()) {  // rows of at most two units per thread: everything in registers, one load batch
        spade_fwd_short_kernel<T, TG, V><<<R, S / V <= 256 ? 128 : kSpNT, 0, st>>>(
            static_cast<const T*>(x), static_cast<const TG*>(gamma), static_cast<const TG*>(beta), static_cast<T*>(out), stats, S, eps);
        return MGA_OK;
    }
    if (vec) { if (staged) MGA_SP_F(V, true); else MGA_SP_F(V, false); }
    else     { if (staged) MGA_SP_F(1, true); else MGA_SP_F(1, false); }
#undef MGA_SP_F
    return MGA_OK;
}

template <typename T, typename TG>
static int spade_bwd_launch(const void* x, const void* g, const void* gamma, const float* stats, void* dx, void* dgamma, int R, int S, cudaStream_t st) {
    constexpr int V = 16 / (int)sizeof(T);
    const bool vec = S % V == 0 && sp_aligned16(x) && sp_aligned16(g) && sp_aligned16(dx) && sp_aligned16(gamma) && sp_aligned16(dgamma);
    const size_t stage = 2 * (size_t)S * sizeof(float);
    const bool staged = stage <= sp_stage_limit(true);
    const size_t smem = staged ? stage : 0;
#define MGA_SP_B(VEC, STG)                                                                                                              \
    do {                                                                                                                                \
        auto kern = spade_bwd_kernel<T, TG, VEC, STG>;                                                                                  \
        if (int rc = sp_prepare(kern, smem)) return rc;                                                                                 \
        kern<<<R, sp_threads(S / VEC), smem, st>>>(static_cast<const T*>(x), static_cast<const T*>(g), static_cast<const TG*>(gamma), stats, static_cast<T*>(dx), \
                                     static_cast<TG*>(dgamma), S);                                                                      \
    } while (0)
    if (vec && S / V <= 32 * sp_kw<V>()) {
        spade_bwd_warp_kernel<T, TG, V><<<(R + kSpNT / 32 - 1) / (kSpNT / 32), kSpNT, 0, st>>>(
            static_cast<const T*>(x), static_cast<const T*>(g), static_cast<const TG*>(gamma), stats, static_cast<T*>(dx), static_cast<TG*>(dgamma), R, S);
        return MGA_OK;
    }
    if (vec && S / V <= 2 * kSpNT && sp_short_rows()) {
        spade_bwd_short_kernel<T, TG, V><<<R, S / V <= 256 ? 128 : kSpNT, 0, st>>>(
            static_cast<const T*>(x), static_cast<const T*>(g), static_cast<const TG*>(gamma), stats, static_cast<T*>(dx), static_cast<TG*>(dgamma), S);
        return MGA_OK;
    }
    if (vec) { if (staged) MGA_SP_B(V, true); else MGA_SP_B(V, false); }
    else     { if (staged) MGA_SP_B(1, true); else MGA_SP_B(1, false); }
#undef MGA_SP_B
    return MGA_OK;
}

static int sp_launch_ok(const char* name) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(MGA_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e));
    return MGA_OK;
}

}  // namespace mga

using namespace mga;

extern "C" {

int mga_spade_forward(const void* x, const void* gamma, const void* beta, void* out, float* stats, int32_t B, int32_t C, int32_t H, int32_t W, float eps,
                      int32_t dtype, int32_t mod_dtype, void* stream) {
    if (!x || !out || !stats) return fail(MGA_ERR_ARG, "mga_spade_forward: null pointer argument");
    if ((gamma == nullptr) != (beta == nullptr)) return fail(MGA_ERR_ARG, "mga_spade_forward: gamma and beta come together");
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || (long long)B * C > 0x7fffffffLL || (long long)H * W > (1 << 28)) return fail(MGA_ERR_ARG, "mga_spade_forward: bad shape");
    if (!(eps >= 0.0f)) return fail(MGA_ERR_ARG, "mga_spade_forward: eps must be >= 0");
    if (gamma && mod_dtype != dtype && mod_dtype != MGA_F32) return fail(MGA_ERR_UNSUPPORTED, "mga_spade_forward: gamma / beta must have the feature dtype or float32");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int R = B * C, S = H * W;
    const bool modf32 = gamma && mod_dtype == MGA_F32;
    int rc = MGA_OK;
    switch (dtype) {
        case MGA_F32: rc = spade_fwd_launch<float, float>(x, gamma, beta, out, stats, R, S, eps, st); break;
        case MGA_BF16: rc = modf32 ? spade_fwd_launch<__nv_bfloat16, float>(x, gamma, beta, out, stats, R, S, eps, st)
                                   : spade_fwd_launch<__nv_bfloat16, __nv_bfloat16>(x, gamma, beta, out, stats, R, S, eps, st); break;
        case MGA_F16: rc = modf32 ? spade_fwd_launch<__half, float>(x, gamma, beta, out, stats, R, S, eps, st)
                                  : spade_fwd_launch<__half, __half>(x, gamma, beta, out, stats, R, S, eps, st); break;
        default: return fail(MGA_ERR_ARG, "mga_spade_forward: bad feature dtype %d", dtype);
    }
    return rc ? rc : sp_launch_ok("mga_spade_forward");
}

int mga_spade_backward(const void* x, const void* grad_out, const void* gamma, const float* stats, void* grad_x, void* grad_gamma, int32_t B, int32_t C,
                       int32_t H, int32_t W, int32_t dtype, int32_t mod_dtype, void* stream) {
    if (!x || !grad_out || !stats || !grad_x) return fail(MGA_ERR_ARG, "mga_spade_backward: null pointer argument");
    if (!gamma && grad_gamma) return fail(MGA_ERR_ARG, "mga_spade_backward: grad_gamma without gamma");
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || (long long)B * C > 0x7fffffffLL || (long long)H * W > (1 << 28)) return fail(MGA_ERR_ARG, "mga_spade_backward: bad shape");
    if (gamma && mod_dtype != dtype && mod_dtype != MGA_F32) return fail(MGA_ERR_UNSUPPORTED, "mga_spade_backward: gamma must have the feature dtype or float32");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int R = B * C, S = H * W;
    const bool modf32 = gamma && mod_dtype == MGA_F32;
    int rc = MGA_OK;
    switch (dtype) {
        case MGA_F32: rc = spade_bwd_launch<float, float>(x, grad_out, gamma, stats, grad_x, grad_gamma, R, S, st); break;
        case MGA_BF16: rc = modf32 ? spade_bwd_launch<__nv_bfloat16, float>(x, grad_out, gamma, stats, grad_x, grad_gamma, R, S, st)
                                   : spade_bwd_launch<__nv_bfloat16, __nv_bfloat16>(x, grad_out, gamma, stats, grad_x, grad_gamma, R, S, st); break;
        case MGA_F16: rc = modf32 ? spade_bwd_launch<__half, float>(x, grad_out, gamma, stats, grad_x, grad_gamma, R, S, st)
                                  : spade_bwd_launch<__half, __half>(x, grad_out, gamma, stats, grad_x, grad_gamma, R, S, st); break;
        default: return fail(MGA_ERR_ARG, "mga_spade_backward: bad feature dtype %d", dtype);
    }
    return rc ? rc : sp_launch_ok("mga_spade_backward");
}

}  // extern "C"
