// cbam_concat.cuh -- fused forward of the build-side `sam_cam_fusion = concat` mode on the 5th-generation tensor cores.
//
//   R   = Conv1x1([x*s ; x*a'], W (C x 2C), bias)            (no reference source, parity unpinned)
//       = (Wa diag(s_b)) X_b + a'_b (.) (Wb X_b) + bias        per sample b, X_b = (C x S) feature map, Wa = W[:, :C], Wb = W[:, C:]
//   out = k0 * x + k1 * R                                      (mga_pyramid_fusion add: k0 = 1 - alpha, k1 = alpha; multiply: k0 = 0, k1 = alpha)
//
// The 2C-channel concat tensor is never formed: the channel gate is folded into per-sample weights (a small pre-kernel writes
// bf16(Wa diag(s_b)) for every sample and bf16(Wb) once), the spatial gate is applied in the epilogue.  Per CTA: a 128 (output channels)
// x 128 (pixels) tile of one sample, two fp32 accumulators (Wa' X and Wb X) in TENSOR MEMORY (256 columns), K = C in steps of 64:
//   warp 0 (one lane)  TMA producer: weights (K-major, 128-byte swizzle) and the feature tile (pixels contiguous = MN-major B operand,
//                      128-byte swizzle, two 64-pixel boxes) into a 4-stage shared-memory ring, mbarrier expect_tx
//   warp 1 (one lane)  tcgen05.mma.cta_group::1.kind::f16 (M 128, N 128, K 16) x 2 accumulators, tcgen05.commit frees the stage
//   warps 2-5          epilogue: tcgen05.ld 32x32b.x16 of both accumulators, a' and bias applied, packed 16-bit stores
// x is read once (the second M tile of a pixel tile hits L2), out is written once: 2N*e of HBM traffic for 4*C*N flops.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace mga {

constexpr int kCcBM = 128, kCcBN = 128, kCcBK = 64, kCcStages = 4;
#ifndef MGA_CC_EPI_WARPS
#define MGA_CC_EPI_WARPS 8
#endif
constexpr int kCcEpiWarps = MGA_CC_EPI_WARPS;        // kCcEpiSplit per TMEM lane quadrant: each takes 1 / kCcEpiSplit of the tile's pixel columns
constexpr int kCcEpiSplit = kCcEpiWarps / 4;         // 2 (64 columns per warp) or 4 (32 columns per warp)
constexpr int kCcEpiChunks = 8 / kCcEpiSplit;        // 16-column accumulator chunks per warp
constexpr int kCcThreads = 64 + 32 * kCcEpiWarps;
constexpr int kCcStageBytes = (2 * kCcBM * kCcBK + kCcBK * kCcBN) * 2;  // two weight tiles + the feature tile, 16-bit elements
constexpr int kCcSmemBytes = kCcStages * kCcStageBytes + 1024 /* alignment slack */ + 256 /* barriers */ + 2 * kCcBN * 4 /* spatial gate of the tile */;

struct ConcatMaps { CUtensorMap wa, wb, x; };

__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading / stride byte offsets (16-byte units),
// version 1 (Blackwell), 128-byte swizzle
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // version
    d |= (uint64_t)2 << 61;   // SWIZZLE_128B
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16: fp32 accumulate, A K-major, B MN-major
__device__ __forceinline__ uint32_t umma_instr_desc(int fmt /* 0 f16, 1 bf16 */, int M, int N) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}

// per-sample weights with the channel gate folded in: wa[b][co][ci] = W[co][ci] * s[b][ci], wb[co][ci] = W[co][C + ci]
// grid (C / 8 rows-of-8, B + 1): a warp owns one weight row, lanes walk its input channels (coalesced, no divisions)
template <typename T>
__global__ void __launch_bounds__(kBlock) concat_fold_kernel(const float* __restrict__ w, const float* __restrict__ s, T* __restrict__ wa, T* __restrict__ wb,
                                                             int B, int C) {
    const int lane = threadIdx.x & 31, co = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5), b = blockIdx.y;
    if (co >= C) return;
    const float* wr = w + (size_t)co * 2 * C;
    if (b < B) {
        T* dst = wa + ((size_t)b * C + co) * C;
        const float* sb = s + (size_t)b * C;
        for (int ci = lane; ci < C; ci += 32) dst[ci] = from_f<T>(__ldg(wr + ci) * __ldg(sb + ci));
    } else {
        T* dst = wb + (size_t)co * C;
        for (int ci = lane; ci < C; ci += 32) dst[ci] = from_f<T>(__ldg(wr + C + ci));
    }
}

// epilogue of one 128 x 128 tile by the 8 epilogue warps (warp indices 2..9): warp w reads TMEM lanes 32 (w % 4) .. (row co = m0 + 32 (w % 4)
// + lane) and the pixel columns [64 h, 64 h + 64) with h = (w - 2) / 4, in chunks of 16 columns of both accumulators (y1 at tacc, y2 at
// tacc + 128): out = k0 x + k1 (y1 + a y2 + bias).  The spatial gate of the tile's 128 pixels sits in shared memory (a_sm).
template <typename T>
__device__ __forceinline__ void concat_epilogue_tile(uint32_t tacc, int warp, int lane, const T* __restrict__ x, const float* a_sm,
                                                     const float* __restrict__ bias, T* __restrict__ out, int b, int m0, int p0, int C, int S,
                                                     float k0, float k1) {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int co = m0 + q * 32 + lane;
    const float bco = __ldg(bias + co);
    const uint32_t trow = tacc + ((uint32_t)(q * 32) << 16);
    const size_t rowoff = ((size_t)b * C + co) * S;
#pragma unroll 1
    for (int c16 = half * kCcEpiChunks; c16 < half * kCcEpiChunks + kCcEpiChunks; ++c16) {
        const int p = p0 + c16 * 16;
        uint32_t y1[16], y2[16];
        tmem_ld16(trow + c16 * 16, y1);
        tmem_ld16(trow + kCcBN + c16 * 16, y2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (p >= S) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ph8 = p + h * 8;
            if (ph8 >= S) break;  // (S % 8 == 0: whole 16-byte pieces)
            const float4 a0 = *reinterpret_cast<const float4*>(a_sm + c16 * 16 + h * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(a_sm + c16 * 16 + h * 8 + 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = k1 * (__uint_as_float(y1[h * 8 + i]) + fmaf(av[i], __uint_as_float(y2[h * 8 + i]), bco));
            if (k0 != 0.0f) {
                float xv[8];
                ldv<T, 8, kLdStream>(x + rowoff + ph8, xv);
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = fmaf(k0, xv[i], r[i]);
            }
            stv<T, 8, true>(out + rowoff + ph8, r);
        }
    }
}
// ---- backward (MODE 1 of the two GEMM kernels below): the same two-accumulator GEMM on the upstream gradient,
//        U = Wa^T g,  V = Wb^T g   (weights transposed by concat_transpose_kernel, shared by the batch; feature operand = g),
// and everything else of the closed-form backward in the epilogue, on the fp32 accumulators (U, V never reach memory):
//   dx   = alpha (s_c U + a_p V) + k0 g          ga = g a_p  (left operand of the dWb GEMM)
//   ds_c = alpha sum_p x U      dbias_c = alpha sum_p g      da_p = alpha sum_c x V      dalpha = sum x (s U + a V) + sum_c bias_c sum_p g - [add] sum g x
// A lane owns one channel row of the tile: the per-channel sums are lane-private over the warp's 64 pixel columns (written per (pixel
// tile, half) -> ds_part / dbias_part (B, 2 nTilesN, C)); the per-pixel sum over channels is a warp reduction over the 32 rows
// (-> da_part (B, C / 32, S)); dalpha per warp (-> dalpha_part (B, nTilesN, C / 128 * 8)).  All partial buffers are written, never
// accumulated: deterministic, no atomics.
struct ConcatBwd {
    const void* x;       // saved features (B, C, S)
    const float* s;      // channel gate (B, C)
    void* ga;            // out: g * a' (B, C, S)
    float* ds_part;
    float* db_part;
    float* da_part;
    float* dal_part;
};

// this thread's pieces of x and grad_out for its row and the 64 pixel columns of its half of the tile: issued BEFORE the wait for the
// accumulators, so their latency hides behind the tile's MMAs (the first version loaded them chunk by chunk after the wait: 1284 us at
// cfg4-P3 against 912 us, every chunk a serial round trip to L2 / HBM)
struct ConcatBwdRegs { uint4 xr[2 * kCcEpiChunks], gr[2 * kCcEpiChunks]; };
template <typename T>
__device__ __forceinline__ void concat_epilogue_bwd_prefetch(ConcatBwdRegs& r, int warp, int lane, const T* __restrict__ g, const ConcatBwd& bw, int b,
                                                             int m0, int p0, int C, int S) {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const size_t rowoff = ((size_t)b * C + m0 + q * 32 + lane) * S;
    const T* __restrict__ xf = static_cast<const T*>(bw.x);
#pragma unroll
    for (int k = 0; k < 2 * kCcEpiChunks; ++k) {
        const int p = p0 + half * (16 * kCcEpiChunks) + k * 8;
        if (p < S) {  // (S % 8 == 0: whole 16-byte pieces; warp-uniform)
            r.xr[k] = __ldg(reinterpret_cast<const uint4*>(xf + rowoff + p));
            r.gr[k] = __ldg(reinterpret_cast<const uint4*>(g + rowoff + p));
        }
    }
}
template <typename T>
__device__ __forceinline__ void unpack8(const uint4& t, float (&v)[8]) {
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

template <typename T>
__device__ __forceinline__ void concat_epilogue_bwd_tile(uint32_t tacc, const ConcatBwdRegs& r, int warp, int lane, const float* a_sm,
                                                         const float* __restrict__ bias, T* __restrict__ dx, const ConcatBwd& bw, int b, int m0,
                                                         int p0, int C, int S, float k0, float alpha, int pyramid_multiply) {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int c = m0 + q * 32 + lane;
    const float sc = __ldg(bw.s + (size_t)b * C + c);
    const uint32_t trow = tacc + ((uint32_t)(q * 32) << 16);
    const size_t rowoff = ((size_t)b * C + c) * S;
    T* __restrict__ gaf = static_cast<T*>(bw.ga);
    float* __restrict__ dap = bw.da_part + ((size_t)b * (C / 32) + (m0 / 32 + q)) * S;
    const int nTilesN = (S + kCcBN - 1) / kCcBN;
    float su = 0.0f, sg = 0.0f, dal = 0.0f;
#pragma unroll
    for (int cc = 0; cc < kCcEpiChunks; ++cc) {
        const int c16 = half * kCcEpiChunks + cc;
        const int p = p0 + c16 * 16;
        uint32_t y1[16], y2[16];
        tmem_ld16(trow + c16 * 16, y1);
        tmem_ld16(trow + kCcBN + c16 * 16, y2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (p >= S) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ph8 = p + h * 8;
            if (ph8 >= S) break;
            const float4 a0 = *reinterpret_cast<const float4*>(a_sm + c16 * 16 + h * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(a_sm + c16 * 16 + h * 8 + 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float xv[8], gv[8], ov[8], gav[8], va[8];
            unpack8<T>(r.xr[cc * 2 + h], xv);
            unpack8<T>(r.gr[cc * 2 + h], gv);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float U = __uint_as_float(y1[h * 8 + i]), V = __uint_as_float(y2[h * 8 + i]);
                const float t = fmaf(sc, U, av[i] * V);
                ov[i] = fmaf(alpha, t, k0 * gv[i]);
                gav[i] = gv[i] * av[i];
                su = fmaf(xv[i], U, su);
                sg += gv[i];
                va[i] = xv[i] * V;
                dal = fmaf(xv[i], t, dal);
                if (!pyramid_multiply) dal = fmaf(-gv[i], xv[i], dal);
            }
            stv<T, 8, true>(dx + rowoff + ph8, ov);
            stv<T, 8, true>(gaf + rowoff + ph8, gav);
            // per-pixel sum over the warp's 32 channel rows: transposing butterfly (9 shuffles for 8 columns), lane l ends with column (l >> 2) & 7
            {
                const bool hi16 = lane & 16, hi8 = lane & 8, hi4 = lane & 4;
                float w4[4], w2[2], w1;
#pragma unroll
                for (int i = 0; i < 4; ++i) {  // lanes with bit 4 keep columns 4..7, the others 0..3
                    const float send = hi16 ? va[i] : va[4 + i];
                    const float got = __shfl_xor_sync(0xffffffffu, send, 16);
                    w4[i] = (hi16 ? va[4 + i] : va[i]) + got;
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {  // bit 3: columns +2, +3 of the kept four
                    const float send = hi8 ? w4[i] : w4[2 + i];
                    const float got = __shfl_xor_sync(0xffffffffu, send, 8);
                    w2[i] = (hi8 ? w4[2 + i] : w4[i]) + got;
                }
                {
                    const float send = hi4 ? w2[0] : w2[1];
                    const float got = __shfl_xor_sync(0xffffffffu, send, 4);
                    w1 = (hi4 ? w2[1] : w2[0]) + got;
                }
                w1 += __shfl_xor_sync(0xffffffffu, w1, 2);
                w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
                if ((lane & 3) == 0) dap[ph8 + (hi16 ? 4 : 0) + (hi8 ? 2 : 0) + (hi4 ? 1 : 0)] = alpha * w1;
            }
        }
    }
    const size_t pi = ((size_t)b * (kCcEpiSplit * nTilesN) + (size_t)(p0 / kCcBN) * kCcEpiSplit + half) * C + c;
    bw.ds_part[pi] = alpha * su;
    bw.db_part[pi] = alpha * sg;
    dal = fmaf(__ldg(bias + c), sg, dal);
    dal = warp_sum(dal);
    if (lane == 0) bw.dal_part[((size_t)b * nTilesN + p0 / kCcBN) * (size_t)(C / kCcBM * kCcEpiWarps) + (m0 / kCcBM) * kCcEpiWarps + (warp - 2)] = dal;
}

// W (C, 2C) fp32 -> wat[ci][co] = W[co][ci], wbt[ci][co] = W[co][C + ci] in the feature dtype (A operands of the backward GEMMs, K = co contiguous)
template <typename T>
__global__ void __launch_bounds__(kBlock) concat_transpose_kernel(const float* __restrict__ w, T* __restrict__ wat, T* __restrict__ wbt, int C) {
    const int e = blockIdx.x * kBlock + threadIdx.x;
    if (e >= C * C) return;
    const int ci = e / C, co = e - ci * C;
    wat[e] = from_f<T>(__ldg(w + (size_t)co * 2 * C + ci));
    wbt[e] = from_f<T>(__ldg(w + (size_t)co * 2 * C + C + ci));
}

// the spatial gate of pixels [p0, p0 + 128) of sample b into shared memory (epilogue threads only), then a barrier among them
__device__ __forceinline__ void concat_stage_gate(float* a_sm, const float* __restrict__ agate, int b, int p0, int S) {
    const int i = (int)threadIdx.x - 64;
    if (i < kCcBN) a_sm[i] = (p0 + i < S) ? __ldg(agate + (size_t)b * S + p0 + i) : 0.0f;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kCcEpiWarps) : "memory");
}

template <typename T, int MODE = 0 /* 0: forward, 1: backward (x := grad_out, out := grad_x, weights transposed and batch-shared) */>
__global__ void __launch_bounds__(kCcThreads, 1) concat_fwd_kernel(const __grid_constant__ ConcatMaps maps, const T* __restrict__ x,
                                                                   const float* __restrict__ agate, const float* __restrict__ bias,
                                                                   const float* __restrict__ beta, T* __restrict__ out, int C, int S, int pyramid_multiply,
                                                                   const ConcatBwd bw) {
    extern __shared__ __align__(1024) unsigned char ccsm_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ccsm_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kCcStages * kCcStageBytes);
    uint64_t* empty = full + kCcStages;
    uint64_t* tmem_full = empty + kCcStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    float* a_sm = reinterpret_cast<float*>(smem + kCcStages * kCcStageBytes + 256);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p0 = blockIdx.x * kCcBN, m0 = blockIdx.y * kCcBM, b = blockIdx.z;
    const int nkb = C / kCcBK;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kCcStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        fence_mbar_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.wa) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.wb) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x) : "memory");
    }
    if (warp == 1) {  // tensor memory: two 128-column fp32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * kCcBN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % kCcStages, ph = (kb / kCcStages) & 1;
                mbar_wait(&empty[st], ph ^ 1);
                unsigned char* sa = smem + st * kCcStageBytes;
                mbar_expect_tx(&full[st], kCcStageBytes);
                tma_load_2d(sa, &maps.wa, kb * kCcBK, MODE == 0 ? b * C + m0 : m0, &full[st]);
                tma_load_2d(sa + kCcBM * kCcBK * 2, &maps.wb, kb * kCcBK, m0, &full[st]);
                unsigned char* sb = sa + 2 * kCcBM * kCcBK * 2;
                tma_load_2d(sb, &maps.x, p0, b * C + kb * kCcBK, &full[st]);
                tma_load_2d(sb + kCcBK * 64 * 2, &maps.x, p0 + 64, b * C + kb * kCcBK, &full[st]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer
            const uint32_t idesc = umma_instr_desc(std::is_same<T, __nv_bfloat16>::value ? 1 : 0, kCcBM, kCcBN);
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % kCcStages, ph = (kb / kCcStages) & 1;
                mbar_wait(&full[st], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + st * kCcStageBytes);
                const uint32_t sa2 = sa + kCcBM * kCcBK * 2, sb = sa + 2 * kCcBM * kCcBK * 2;
#pragma unroll
                for (int k = 0; k < kCcBK / 16; ++k) {
                    // A (weights): K-major, rows of 128 B, 8-row groups 1024 B apart; 16 K elements = 32 B inside the swizzle atom
                    const uint64_t a1 = umma_smem_desc(sa + k * 32, 16, 1024);
                    const uint64_t a2 = umma_smem_desc(sa2 + k * 32, 16, 1024);
                    // B (features): MN-major, 64-pixel atoms 8 KB apart (the two boxes), 8-channel groups 1024 B apart; 16 K = 2 groups
                    const uint64_t bd = umma_smem_desc(sb + k * 2048, kCcBK * 64 * 2, 1024);
                    const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
                    umma_f16(tmem_base, a1, bd, idesc, acc);
                    umma_f16(tmem_base + kCcBN, a2, bd, idesc, acc);
                }
                umma_commit(&empty[st]);  // (implies tcgen05.fence::before_thread_sync) the stage is free once these MMAs have read it
            }
            umma_commit(tmem_full);
        }
    } else {
        // ---- epilogue
        const float alpha = softplusf_acc(__ldg(beta));
        concat_stage_gate(a_sm, agate, b, p0, S);
        [[maybe_unused]] ConcatBwdRegs pre;
        if constexpr (MODE == 1) concat_epilogue_bwd_prefetch<T>(pre, warp, lane, x, bw, b, m0, p0, C, S);
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if constexpr (MODE == 0)
            concat_epilogue_tile<T>(tmem_base, warp, lane, x, a_sm, bias, out, b, m0, p0, C, S, pyramid_multiply ? 0.0f : 1.0f - alpha, alpha);
        else
            concat_epilogue_bwd_tile<T>(tmem_base, pre, warp, lane, a_sm, bias, out, bw, b, m0, p0, C, S, pyramid_multiply ? 0.0f : 1.0f - alpha, alpha,
                                        pyramid_multiply);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * kCcBN) : "memory");
}


// ---------------------------------------------------------------- C <= 256: weights RESIDENT in shared memory, persistent CTAs
// A work item = (sample b, 128 output channels, a run of pixel tiles).  The two weight tiles of the item (2 x 128 x C 16-bit = 128 KB at
// C = 256) are loaded once and stay; only the feature tiles stream through the 4-stage ring (one 64-channel k-block per stage), and
// the two accumulators are double-buffered in tensor memory (512 columns), so the epilogue of tile t overlaps the MMAs of tile t + 1.
// Weight traffic per output element drops from 2 * 2C * e / 128 bytes (every tile reloads them) to ~0: HBM/L2 see x and out only.
constexpr int kCcResStageBytes = kCcBK * kCcBN * 2;  // one feature k-block
__host__ __device__ inline int concat_res_smem(int C) { return (C / kCcBK) * 2 * kCcBM * kCcBK * 2 + kCcStages * kCcResStageBytes + 1024 + 256 + 2 * kCcBN * 4; }

template <typename T, int MODE = 0>
__global__ void __launch_bounds__(kCcThreads, 1) concat_fwd_res_kernel(const __grid_constant__ ConcatMaps maps, const T* __restrict__ x,
                                                                       const float* __restrict__ agate, const float* __restrict__ bias,
                                                                       const float* __restrict__ beta, T* __restrict__ out, int B, int C, int S,
                                                                       int tiles_per_item, int pyramid_multiply, const ConcatBwd bw) {
    extern __shared__ __align__(1024) unsigned char ccsm_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ccsm_raw) + 1023) & ~(uintptr_t)1023);
    const int nkb = C / kCcBK;
    unsigned char* wsm = smem;                                   // [nkb][Wa' tile 16 KB | Wb tile 16 KB]
    unsigned char* ring = smem + nkb * 2 * kCcBM * kCcBK * 2;    // [kCcStages][feature k-block 16 KB]
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + kCcStages * kCcResStageBytes);
    uint64_t* empty = full + kCcStages;
    uint64_t* tfull = empty + kCcStages;    // [2] accumulators of buffer i are complete
    uint64_t* tempty = tfull + 2;           // [2] ... have been drained by the epilogue
    uint64_t* wfull = tempty + 2;
    uint64_t* wempty = wfull + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wempty + 1);
    float* a_sm = reinterpret_cast<float*>(ring + kCcStages * kCcResStageBytes + 256);  // [2][128]: spatial gate of the current / next tile
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nTilesN = (S + kCcBN - 1) / kCcBN, nMt = C / kCcBM;
    const int chunks = (nTilesN + tiles_per_item - 1) / tiles_per_item;
    const int nItems = B * nMt * chunks;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kCcStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kCcEpiWarps); }
        mbar_init(wfull, 1);
        mbar_init(wempty, 1);
        fence_mbar_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.wa) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.wb) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(4 * kCcBN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer
            unsigned kit = 0, nit = 0;
            for (int it = blockIdx.x; it < nItems; it += gridDim.x, ++nit) {
                const int ch = it % chunks, mt = (it / chunks) % nMt, b = it / (chunks * nMt);
                const int m0 = mt * kCcBM, t0 = ch * tiles_per_item, t1 = min(nTilesN, t0 + tiles_per_item);
                mbar_wait(wempty, (nit & 1) ^ 1);  // the previous item's MMAs have finished reading the weights
                mbar_expect_tx(wfull, (uint32_t)(nkb * 2 * kCcBM * kCcBK * 2));
                for (int kb = 0; kb < nkb; ++kb) {
                    tma_load_2d(wsm + kb * 2 * kCcBM * kCcBK * 2, &maps.wa, kb * kCcBK, MODE == 0 ? b * C + m0 : m0, wfull);
                    tma_load_2d(wsm + kb * 2 * kCcBM * kCcBK * 2 + kCcBM * kCcBK * 2, &maps.wb, kb * kCcBK, m0, wfull);
                }
                for (int t = t0; t < t1; ++t)
                    for (int kb = 0; kb < nkb; ++kb, ++kit) {
                        const int st = kit % kCcStages, ph = (kit / kCcStages) & 1;
                        mbar_wait(&empty[st], ph ^ 1);
                        unsigned char* sb = ring + st * kCcResStageBytes;
                        mbar_expect_tx(&full[st], kCcResStageBytes);
                        tma_load_2d(sb, &maps.x, t * kCcBN, b * C + kb * kCcBK, &full[st]);
                        tma_load_2d(sb + kCcBK * 64 * 2, &maps.x, t * kCcBN + 64, b * C + kb * kCcBK, &full[st]);
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer
            const uint32_t idesc = umma_instr_desc(std::is_same<T, __nv_bfloat16>::value ? 1 : 0, kCcBM, kCcBN);
            unsigned kit = 0, nit = 0, tt = 0;
            for (int it = blockIdx.x; it < nItems; it += gridDim.x, ++nit) {
                const int ch = it % chunks;
                const int t0 = ch * tiles_per_item, t1 = min(nTilesN, t0 + tiles_per_item);
                mbar_wait(wfull, nit & 1);
                for (int t = t0; t < t1; ++t, ++tt) {
                    const int buf = tt & 1;
                    mbar_wait(&tempty[buf], ((tt >> 1) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t tacc = tmem_base + buf * 2 * kCcBN;
                    for (int kb = 0; kb < nkb; ++kb, ++kit) {
                        const int st = kit % kCcStages, ph = (kit / kCcStages) & 1;
                        mbar_wait(&full[st], ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t sa = smem_u32(wsm + kb * 2 * kCcBM * kCcBK * 2);
                        const uint32_t sa2 = sa + kCcBM * kCcBK * 2, sb = smem_u32(ring + st * kCcResStageBytes);
#pragma unroll
                        for (int k = 0; k < kCcBK / 16; ++k) {
                            const uint64_t a1 = umma_smem_desc(sa + k * 32, 16, 1024);
                            const uint64_t a2 = umma_smem_desc(sa2 + k * 32, 16, 1024);
                            const uint64_t bd = umma_smem_desc(sb + k * 2048, kCcBK * 64 * 2, 1024);
                            const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
                            umma_f16(tacc, a1, bd, idesc, acc);
                            umma_f16(tacc + kCcBN, a2, bd, idesc, acc);
                        }
                        umma_commit(&empty[st]);
                    }
                    umma_commit(&tfull[buf]);
                }
                umma_commit(wempty);
            }
        }
    } else {
        // ---- epilogue warps
        const float alpha = softplusf_acc(__ldg(beta));
        const float k0 = pyramid_multiply ? 0.0f : 1.0f - alpha;
        unsigned tt = 0;
        for (int it = blockIdx.x; it < nItems; it += gridDim.x) {
            const int ch = it % chunks, mt = (it / chunks) % nMt, b = it / (chunks * nMt);
            const int t0 = ch * tiles_per_item, t1 = min(nTilesN, t0 + tiles_per_item);
            for (int t = t0; t < t1; ++t, ++tt) {
                const int buf = tt & 1;
                concat_stage_gate(a_sm + buf * kCcBN, agate, b, t * kCcBN, S);  // (its previous reader, tile tt - 2, is two barriers back)
                [[maybe_unused]] ConcatBwdRegs pre;
                if constexpr (MODE == 1) concat_epilogue_bwd_prefetch<T>(pre, warp, lane, x, bw, b, mt * kCcBM, t * kCcBN, C, S);
                mbar_wait(&tfull[buf], (tt >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if constexpr (MODE == 0)
                    concat_epilogue_tile<T>(tmem_base + buf * 2 * kCcBN, warp, lane, x, a_sm + buf * kCcBN, bias, out, b, mt * kCcBM, t * kCcBN, C, S, k0, alpha);
                else
                    concat_epilogue_bwd_tile<T>(tmem_base + buf * 2 * kCcBN, pre, warp, lane, a_sm + buf * kCcBN, bias, out, bw, b, mt * kCcBM, t * kCcBN, C,
                                                S, k0, alpha, pyramid_multiply);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty[buf])) : "memory");
            }
        }
    }
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(4 * kCcBN) : "memory");
}

}  // namespace mga

namespace mga {

// ---------------------------------------------------------------- backward of the concat forward: the elementwise / reduction part.
// With U = Wa^T g and V = Wb^T g (library GEMMs on the raw upstream gradient g, stacked as uv (B, 2C, S)):
//   dx      = alpha (s_c U + a_p V) + k0 g                     ga = g * a_p   (left operand of the dWb GEMM)
//   ds_c    = alpha sum_p x U          da_p = alpha sum_c x V   dbias_c = alpha sum_p g
//   dalpha  = sum x (s U + a V) + sum_c bias_c sum_p g - [add] sum g x
// grid (tiles of 32 units, B); warp w owns channels w, w+8, ...; a lane owns one 16-byte unit.  Per-tile partial sums go to scratch rows
// (ds_part / dbias_part: (B, nTiles, C), dalpha_part: (B, nTiles)); da is complete per pixel (a CTA covers every channel of its pixels).
template <typename T>
__global__ void __launch_bounds__(kBlock) concat_bwd_elem_kernel(const T* __restrict__ x, const T* __restrict__ g, const T* __restrict__ uv,
                                                                 const float* __restrict__ s, const float* __restrict__ agate,
                                                                 const float* __restrict__ bias, const float* __restrict__ beta, T* __restrict__ dx,
                                                                 T* __restrict__ ga, float* __restrict__ ds_part, float* __restrict__ dbias_part,
                                                                 float* __restrict__ da, float* __restrict__ dalpha_part, int C, int S,
                                                                 int pyramid_multiply) {
    constexpr int VEC = 8;
    __shared__ float part[kWarpsPerBlock][32 * VEC];
    __shared__ float red[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.y, tile = blockIdx.x, nT = gridDim.x, U = S / VEC;
    const int u = tile * 32 + lane;
    const bool act = u < U;
    const float alpha = softplusf_acc(__ldg(beta));
    const float k0 = pyramid_multiply ? 0.0f : 1.0f - alpha;
    float av[VEC], vacc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) { av[e] = 0.0f; vacc[e] = 0.0f; }
    if (act) ldf<VEC>(agate + (size_t)b * S + (size_t)u * VEC, av);
    float dal = 0.0f;
    for (int c = w; c < C; c += kWarpsPerBlock) {
        float su = 0.0f, sg = 0.0f;
        if (act) {
            const size_t o = ((size_t)b * C + c) * S + (size_t)u * VEC;
            const size_t ou = ((size_t)b * 2 * C + c) * S + (size_t)u * VEC;
            float xv[VEC], gv[VEC], uu[VEC], vv[VEC], ov[VEC], gav[VEC];
            ldv<T, VEC, kLdStream>(x + o, xv);
            ldv<T, VEC, kLdStream>(g + o, gv);
            ldv<T, VEC, kLdStream>(uv + ou, uu);
            ldv<T, VEC, kLdStream>(uv + ou + (size_t)C * S, vv);
            const float sc = __ldg(s + b * C + c);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const float t = fmaf(sc, uu[e], av[e] * vv[e]);
                ov[e] = fmaf(alpha, t, k0 * gv[e]);
                gav[e] = gv[e] * av[e];
                su = fmaf(xv[e], uu[e], su);
                sg += gv[e];
                vacc[e] = fmaf(xv[e], vv[e], vacc[e]);
                dal = fmaf(xv[e], t, dal);
                if (!pyramid_multiply) dal = fmaf(-gv[e], xv[e], dal);
            }
            stv<T, VEC, true>(dx + o, ov);
            stv<T, VEC, true>(ga + o, gav);
        }
        su = warp_sum(su);
        sg = warp_sum(sg);
        if (lane == 0) {
            const size_t pi = ((size_t)b * nT + tile) * C + c;
            ds_part[pi] = alpha * su;
            dbias_part[pi] = alpha * sg;
            dal = fmaf(__ldg(bias + c), sg, dal);
        }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) part[w][lane * VEC + e] = vacc[e];
    const float dtot = block_sum(dal, red);  // (barriers inside: `part` is complete)
    if (threadIdx.x == 0) dalpha_part[(size_t)b * nT + tile] = dtot;
    for (int i = threadIdx.x; i < 32 * VEC; i += kBlock) {
        const int p = tile * 32 * VEC + i;
        if (p >= S) continue;
        float r = 0.0f;
#pragma unroll
        for (int q = 0; q < kWarpsPerBlock; ++q) r += part[q][i];
        da[(size_t)b * S + p] = alpha * r;
    }
}


// ---------------------------------------------------------------- weight gradient: batch reduction of the per-sample GEMM results
//   dW[o][i]     = alpha * sum_b Ga[b][o][i] * s[b][i]      Ga = g X^T   (B, C, C)
//   dW[o][C + i] = alpha * sum_b Gb[b][o][i]                Gb = (g a) X^T
// One thread per (o, i), fixed summation order over the batch (deterministic), fp32 accumulation; TG = element type of the GEMM results.
template <typename TG>
__global__ void __launch_bounds__(kBlock) concat_wgrad_reduce_kernel(const TG* __restrict__ Ga, const TG* __restrict__ Gb, const float* __restrict__ s,
                                                                     const float* __restrict__ beta, float* __restrict__ dw, int B, int C) {
    const int e = blockIdx.x * kBlock + threadIdx.x;
    if (e >= C * C) return;
    const int o = e / C, i = e - o * C;
    const size_t CC = (size_t)C * C;
    float accA = 0.0f, accB = 0.0f;
    int b = 0;
    for (; b + 4 <= B; b += 4) {
        float ga[4], gb[4], sv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            ga[q] = to_f<TG>(Ga[(size_t)(b + q) * CC + e]);
            gb[q] = to_f<TG>(Gb[(size_t)(b + q) * CC + e]);
            sv[q] = __ldg(s + (size_t)(b + q) * C + i);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) { accA = fmaf(ga[q], sv[q], accA); accB += gb[q]; }
    }
    for (; b < B; ++b) {
        accA = fmaf(to_f<TG>(Ga[(size_t)b * CC + e]), __ldg(s + (size_t)b * C + i), accA);
        accB += to_f<TG>(Gb[(size_t)b * CC + e]);
    }
    const float alpha = softplusf_acc(__ldg(beta));
    dw[(size_t)o * 2 * C + i] = alpha * accA;
    dw[(size_t)o * 2 * C + C + i] = alpha * accB;
}

}  // namespace mga
