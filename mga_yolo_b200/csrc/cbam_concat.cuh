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
constexpr int kCcThreads = 192;
constexpr int kCcStageBytes = (2 * kCcBM * kCcBK + kCcBK * kCcBN) * 2;  // two weight tiles + the feature tile, 16-bit elements
constexpr int kCcSmemBytes = kCcStages * kCcStageBytes + 1024 /* alignment slack */ + 256 /* barriers */;

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
template <typename T>
__global__ void __launch_bounds__(kBlock) concat_fold_kernel(const float* __restrict__ w, const float* __restrict__ s, T* __restrict__ wa, T* __restrict__ wb,
                                                             int B, int C) {
    const size_t n = (size_t)(B + 1) * C * C;
    for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (size_t)gridDim.x * kBlock) {
        const int b = (int)(i / ((size_t)C * C));
        const int r = (int)(i - (size_t)b * C * C);
        const int co = r / C, ci = r - co * C;
        if (b < B) wa[i] = from_f<T>(__ldg(w + (size_t)co * 2 * C + ci) * __ldg(s + b * C + ci));
        else wb[r] = from_f<T>(__ldg(w + (size_t)co * 2 * C + C + ci));
    }
}

template <typename T>
__global__ void __launch_bounds__(kCcThreads, 1) concat_fwd_kernel(const __grid_constant__ ConcatMaps maps, const T* __restrict__ x,
                                                                   const float* __restrict__ agate, const float* __restrict__ bias,
                                                                   const float* __restrict__ beta, T* __restrict__ out, int C, int S, int pyramid_multiply) {
    extern __shared__ __align__(1024) unsigned char ccsm_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ccsm_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kCcStages * kCcStageBytes);
    uint64_t* empty = full + kCcStages;
    uint64_t* tmem_full = empty + kCcStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
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
                tma_load_2d(sa, &maps.wa, kb * kCcBK, b * C + m0, &full[st]);
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
        // ---- epilogue: warp w reads TMEM lanes 32 (w % 4) ..: row co = m0 + 32 (w % 4) + lane, 128 pixels in chunks of 16 columns
        const int q = warp & 3;
        const int co = m0 + q * 32 + lane;
        const float bt = __ldg(beta);
        const float alpha = softplusf_acc(bt);
        const float k0 = pyramid_multiply ? 0.0f : 1.0f - alpha, k1 = alpha;
        const float bco = __ldg(bias + co);
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        const size_t rowoff = ((size_t)b * C + co) * S;
        const float* ap = agate + (size_t)b * S;
#pragma unroll 1
        for (int c16 = 0; c16 < kCcBN / 16; ++c16) {
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
                const float4 a0 = __ldg(reinterpret_cast<const float4*>(ap + ph8));
                const float4 a1 = __ldg(reinterpret_cast<const float4*>(ap + ph8 + 4));
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
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * kCcBN) : "memory");
}

}  // namespace mga
