// common.cuh -- shared device helpers for the mask-guided CBAM kernels (sm_100a).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <type_traits>

#include "mga_cbam.h"

namespace mga {

constexpr int kWarp = 32;
constexpr int kBlock = 256;          // threads per CTA for the streaming kernels
constexpr int kWarpsPerBlock = kBlock / kWarp;
constexpr int kSMs = 148;            // B200
// tuning knobs of the tile kernels (sam_reduce, bwd_reduce1/2, bwd_dx): resident CTAs per SM the register
// allocation must allow, and channels per load batch (x-only kernels / x+g kernels)
#ifndef MGA_TILE_MINB_16
#define MGA_TILE_MINB_16 3  // ... of their 16-bit instantiations (80 registers: six raw loads in flight; measured on cfg5: 0.650 -> 0.630 ms)
#endif
#ifndef MGA_TILE_MINB
#define MGA_TILE_MINB 4
#endif
#ifndef MGA_KB1
#define MGA_KB1 4
#endif
#ifndef MGA_KB1_16
#define MGA_KB1_16 6  // ... of the 16-bit instantiations, which hold their loads RAW (4 registers per 16 bytes) and unpack at use: as many bytes in flight as fp32 (was 2 loads: unpacked fp32 registers)
#endif
#ifndef MGA_KB2_16
#define MGA_KB2_16 2  // 16-bit instantiations of the (x, g) kernels: channels per batch, loads held raw (was 1)
#endif
#ifndef MGA_KB2
#define MGA_KB2 2
#endif
constexpr int kMaxK = 7;             // largest spatial-attention kernel that is built
constexpr int kConvTW = 32, kConvTH = 8;  // output tile of the plane-convolution kernels

// ---------------------------------------------------------------- element conversion
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

template <typename T> struct VecOf { static constexpr int V = 16 / sizeof(T); };  // elements per 128-bit access

// L2 eviction-priority policies for 128-bit accesses (the plain .L2::evict_* qualifiers need 256-bit accesses on sm_100).
// keep  : data that a later kernel of the same forward/backward re-reads (x after pooling, x/g after the first reduce)
// last  : final read of that data -> first candidate for eviction
#ifndef MGA_L2_HINTS
#define MGA_L2_HINTS 0  // measured on B200 (r1): no gain at batch 64 (x does not stay L2-resident between passes); kept for chunked launches
#endif
enum LdMode { kLdDefault = 0, kLdStream = 1, kLdKeepL2 = 2, kLdLastUse = 3 };
__device__ __forceinline__ unsigned long long l2_policy_keep() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// The same load kept RAW (4 registers per 16 bytes whatever the element type) and unpacked when it is consumed: a 16-bit unit unpacks to
// 8 fp32 registers, so loops that keep several loads in flight hold them raw (twice the bytes in flight per register).
struct RawV { uint4 t; };
template <typename T, int VEC, int MODE = 0>
__device__ __forceinline__ RawV ldraw(const T* __restrict__ p) {
    constexpr int kMode = (MGA_L2_HINTS || MODE < 2) ? MODE : 1;
    RawV r;
    if constexpr (VEC == 1) {
        r.t = make_uint4(__float_as_uint(to_f<T>(p[0])), 0u, 0u, 0u);
    } else if constexpr (kMode == kLdStream) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.t.x), "=r"(r.t.y), "=r"(r.t.z), "=r"(r.t.w) : "l"(p));
    } else if constexpr (kMode == kLdKeepL2) {
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=r"(r.t.x), "=r"(r.t.y), "=r"(r.t.z), "=r"(r.t.w) : "l"(p), "l"(l2_policy_keep()));
    } else if constexpr (kMode == kLdLastUse) {
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=r"(r.t.x), "=r"(r.t.y), "=r"(r.t.z), "=r"(r.t.w) : "l"(p), "l"(l2_policy_last()));
    } else {
        r.t = __ldg(reinterpret_cast<const uint4*>(p));
    }
    return r;
}
template <typename T, int VEC>
__device__ __forceinline__ void unpackv(const RawV& r, float (&v)[VEC]) {
    if constexpr (VEC == 1) {
        v[0] = __uint_as_float(r.t.x);
    } else if constexpr (sizeof(T) == 4) {
        static_assert(VEC == 4, "fp32 vector is 4 wide");
        v[0] = __uint_as_float(r.t.x); v[1] = __uint_as_float(r.t.y); v[2] = __uint_as_float(r.t.z); v[3] = __uint_as_float(r.t.w);
    } else {
        static_assert(VEC == 8, "16-bit vector is 8 wide");
        const uint32_t w[4] = {r.t.x, r.t.y, r.t.z, r.t.w};
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

// 128-bit (or scalar) global load of VEC consecutive elements into fp32 registers.
// MODE: true / kLdStream = no L1 allocation (data touched once by this kernel); kLdKeepL2 / kLdLastUse add an L2 policy.
template <typename T, int VEC, int MODE = 0>
__device__ __forceinline__ void ldv(const T* __restrict__ p, float (&v)[VEC]) {
    constexpr int kMode = (MGA_L2_HINTS || MODE < 2) ? MODE : 1;
    if constexpr (VEC == 1) {
        v[0] = to_f<T>(p[0]);
    } else {
        uint4 t;
        if constexpr (kMode == kLdStream) {
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p));
        } else if constexpr (kMode == kLdKeepL2) {
            asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                         : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p), "l"(l2_policy_keep()));
        } else if constexpr (kMode == kLdLastUse) {
            asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                         : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p), "l"(l2_policy_last()));
        } else {
            t = __ldg(reinterpret_cast<const uint4*>(p));
        }
        if constexpr (sizeof(T) == 4) {
            static_assert(VEC == 4, "fp32 vector is 4 wide");
            v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
        } else {
            static_assert(VEC == 8, "16-bit vector is 8 wide");
            const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                    v[2 * i] = __uint_as_float(w[i] << 16);
                    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
                } else {
                    const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
                    const float2 f = __half22float2(h);
                    v[2 * i] = f.x;
                    v[2 * i + 1] = f.y;
                }
            }
        }
    }
}

template <typename T, int VEC, bool kStream = false>
__device__ __forceinline__ void stv(T* __restrict__ p, const float (&v)[VEC]) {
    if constexpr (VEC == 1) {
        p[0] = from_f<T>(v[0]);
    } else if constexpr (sizeof(T) == 4) {
        float4 t = make_float4(v[0], v[1], v[2], v[3]);
        if constexpr (kStream) __stcs(reinterpret_cast<float4*>(p), t);
        else *reinterpret_cast<float4*>(p) = t;
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
        uint4 t = make_uint4(w[0], w[1], w[2], w[3]);
        if constexpr (kStream) __stcs(reinterpret_cast<uint4*>(p), t);
        else *reinterpret_cast<uint4*>(p) = t;
    }
}

// fp32 / int32 planes (mask, attention map, indices) read VEC at a time
template <int VEC>
__device__ __forceinline__ void ldf(const float* __restrict__ p, float (&v)[VEC]) {
    if constexpr (VEC == 1) {
        v[0] = __ldcg(p);
    } else {
#pragma unroll
        for (int i = 0; i < VEC / 4; ++i) {
            const float4 t = __ldcg(reinterpret_cast<const float4*>(p) + i);
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
    }
}
template <int VEC>
__device__ __forceinline__ void ldi(const int* __restrict__ p, int (&v)[VEC]) {
    if constexpr (VEC == 1) {
        v[0] = __ldcg(p);
    } else {
#pragma unroll
        for (int i = 0; i < VEC / 4; ++i) {
            const int4 t = __ldcg(reinterpret_cast<const int4*>(p) + i);
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
    }
}
template <int VEC>
__device__ __forceinline__ void stf(float* __restrict__ p, const float (&v)[VEC]) {
    if constexpr (VEC == 1) {
        p[0] = v[0];
    } else {
#pragma unroll
        for (int i = 0; i < VEC / 4; ++i)
            reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
}
template <int VEC>
__device__ __forceinline__ void sti(int* __restrict__ p, const int (&v)[VEC]) {
    if constexpr (VEC == 1) {
        p[0] = v[0];
    } else {
#pragma unroll
        for (int i = 0; i < VEC / 4; ++i)
            reinterpret_cast<int4*>(p)[i] = make_int4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
}

// ---------------------------------------------------------------- math
__device__ __forceinline__ float sigmoidf_acc(float v) { return 1.0f / (1.0f + expf(-v)); }  // same form ATen uses
__device__ __forceinline__ float softplusf_acc(float b) { return b > 20.0f ? b : log1pf(expf(b)); }  // F.softplus threshold=20

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide fp64 sum (used only for the few scalars that feed d beta)
__device__ __forceinline__ double block_sum_d(double v, double* sh /* >= 32 doubles */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum_d(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double r = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0.0;
    return warp_sum_d(r);
}

// block-wide sum for kBlock threads; result valid in every thread
__device__ __forceinline__ float block_sum(float v, float* sh /* >= 32 floats */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    float r = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0.0f;
    r = warp_sum(r);
    return r;
}

// ---------------------------------------------------------------- PTX helpers (mbarrier + TMA)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> this CTA's shared memory, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 3-D tiled TMA load (tensor map made by cuTensorMapEncodeTiled on the host): box -> this CTA's shared memory, out-of-bounds
// elements arrive as zeros, completion counted on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_3d(void* dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}


// ---------------------------------------------------------------- small helpers shared by the cluster and per-phase kernels
constexpr int kSmemLimit = 232448;  // 227 KB opt-in maximum per CTA on sm_100

// In-kernel phase timeline (tools/timeline.py): compiled in only with -DMGA_TUNING; thread 0 of each CTA stamps %globaltimer.
#ifdef MGA_TUNING
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
#else
__device__ __forceinline__ void stamp(int) {}
#endif

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

// ---------------------------------------------------------------- block coordinates as data
// Every phase is written as a __device__ body that takes its block coordinates as an argument, so the same code runs
// as its own kernel (coordinates = blockIdx); the indirection lets tools re-drive a phase with synthetic coordinates.
struct Blk {
    int x, y, z, gx;  // blockIdx.x/y/z and gridDim.x of the equivalent stand-alone launch
};
__device__ __forceinline__ Blk this_block() { return Blk{(int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z, (int)gridDim.x}; }

// loads of data that ANOTHER CTA of the same kernel may have produced (dataflow mode): L2-coherent, never the read-only path
__device__ __forceinline__ float ldc(const float* p) { return __ldcg(p); }
__device__ __forceinline__ int ldc(const int* p) { return __ldcg(p); }

// ---------------------------------------------------------------- tile kernels: thread -> (channel, unit) mapping
// A CTA of 8 warps owns a tile of LPT*UPT units (16-byte groups of pixels) of one sample and ALL channels.
// A warp load covers 32/LPT channels x LPT units (each channel's LPT units are contiguous: LPT*16 B >= 128 B), so small
// feature maps (few units per sample) still produce enough CTAs: LPT = 32 for large planes, 16 / 8 for small ones.
template <int LPT, int UPT, int VEC>
struct TileMap {
    static constexpr int CPW = 32 / LPT;                       // channels per warp load
    static constexpr int kChanStep = kWarpsPerBlock * CPW;     // channel stride between two loads of the same thread
    static constexpr int TU = LPT * UPT;                       // units per tile
    static constexpr int TP = TU * VEC;                        // pixels per tile
    int w, sub, ul;
    __device__ __forceinline__ TileMap() {
        const int lane = threadIdx.x & 31;
        w = threadIdx.x >> 5;
        sub = lane / LPT;
        ul = lane % LPT;
    }
    __device__ __forceinline__ int chan0() const { return w * CPW + sub; }                // first channel of this thread
    __device__ __forceinline__ int unit(int tile, int k) const { return tile * TU + k * LPT + ul; }
    __device__ __forceinline__ int slot(int k, int i) const { return (k * LPT + ul) * VEC + i; }  // pixel slot inside the tile
};
// sum over the LPT lanes that share a channel (result in every lane of the group)
template <int LPT>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = LPT / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// sum over the 32/LPT channel groups of a warp (lanes with equal `ul`); result in every lane
template <int LPT>
__device__ __forceinline__ float cross_group_sum(float v) {
#pragma unroll
    for (int o = LPT; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------- shapes, context and scratch layouts
struct Shape {
    int B, C, H, W, S, hidden, k;
    int flags;
    float tiny_thr, eps;
    __host__ __device__ bool has_mask() const { return flags & MGA_HAS_MASK; }
    __host__ __device__ bool sigmoid_mask() const { return flags & MGA_SIGMOID_MASK; }
    __host__ __device__ bool gate_clamp() const { return flags & MGA_GATE_CLAMP; }
    __host__ __device__ bool samcam_add() const { return flags & MGA_SAMCAM_ADD; }
    __host__ __device__ bool pyramid_multiply() const { return flags & MGA_PYRAMID_MULTIPLY; }
    __host__ __device__ bool gates_only() const { return flags & MGA_GATES_ONLY; }
    __host__ __device__ bool gates_acc() const { return flags & MGA_GATES_ACC; }
    __host__ __device__ bool no_save() const { return flags & MGA_NO_SAVE; }
};

// saved-for-backward context (device pointers into the caller's ctx buffer)
struct Ctx {
    float* consts;  // [0]=k0 [1]=k1 [2]=alpha [3]=sigmoid(beta)
    float* m;       // (B,S) mask after clamp/sigmoid; zeros without a mask
    float* a;       // (B,S) spatial attention map
    float* pmax;    // (B,S) channel max of x*q
    float* pavg;    // (B,S) channel mean of x*q
    int* idx;       // (B,S) arg max channel
    float* s;       // (B,C) channel attention
    float* avg;     // (B,C) pooled descriptor 1 (after fall-back blend)
    float* mx;      // (B,C) pooled descriptor 2 (after fall-back)
    float* apool;   // (B,C) masked average A (before blend)
    int* amax;      // (B,C) arg max pixel of the masked max, -1 when no pixel was valid
    float* use;     // (B)   1 when mean(mask) >= tiny_thr
    float* den;     // (B)   clamp_min(sum(mask), eps)
    float* msum;    // (B)   sum(mask)
    float* ha;      // (B,hidden) relu(W1 avg + b1)
    float* hm;      // (B,hidden) relu(W1 mx + b1)
};

struct FwdScratch {
    float* sxm;   // (B,C) sum x*m
    float* sx;    // (B,C) sum x
    float* best;  // (B,C) masked max
    int* bidx;    // (B,C) its pixel
    float* mpart; // (B,nMaskTiles) partial sums of the mask
};

struct BwdScratch {
    float* T;        // (B,S)   sum_c g*x*q
    float* dcat;     // (3,B,S) gradient wrt [pmax, pavg, m] planes
    float* epart;    // (B,nT,C) partial sum_p g*x*(a or 1)
    float* gxpart;   // (B,nT,C) partial sum_p g*x
    float* qpart;    // (B,nT,C) partial sum_p x*dy1            (multiply mode)
    float* psum;     // (3,B,C) the three partial arrays summed over their tiles (large planes: bwd_partsum_kernel)
    float* cA;       // (B,C) use*davg/den
    float* cG;       // (B,C) ((1-use)*davg + dead*dmx)/S
    float* cM;       // (B,C) (1-dead)*dmx
    float* kb;       // (B)   sum_c cA*A*[msum>=eps]
    float* dz;       // (B,C)
    float* dha;      // (B,hidden)
    float* dhm;      // (B,hidden)
    float* convpart; // (nConvCta, 3*k*k + 1)  dWsam partials + sum a*T
    double* alphapart;// (B) per-sample pieces of d alpha that come from per-channel sums (fp64: d beta sums all N elements)
    double* atpart;   // (nConvCta) per-CTA sum_p a_p T_p
};

// records a message for mga_last_error() and returns `code` (defined in mga_cbam.cu)
int fail(int code, const char* fmt, ...);

inline __host__ size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

}  // namespace mga
