// cbam_flow.cuh -- wavefront-ordered dataflow kernels: ALL phases of the forward (or backward) pass in ONE launch.
//
// Why: as separate kernels the passes re-read x (and g) from HBM -- at batch 64 a 105 MB feature map does not survive in
// L2 from one kernel to the next -- and every small phase pays a ~4 us launch floor.  Here every phase is a ROLE of one
// kernel.  The grid is a sequence of "steps"; step t holds the CTAs of  phase_0(sample t), phase_1(sample t - L),
// phase_2(sample t - 2L) ...  so the phases of one sample are dispatched a few microseconds apart: the sample's planes
// (1.6 MB at P3) are still in L2 when the next phase reads them, and only ~5L samples are in flight at any time.
// A role CTA starts by waiting for the per-sample completion counter of the phase it depends on; because that phase was
// dispatched L steps earlier (lower block indices are dispatched first) the wait is normally already satisfied, and the
// lowest-numbered unfinished CTA never waits on anything that is not already resident -> no deadlock.  The spin is
// bounded anyway: on time-out an error flag is raised and the CTA continues (the launch never hangs).
//
// Memory model: producers write, __threadfence(), barrier, then one atomicAdd on the counter; consumers spin with
// ld.acquire.gpu, barrier, and read the produced planes / statistics with ld.global.cg (never the read-only path).
#pragma once
#include "cbam_bwd.cuh"
#include "cbam_conv.cuh"
#include "cbam_fwd.cuh"
#include "common.cuh"

namespace mga {

constexpr int kFlowMaxRoles = 8;
struct FlowSched {
    int nRoles;
    int cnt[kFlowMaxRoles];      // CTAs of each role per sample
    int lag[kFlowMaxRoles];      // step offset of each role
    int dep[kFlowMaxRoles];      // role whose per-sample counter must reach cnt[dep] first (-1: none)
    int pre[kFlowMaxRoles + 1];  // prefix sums of cnt
    int per_step;
    int steps;
    unsigned* done;              // [nRoles][B] completion counters (zeroed before the launch)
    int* err;                    // raised when a wait timed out
};

__device__ __forceinline__ void flow_wait(const unsigned* ctr, unsigned expected, int* err) {
    if (threadIdx.x == 0) {
        unsigned v;
        int spins = 0;
        while (true) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if (v >= expected) break;
            __nanosleep(64);
            if (++spins > (1 << 22)) { *err = 1; break; }  // ~0.3 s: never hang the GPU
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void flow_signal(unsigned* ctr) {
    __syncthreads();  // every thread's writes happen-before thread 0 (CTA scope) ...
    if (threadIdx.x == 0) {
        __threadfence();  // ... and the cumulative gpu-scope fence publishes them before the counter moves
        atomicAdd(ctr, 1u);
    }
}

// plane class (chosen from the per-sample plane size only): threads per plane of the pooling role, tile shape of the reduce roles
template <int CLS, int VEC> struct PlaneClass;
template <int VEC> struct PlaneClass<0, VEC> { static constexpr int TPP = 256, LPT = 32, UPT = (VEC == 8 ? 1 : 2); };  // U >= 1024
template <int VEC> struct PlaneClass<1, VEC> { static constexpr int TPP = 128, LPT = 16, UPT = 1; };                    // 256 <= U < 1024
template <int VEC> struct PlaneClass<2, VEC> { static constexpr int TPP = 32, LPT = 8, UPT = 1; };                      // U < 256

enum { kFwdMask = 0, kFwdPool, kFwdMlp, kFwdReduce, kFwdConv, kFwdRescale, kFwdRoles };

template <typename T, int CLS>
__global__ void __launch_bounds__(kBlock, 4) flow_fwd_kernel(const T* __restrict__ x, const void* __restrict__ mask, int mdt, T* __restrict__ out,
                                                             Shape sh, mga_cbam_params prm, Ctx ctx, FwdScratch fs, ConvGeom cg, FlowSched sc,
                                                             int nMaskTiles) {
    constexpr int VEC = 16 / sizeof(T);
    using PC = PlaneClass<CLS, VEC>;
    extern __shared__ __align__(128) float dsm_flow[];
    const int step = blockIdx.x / sc.per_step, off = blockIdx.x - step * sc.per_step;
    int role = 0;
    while (off >= sc.pre[role + 1]) ++role;
    const int tile = off - sc.pre[role];
    const int b = step - sc.lag[role];
    if (b < 0 || b >= sh.B) return;
    const int dep = sc.dep[role];
    if (dep >= 0 && sc.cnt[dep] > 0) flow_wait(sc.done + (size_t)dep * sh.B + b, (unsigned)sc.cnt[dep], sc.err);

    switch (role) {
        case kFwdMask: {
            const Blk blk{tile, b, 0, nMaskTiles};
            if (mdt == MGA_F32) mask_prep_body<float>(static_cast<const float*>(mask), sh, ctx, fs, blk);
            else if (mdt == MGA_BF16) mask_prep_body<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(mask), sh, ctx, fs, blk);
            else mask_prep_body<__half>(static_cast<const __half*>(mask), sh, ctx, fs, blk);
            break;
        }
        case kFwdPool: {
            const Blk blk{b * sc.cnt[kFwdPool] + tile, 0, 0, 0};  // C is a multiple of the planes per CTA (checked on the host)
            cam_pool_body<T, VEC, PC::TPP>(x, sh, ctx, fs, blk);
            break;
        }
        case kFwdMlp: {
            const Blk blk{b, 0, 0, 0};
            cam_mlp_body(sh, prm, ctx, fs, nMaskTiles, blk, dsm_flow);
            break;
        }
        case kFwdReduce: {
            const Blk blk{tile, b, 0, 0};
            sam_reduce_body<T, VEC, PC::LPT, PC::UPT>(x, sh, ctx, blk);
            break;
        }
        case kFwdConv: {
            const Blk blk{tile, b, 0, 0};
            sam_conv4_body(sh, prm.wsam, ctx, cg, nullptr, blk, dsm_flow);
            break;
        }
        default: {
            const float alpha = softplusf_acc(__ldg(prm.beta));
            const float k0 = sh.pyramid_multiply() ? 0.0f : 1.0f - alpha;
            rescale_tile_body<T, VEC>(x, out, sh, ctx, k0, alpha, b, tile);
            break;
        }
    }
    flow_signal(sc.done + (size_t)role * sh.B + b);
}

}  // namespace mga
