// cbam_cluster.cuh -- cluster-per-sample fused forward / backward kernels (the default path for samples that fit).
//
// One thread-block CLUSTER owns one sample for the whole pass.  CTA r of the cluster owns image rows
// [r*rowsPer, (r+1)*rowsPer) of that sample for ALL channels, every phase of the pass runs inside the same kernel and
// phases are ordered by the hardware cluster barrier.  Nothing of the sample is staged in shared memory: the first
// pass over x (and g) comes from HBM, the later passes re-read the same lines a few microseconds later, while only
// (resident CTAs) x (~200 KB) ~ 60 MB of feature map are in flight on the whole GPU -- they are L2 hits.  HBM therefore
// sees the algorithmic traffic only: forward 2N (+ planes), backward 3N (+ planes), instead of 4N / 6N for the
// one-kernel-per-phase path (cbam_fwd.cuh / cbam_bwd.cuh), and the six small kernels per direction disappear.
//
// Cross-CTA data is tiny and travels through distributed shared memory (DSMEM):
//   * per-channel partial sums are PUSHED into every CTA's shared memory before a cluster barrier, so after the barrier
//     every CTA recomputes the tiny shared MLP (forward or backward) from local data;
//   * the 7x7 convolution needs 3 halo rows of the 3-plane map (forward) / of dpre (backward) from the neighbouring
//     CTAs' tiles: read through DSMEM after a cluster barrier.
//
// Two thread -> data mappings (both runtime-shaped, so one instantiation per element type).  The kernels are
// instruction-issue bound before they are bandwidth bound, so both are written for few instructions per element:
// addresses are one mad.wide per access off a running row pointer, per-channel constants sit in shared memory in the
// order the thread walks them (128-bit reads), reductions across threads go through shared memory, not shuffles.
//   T1 "channel rows": LPT consecutive threads own one channel slot, thread = K units (16 B of pixels) of the CTA's
//      pixel range -> per-CHANNEL sums are thread-private; the LPT partials of a channel are staged in shared memory and
//      summed in a fixed order afterwards (pool, bwd reduce 1 / 2);
//   T2 "pixel columns": thread = (unit, channel group), loops over the channels of its group -> per-PIXEL sums / max are
//      thread-private (channel max/mean, rescale, dx); the G channel groups merge through shared memory.
//
// Reference lines: mga_yolo/nn/modules/masked_cbam.py:87-171 (forward); SURVEY.md section 8a "Backward" (closed form).
#pragma once
#include <cooperative_groups.h>

#include "cbam_conv.cuh"
#include "common.cuh"

namespace mga {
namespace cgx = cooperative_groups;

#ifndef MGA_CL_NT_F
#define MGA_CL_NT_F 512  // threads per CTA of the forward kernel (64 registers per thread, 2 CTAs per SM = 32 warps per SM)
#endif
#ifndef MGA_CL_NT_B
#define MGA_CL_NT_B 256  // ... of the backward kernel (128 registers per thread)
#endif
constexpr int kClNTF = MGA_CL_NT_F, kClNTB = MGA_CL_NT_B;
constexpr int kClWmax = 16;  // most warps per CTA (sizes the warp-private bins of the backward kernel)
#ifndef MGA_CL_K_F
#define MGA_CL_K_F 5
#endif
#ifndef MGA_CL_K_B
#define MGA_CL_K_B 5
#endif
#ifndef MGA_CL_DXKB
#define MGA_CL_DXKB 4
#endif
#ifndef MGA_CL_K_B16
#define MGA_CL_K_B16 3  // 16-bit backward: 8-element units double the per-slot T partials; 3 units keep 2 CTAs per SM (shared memory)
#endif
constexpr int kClKF = MGA_CL_K_F, kClKB = MGA_CL_K_B, kClKB16 = MGA_CL_K_B16;  // T1: units per thread and channel

#ifndef MGA_CL_HINTS
#define MGA_CL_HINTS 1  // L2 eviction-priority hints: first touch of x/g = keep, last read = evict first
#endif
#ifndef MGA_CL_GPOL
#define MGA_CL_GPOL kPolKeep  // L2 policy of the FIRST read of grad_out in the backward kernel
#endif
#ifndef MGA_CL_MINB_F
#define MGA_CL_MINB_F 2  // resident CTAs per SM the register allocation of the forward kernel must allow
#endif
#ifndef MGA_CL_MINB_B
#define MGA_CL_MINB_B 2  // ... of the backward kernel
#endif

struct ClGeom {
    int NT;        // threads per CTA
    int CS;        // CTAs per cluster (= per sample)
    int rowsPer;   // image rows per CTA
    int nUmax;     // 16-byte units per CTA (rowsPer * W / VEC)
    int nPmax;     // pixels per CTA
    int G;         // T2: channel groups = NT / nUmax
    int CG;        // T2: channels per group, rounded up to a multiple of 8 (group-major constant arrays)
    int K;         // T1: units per thread and channel of this launch
    int LPT;       // T1: threads per channel slot = ceil(nUmax / K)
    int slots;     // T1: channel slots = NT / LPT
    int tileRows;  // rowsPer + 6
    int TWp;       // W + 8
    int planeT;    // floats per tile plane (multiple of 32)
    int smem_bytes;
    int stage_w;   // 1: the cam_mlp weight matrices are staged in shared memory (decided on the host: only when the second resident CTA per SM survives)
    int prefetch;  // 1: bulk L2 prefetch of the sample at kernel start, so the HBM stream overlaps the latency-bound prologue
    int use_tma;   // backward: 1 = the padded [pmax, pavg, m] tile arrives by tensor-map TMA (host made the maps), 0 = plain loads
};

// the two weight matrices of the shared MLP may be staged in shared memory (16-byte cp.async copies) when they are small
__host__ __device__ inline bool cl_mlp_in_smem(int C, int Hd) { return (C * Hd) % 4 == 0 && 2 * C * Hd * 4 <= 16 * 1024; }
constexpr int kClTwoCtaSmem = 114 * 1024;  // dynamic shared memory per CTA up to which two CTAs fit one SM (228 KB, 1 KB reserved each)

// ---------------------------------------------------------------- shared-memory layouts (float offsets)
struct ClFwdOff { int wk, red, avg, mx, ha, hm, sAB, q, part, msum, mloc, aloc, tile, mg, stage, mlpw, total; };
__host__ __device__ inline ClFwdOff cl_fwd_off(int C, int Hd, const ClGeom& g) {
    ClFwdOff o;
    int p = 0;
    auto take = [&p](int n) { const int r = p; p += (n + 3) & ~3; return r; };
    o.wk = take(3 * kMaxK * kMaxK);
    o.red = take(64);
    o.avg = take(C); o.mx = take(C); o.ha = take(Hd); o.hm = take(Hd);
    o.sAB = take(2 * g.G * g.CG);   // group-major (A,B) pairs of the rescale
    o.q = take(g.G * g.CG);         // group-major q
    o.part = take(4 * C * g.CS);    // pool partials of every rank (pushed)
    o.msum = take(16);
    o.mloc = take(g.LPT * g.K * 8);  // m of the own pixels, zero padded to the T1 footprint
    // the T1 per-thread partials [4][C][LPT|1] of the pooling phase are dead before the planes exist: same storage
    o.stage = p;
    o.aloc = take(g.nPmax);
    o.tile = take(3 * g.planeT);
    o.mg = take(3 * g.G * g.nPmax);  // T2 merge buffers [max | sum | idx][G][nPmax]; later the conv's per-plane partials [3][nPmax]
    const int stage_end = o.stage + ((4 * C * (g.LPT | 1) + 3) & ~3);
    if (p < stage_end) p = stage_end;
    o.mlpw = g.stage_w ? take(2 * C * Hd) : -1;  // cam_mlp weights [W1 (Hd,C) | W2 (C,Hd)], staged with cp.async at kernel start
    o.total = p;
    return o;
}

struct ClBwdOff {
    int wsm, red, bar, s, chA, chG, amx, cM, dz, eloc, qloc, binloc, dha, dhm, q, epart, qpart, binpart, gxpart;
    int aloc, ae, mloc, idx, pmx, d0, d1s, d2, dpre, cat, tp, dwp, binw, stage, mlpw, total;
};
__host__ __device__ inline ClBwdOff cl_bwd_off(int C, int Hd, const ClGeom& g) {
    ClBwdOff o;
    int p = 0;
    auto take = [&p](int n) { const int r = p; p += (n + 3) & ~3; return r; };
    o.wsm = take(3 * kMaxK * kMaxK);
    o.red = take(64);  // 32 doubles
    o.bar = take(4);   // two mbarriers: [0] context planes (bulk copies), [1] the [pmax, pavg, m] tile (tensor-map TMA)
    o.s = take(C);
    o.chA = take(4 * g.G * g.CG);  // group-major (A, B, q, cA)
    o.chG = take(g.G * g.CG);      // group-major cG
    o.amx = take(C); o.cM = take(C);
    o.dz = take(C); o.eloc = take(C); o.qloc = take(C); o.binloc = take(C);
    o.dha = take(Hd); o.dhm = take(Hd);
    o.q = take(C);
    o.epart = take(C * g.CS); o.qpart = take(C * g.CS); o.binpart = take(C * g.CS);
    o.gxpart = take(2 * 16);  // doubles
    const int padP = g.LPT * g.K * 8;  // T1 footprint in pixels (>= nPmax), zero padded
    o.aloc = take(g.nPmax); o.ae = take(padP); o.mloc = take(g.nPmax); o.idx = take(g.nPmax); o.pmx = take(g.nPmax);
    o.d0 = take(g.nPmax); o.d1s = take(padP); o.d2 = take(g.nPmax);
    p = (p + 31) & ~31;  // 128-byte alignment: TMA destinations
    o.dpre = take(g.planeT);
    // One region, three tenants with disjoint lifetimes:
    //   phase 1      : tp (T partials of the channel slots) | stage (per-channel thread partials)
    //   phase 2      : cat = the [pmax, pavg, m] tile (lands asynchronously after the T merge) | dwp (dW team partials)
    //   phase 3 / 5  : stage again (Q partials) / tp again (R partials of the G channel groups)
    const int tpn = (((g.slots > g.G ? g.slots : g.G) * g.nPmax) + 3) & ~3;
    const int stagen = (C * (g.LPT | 1) + 3) & ~3;
    const int dwpn = 21 * 12 * kMaxK;
    const int a = tpn + stagen, b = 3 * g.planeT + dwpn;
    o.tp = take(a > b ? a : b);
    o.stage = o.tp + tpn;
    o.cat = o.tp;
    o.dwp = o.cat + 3 * g.planeT;
    o.binw = take((g.NT / 32) * C);
    o.mlpw = g.stage_w ? take(2 * C * Hd) : -1;  // cam_mlp weights [W1 | W2] staged with cp.async
    o.total = p;
    return o;
}

// ---------------------------------------------------------------- 128-bit global access with optional L2 policy
enum { kPolNone = 0, kPolKeep = 1, kPolLast = 2 };
struct ClPol { unsigned long long keep, last; };
__device__ __forceinline__ ClPol cl_policies() {
    ClPol p;
#if MGA_CL_HINTS
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p.keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p.last));
#else
    p.keep = 0; p.last = 0;
#endif
    return p;
}
template <int POL>
__device__ __forceinline__ uint4 ldg128(const void* p, const ClPol& pol) {
    uint4 t;
#if MGA_CL_HINTS
    if constexpr (POL == kPolNone) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p));
    } else {
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p), "l"(POL == kPolKeep ? pol.keep : pol.last));
    }
#else
    (void)pol;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p));
#endif
    return t;
}
// base + 32-bit byte offset in ONE instruction (IMAD.WIDE.U32)
__device__ __forceinline__ const char* addw(const char* base, unsigned off) {
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(r) : "r"(off), "l"(base));
    return reinterpret_cast<const char*>(r);
}
__device__ __forceinline__ char* addw(char* base, unsigned off) {
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(r) : "r"(off), "l"(base));
    return reinterpret_cast<char*>(r);
}
template <typename T, int VEC>
__device__ __forceinline__ void unpack(const uint4& t, float (&v)[VEC]) {
    if constexpr (sizeof(T) == 4) {
        v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
    } else {
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
template <typename T, int VEC>
__device__ __forceinline__ uint4 pack(const float (&v)[VEC]) {
    uint4 t;
    if constexpr (sizeof(T) == 4) {
        t = make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
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
        t = make_uint4(w[0], w[1], w[2], w[3]);
    }
    return t;
}
// VEC consecutive fp32 / int32 values from shared memory (16-byte aligned)
template <int VEC>
__device__ __forceinline__ void lds_f(const float* p, float (&v)[VEC]) {
#pragma unroll
    for (int i = 0; i < VEC / 4; ++i) {
        const float4 t = reinterpret_cast<const float4*>(p)[i];
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}
template <int VEC>
__device__ __forceinline__ void lds_i(const int* p, int (&v)[VEC]) {
#pragma unroll
    for (int i = 0; i < VEC / 4; ++i) {
        const int4 t = reinterpret_cast<const int4*>(p)[i];
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}

// Asynchronous L2 prefetch of this CTA's 1/CS share of the sample's CONTIGUOUS bytes (the cluster needs the whole sample):
// a few bulk-prefetch instructions issued by one thread, executed by the copy engine -- no registers, no scoreboard, no
// back-pressure on the compute warps.  The HBM -> L2 stream then runs at full speed underneath the latency-bound prologue.
__device__ __forceinline__ void cl_prefetch_share(const void* sample, size_t sample_bytes, int r, int CS) {
    if (threadIdx.x != 0) return;
    size_t share = (sample_bytes / (size_t)CS + 15) & ~(size_t)15;
    size_t lo = (size_t)r * share;
    if (lo >= sample_bytes) return;
    size_t hi = lo + share < sample_bytes ? lo + share : (sample_bytes & ~(size_t)15);
    const char* base = static_cast<const char*>(sample);
    constexpr size_t kChunk = 32768;
    for (size_t o = lo; o < hi; o += kChunk) {
        const unsigned n = (unsigned)(hi - o < kChunk ? hi - o : kChunk);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + o), "r"(n) : "memory");
    }
}

// n / d for small non-negative n (< 2^21) with inv = 1.0f / d precomputed: three instructions instead of ~20 for an integer division
// by a run-time value (the strip / pixel -> (row, column) maps below sit in loops every thread runs for every item)
__device__ __forceinline__ int fdiv_small(int n, float inv) { return __float2int_rz(((float)n + 0.5f) * inv); }

template <int NT>
__device__ __forceinline__ void cl_load_weights7(const float* __restrict__ wsam, int k, bool flip, float* w /* [3][49] */) {
    const int off = (kMaxK - k) / 2;
    for (int t = threadIdx.x; t < 3 * kMaxK * kMaxK; t += NT) {
        const int pl = t / (kMaxK * kMaxK);
        int ii = (t / kMaxK) % kMaxK, jj = t % kMaxK;
        if (flip) { ii = kMaxK - 1 - ii; jj = kMaxK - 1 - jj; }
        ii -= off; jj -= off;
        w[t] = (ii >= 0 && ii < k && jj >= 0 && jj < k) ? __ldg(wsam + (pl * k + ii) * k + jj) : 0.0f;
    }
}
// the three planes [pmax, pavg, m] of rows [y_lo, y_lo + rows): float4 chunks, zero outside the image / in the pad columns
template <int NT>
__device__ __forceinline__ void cl_stage_three(float* tile, int planeT, const float* const (&planes)[3], int y_lo, int rows, int H, int W, int TWp) {
    const int cpr = TWp / 4, total = rows * cpr;
    const float icpr = 1.0f / (float)cpr;
    for (int i = threadIdx.x; i < total; i += NT) {
        const int r = fdiv_small(i, icpr), c = i - r * cpr, yy = y_lo + r;
        const bool in = yy >= 0 && yy < H && c >= 1 && c < cpr - 1;
        const int o = yy * W + (c - 1) * 4;
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in && planes[pl] != nullptr) v = __ldcg(reinterpret_cast<const float4*>(planes[pl] + o));
            reinterpret_cast<float4*>(tile + pl * planeT)[i] = v;
        }
    }
}

// position of channel c in the group-major constant arrays: thread group g walks g, g+G, g+2G ... -> consecutive floats
__device__ __forceinline__ int cl_gm(int c, int G, int CG) { const int j = c / G; return (c - j * G) * CG + j; }

// sum the LPT staged partials of every channel in a fixed order: dst[c] = sum_j stage[c*(LPT|1) + j]
template <int NT>
__device__ __forceinline__ void cl_stage_sum(const float* stage, float* dst, int C, int LPT) {
    const int LPTp = LPT | 1;
    const int tpc = C * 2 > NT ? 1 : (C * 4 > NT ? 2 : (C * 8 > NT ? 4 : 8));  // threads per channel (power of two)
    const int sl = threadIdx.x & (tpc - 1);
    for (int c0 = 0; c0 < C; c0 += NT / tpc) {
        const int c = c0 + (int)threadIdx.x / tpc;
        float acc = 0.0f;
        if (c < C)
            for (int j = sl; j < LPT; j += tpc) acc += stage[c * LPTp + j];
        for (int of = tpc >> 1; of > 0; of >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, of);
        if (c < C && sl == 0) dst[c] = acc;
    }
}

// fetch the halo rows of `nplanes` tile planes from the CTAs that own them (DSMEM); own rows and padding stay untouched
template <int NT>
__device__ __forceinline__ void cl_fetch_halo(cgx::cluster_group& cluster, float* tile, int nplanes, const ClGeom& gm, int y0, int rows, int H,
                                              int W) {
    const int cpr = W / 4;  // float4 chunks per image row
    const int per = 6 * cpr;
    const float iper = 1.0f / (float)per, icpr = 1.0f / (float)cpr;
    for (int i = threadIdx.x; i < nplanes * per; i += NT) {
        const int pl = fdiv_small(i, iper), rem = i - pl * per;
        const int hr = fdiv_small(rem, icpr), ch = rem - hr * cpr;       // halo row 0..5: 3 above, 3 below
        const int tr = hr < 3 ? hr : rows + hr;              // tile row (own rows are 3 .. 3+rows-1)
        const int yy = y0 - 3 + tr;
        if (yy < 0 || yy >= H) continue;
        const int rr = yy / gm.rowsPer;
        const int rtr = yy - rr * gm.rowsPer + 3;
        const float* rt = cluster.map_shared_rank(tile, rr);
        const float4 v = *reinterpret_cast<const float4*>(rt + (size_t)pl * gm.planeT + rtr * gm.TWp + 4 + ch * 4);
        *reinterpret_cast<float4*>(tile + (size_t)pl * gm.planeT + tr * gm.TWp + 4 + ch * 4) = v;
    }
}

// ================================================================== forward
template <typename T>
__global__ void __launch_bounds__(kClNTF, MGA_CL_MINB_F) cl_fwd_kernel(const T* __restrict__ x, const void* __restrict__ mask, int mdt,
                                                                  T* __restrict__ out, Shape sh, mga_cbam_params prm, Ctx ctx, const ClFwdOff o,
                                                                  ClGeom gm) {
    constexpr int VEC = 16 / sizeof(T);
    constexpr int K = kClKF;
    constexpr int NT = kClNTF, NW = NT / 32;
    constexpr int kGrp = NT / 3;  // conv: threads per input plane
    extern __shared__ __align__(128) float clsm[];
    float* const csm = clsm;
    cgx::cluster_group cluster = cgx::this_cluster();
    const int CS = gm.CS;
    const int r = (int)cluster.block_rank();
    const int b = blockIdx.x / CS;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int C = sh.C, S = sh.S, Hd = sh.hidden, W = sh.W, H = sh.H;
    const float invW = 1.0f / (float)W;
    const int y0 = r * gm.rowsPer;
    const int rows = max(0, min(gm.rowsPer, H - y0));
    const int nP = rows * W, nU = nP / VEC, p0 = y0 * W;
    const int nPmax = gm.nPmax, nUmax = gm.nUmax, TWp = gm.TWp, planeT = gm.planeT;
    const int G = gm.G, CG = gm.CG;
    const bool has_mask = sh.has_mask();
    const bool save = !sh.no_save();  // inference: the planes only the backward reads are not written
    // (the shared-memory layout `o` is computed on the host: per thread it cost ~1000 instructions per warp at kernel start)
    float* wk = csm + o.wk;
    float* red = csm + o.red;
    float* s_avg = csm + o.avg;
    float* s_mx = csm + o.mx;
    float* s_ha = csm + o.ha;
    float* s_hm = csm + o.hm;
    float2* sAB = reinterpret_cast<float2*>(csm + o.sAB);
    float* s_q = csm + o.q;
    float* part = csm + o.part;
    float* msum_part = csm + o.msum;
    float* mloc = csm + o.mloc;
    float* aloc = csm + o.aloc;
    float* tile = csm + o.tile;  // [pmax | pavg | m][tileRows][TWp]
    float* mg = csm + o.mg;
    float* stage = csm + o.stage;
    const ClPol pol = cl_policies();
    const unsigned rowB = (unsigned)S * (unsigned)sizeof(T);  // bytes per channel plane

    const size_t bS = (size_t)b * S;
    // ---- phase 0: conv weights, MLP weights (async), first pooling loads in flight, mask -> m (own pixels), sum(m)
    stamp(0);
    cluster.barrier_arrive();  // #0 (waited before the first push): every CTA of the cluster has started
    if (gm.prefetch) cl_prefetch_share(x + (size_t)b * C * S, (size_t)C * rowB, r, CS);
    const float* w1p = prm.w1;
    const float* w2p = prm.w2;
    if (o.mlpw >= 0 && ((reinterpret_cast<uintptr_t>(prm.w1) | reinterpret_cast<uintptr_t>(prm.w2)) & 15) == 0) {  // 16-byte asynchronous copies: no registers, waited for just before the MLP
        float* mw = csm + o.mlpw;
        const int n4 = C * Hd / 4;
        for (int i = tid; i < n4; i += NT) {
            cp_async16(mw + 4 * i, prm.w1 + 4 * i);
            cp_async16(mw + C * Hd + 4 * i, prm.w2 + 4 * i);
        }
        cp_async_commit();
        w1p = mw;
        w2p = mw + C * Hd;
    }
    const int LPT = gm.LPT, LPTp = LPT | 1, slots = gm.slots;
    const int slot = tid / LPT, ul = tid - slot * LPT;
    const char* xbytes = reinterpret_cast<const char*>(x + (size_t)b * C * S + p0);  // the CTA's pixel range of channel 0
    unsigned offk[K];
    int nk = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int u = ul + k * LPT;
        offk[k] = (unsigned)u * 16u;
        nk += (slot < slots && u < nU) ? 1 : 0;
    }
    uint4 raw[K];  // the first channel of this thread's walk: in flight while the mask is prepared
    if (slot < C) {
        const char* row = xbytes + (size_t)slot * rowB;
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (k < nk) raw[k] = ldg128<kPolKeep>(addw(row, offk[k]), pol);
    }
    for (int i = nP + tid; i < gm.LPT * K * VEC; i += NT) mloc[i] = 0.0f;  // padding of the T1 footprint
    cl_load_weights7<NT>(prm.wsam, sh.k, false, wk);
    float mtot = 0.0f;
    {
        float macc = 0.0f;
        for (int i = tid; i < nP; i += NT) {
            float v = 1.0f;
            if (has_mask) {
                v = load_mask_any(mask, mdt, bS + p0 + i);
                if (sh.gate_clamp()) v = fminf(fmaxf(v, 0.0f), 1.0f);
                if (sh.sigmoid_mask()) v = sigmoidf_acc(v);
                if (save) ctx.m[bS + p0 + i] = v;
                macc += v;
            }
            mloc[i] = v;
        }
        mtot = block_sum(macc, red);
    }
    __syncthreads();

    // ---- phase 1 (T1): per channel over the CTA's pixels: sum x*m, sum x, masked max + arg max
    stamp(1);
    if (nk > 0) {
        for (int c = slot; c < C; c += slots) {
            if (c != slot) {
                const char* row = xbytes + (size_t)c * rowB;
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (k < nk) raw[k] = ldg128<kPolKeep>(addw(row, offk[k]), pol);
            }
            float sxm = 0.0f, sx = 0.0f, best = -INFINITY;
            int bpos = -1;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (k < nk) {
                    float v[VEC], mv[VEC];
                    unpack<T, VEC>(raw[k], v);
                    lds_f<VEC>(mloc + (ul + k * LPT) * VEC, mv);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        sx += v[i];
                        sxm = fmaf(v[i], mv[i], sxm);
                        if (mv[i] > 0.5f && v[i] > best) { best = v[i]; bpos = k * VEC + i; }  // scan order = pixel order within the thread
                    }
                }
            }
            int bidx = -1;
            if (bpos >= 0) { const int kk = bpos / VEC; bidx = p0 + (ul + kk * LPT) * VEC + (bpos - kk * VEC); }
            float* st = stage + c * LPTp + ul;
            st[0] = sxm;
            st[C * LPTp] = sx;
            st[2 * C * LPTp] = best;
            st[3 * C * LPTp] = __int_as_float(bidx);
        }
    } else if (slot < slots) {
        for (int c = slot; c < C; c += slots) {  // this thread owns no valid unit (short last rank): neutral partials
            float* st = stage + c * LPTp + ul;
            st[0] = 0.0f;
            st[C * LPTp] = 0.0f;
            st[2 * C * LPTp] = -INFINITY;
            st[3 * C * LPTp] = __int_as_float(-1);
        }
    }
    __syncthreads();
    // per channel: combine the LPT thread partials (4 threads per channel, fixed order; max ties -> lower pixel index), push to every rank
    cluster.barrier_wait();  // #0
    if (has_mask && tid < CS) cluster.map_shared_rank(msum_part, tid)[r] = mtot;
    for (int c0 = 0; c0 < C; c0 += NT / 4) {
        const int c = c0 + tid / 4, sl = tid & 3;
        float sxm = 0.0f, sx = 0.0f, best = -INFINITY;
        int bidx = -1;
        if (c < C) {
            const float* st = stage + c * LPTp;
            for (int j = sl; j < LPT; j += 4) {
                sxm += st[j];
                sx += st[C * LPTp + j];
                const float ob = st[2 * C * LPTp + j];
                const int oi = __float_as_int(st[3 * C * LPTp + j]);
                if ((oi >= 0) && (bidx < 0 || ob > best || (ob == best && oi < bidx))) { best = ob; bidx = oi; }
            }
        }
#pragma unroll
        for (int of = 2; of > 0; of >>= 1) {
            sxm += __shfl_xor_sync(0xffffffffu, sxm, of);
            sx += __shfl_xor_sync(0xffffffffu, sx, of);
            const float ob = __shfl_xor_sync(0xffffffffu, best, of);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, of);
            if ((oi >= 0) && (bidx < 0 || ob > best || (ob == best && oi < bidx))) { best = ob; bidx = oi; }
        }
        if (c < C) {
            const float val = sl == 0 ? sxm : (sl == 1 ? sx : (sl == 2 ? best : __int_as_float(bidx)));
            for (int rr = 0; rr < CS; ++rr) cluster.map_shared_rank(part, rr)[r * 4 * C + sl * C + c] = val;  // lane sl pushes quantity sl
        }
    }
    const float beta = __ldg(prm.beta);
    stamp(2);
    cluster.sync();  // #1: partials of all ranks are in local shared memory

    // ---- phase 2: pooled descriptors, shared MLP, sigmoid (recomputed by every CTA)
    stamp(3);
    // the staging area is dead: build the conv tile in it (zero padding; own rows of the m plane)
    for (int i = tid; i < 3 * planeT; i += NT) tile[i] = 0.0f;
    __syncthreads();
    if (has_mask)
        for (int i = tid; i < nP; i += NT) {
            const int ry = fdiv_small(i, invW), cx = i - ry * W;
            tile[2 * planeT + (3 + ry) * TWp + 4 + cx] = mloc[i];
        }
    {
        float use = 0.0f, den = 1.0f;
        if (has_mask) {
            float tot = 0.0f;
            for (int rr = 0; rr < CS; ++rr) tot += msum_part[rr];
            use = (tot / (float)S >= sh.tiny_thr) ? 1.0f : 0.0f;
            den = fmaxf(tot, sh.eps);
            if (r == 0 && tid == 0) { ctx.msum[b] = tot; ctx.use[b] = use; ctx.den[b] = den; }
        }
        const float invS = 1.0f / (float)S;
        for (int c = tid; c < C; c += NT) {
            float sxm = 0.0f, sx = 0.0f, best = -INFINITY;
            int bidx = -1;
            for (int rr = 0; rr < CS; ++rr) {  // ranks own increasing pixel ranges: strict > keeps the first maximum
                const float* pp = part + rr * 4 * C;
                sxm += pp[c];
                sx += pp[C + c];
                const float ob = pp[2 * C + c];
                const int oi = __float_as_int(pp[3 * C + c]);
                if (oi >= 0 && (bidx < 0 || ob > best)) { best = ob; bidx = oi; }
            }
            const float Gm = sx * invS;
            const float A = has_mask ? sxm / den : Gm;
            const float avg = has_mask ? (A * use + Gm * (1.0f - use)) : Gm;
            const bool dead = bidx < 0;  // no pixel with m > 0.5 (masked_cbam.py:118-121)
            const float mx = dead ? Gm : best;
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
    if (b == 0 && r == 0 && tid == 0) { ctx.consts[0] = k0; ctx.consts[1] = k1; ctx.consts[2] = alpha; ctx.consts[3] = sigmoidf_acc(beta); }
    cp_async_wait_all();  // the staged MLP weights of this thread ...
    __syncthreads();      // ... and of every other thread
    for (int j = w; j < Hd; j += NW) {
        const float* wr = w1p + (size_t)j * C;
        float pa = 0.0f, pm = 0.0f;
#pragma unroll 4
        for (int c = lane; c < C; c += 32) {
            const float wv = wr[c];
            pa = fmaf(wv, s_avg[c], pa);
            pm = fmaf(wv, s_mx[c], pm);
        }
        pa = warp_sum(pa);
        pm = warp_sum(pm);
        if (lane == 0) {
            const float bb = __ldg(prm.b1 + j);
            const float ha = fmaxf(pa + bb, 0.0f), hm = fmaxf(pm + bb, 0.0f);
            s_ha[j] = ha;
            s_hm[j] = hm;
            if (r == 0) { ctx.ha[b * Hd + j] = ha; ctx.hm[b * Hd + j] = hm; }
        }
    }
    __syncthreads();
    {
        const bool add = sh.samcam_add();
        for (int c = tid; c < C; c += NT) {
            const float* wr = w2p + (size_t)c * Hd;
            float za = 0.0f, zm = 0.0f;
#pragma unroll 4
            for (int j = 0; j < Hd; ++j) {
                const float wv = wr[j];
                za = fmaf(wv, s_ha[j], za);
                zm = fmaf(wv, s_hm[j], zm);
            }
            const float bb = __ldg(prm.b2 + c);
            const float s = sigmoidf_acc((za + bb) + (zm + bb));  // b2 enters twice (masked_cbam.py:128)
            if (r == 0) ctx.s[b * C + c] = s;
            const int gi = cl_gm(c, G, CG);
            s_q[gi] = add ? 1.0f : s;
            // out = x * (A * a + B):  multiply: k0 + k1*s*a ;  add: k0 + k1*(s + a)
            sAB[gi] = add ? make_float2(k1, fmaf(k1, s, k0)) : make_float2(k1 * s, k0);
        }
    }
    __syncthreads();

    // ---- phase 3 (T2): per pixel max / arg max / sum over channels of x*q
    stamp(4);
    const int g2 = tid / nUmax, ul2 = tid - g2 * nUmax;
    const bool act2 = g2 < G && ul2 < nU;
    const int nj = act2 ? (C - g2 + G - 1) / G : 0;  // channels of this thread's group
    constexpr int KB = (VEC == 4) ? 8 : 4;
    const unsigned gstep = (unsigned)G * rowB;       // bytes between two channels of a group
    {
        float vmax[VEC], vsum[VEC];
        int vidx[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) { vmax[i] = -INFINITY; vsum[i] = 0.0f; vidx[i] = -1; }
        if (act2) {
            const char* p = xbytes + (size_t)g2 * rowB + (size_t)ul2 * 16;
            const float* qp = s_q + g2 * CG;
            int j = 0;
            for (; j + KB <= nj; j += KB) {
                uint4 raw[KB];
#pragma unroll
                for (int kc = 0; kc < KB; ++kc) raw[kc] = ldg128<kPolKeep>(addw(p, kc * gstep), pol);
                float qv[KB];
                lds_f<KB>(qp + j, qv);
#pragma unroll
                for (int kc = 0; kc < KB; ++kc) {
                    float v[VEC];
                    unpack<T, VEC>(raw[kc], v);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const float y = v[i] * qv[kc];
                        vsum[i] += y;
                        if (y > vmax[i]) { vmax[i] = y; vidx[i] = j + kc; }
                    }
                }
                p += (size_t)KB * gstep;
            }
            for (; j < nj; ++j) {
                float v[VEC];
                unpack<T, VEC>(ldg128<kPolKeep>(p, pol), v);
                const float q = qp[j];
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const float y = v[i] * q;
                    vsum[i] += y;
                    if (y > vmax[i]) { vmax[i] = y; vidx[i] = j; }
                }
                p += gstep;
            }
        }
        if (g2 < G && ul2 < nUmax) {
            float* d = mg + (size_t)g2 * nPmax + ul2 * VEC;
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                d[i] = vmax[i];
                d[(size_t)G * nPmax + i] = vsum[i];
                d[(size_t)2 * G * nPmax + i] = __int_as_float(vidx[i] < 0 ? 0x7fffffff : g2 + vidx[i] * G);  // channel index
            }
        }
    }
    __syncthreads();
    {
        const float invC = 1.0f / (float)C;
        for (int p = tid; p < nP; p += NT) {
            float bm = mg[p], bsum = mg[(size_t)G * nPmax + p];
            int bi = __float_as_int(mg[(size_t)2 * G * nPmax + p]);
            for (int j = 1; j < G; ++j) {
                const float om = mg[(size_t)j * nPmax + p];
                const int oi = __float_as_int(mg[(size_t)(2 * G + j) * nPmax + p]);
                bsum += mg[(size_t)(G + j) * nPmax + p];
                if (om > bm || (om == bm && oi < bi)) { bm = om; bi = oi; }  // torch.max: first maximal channel
            }
            const float pavg = bsum * invC;
            const int ry = fdiv_small(p, invW), cx = p - ry * W;
            tile[(3 + ry) * TWp + 4 + cx] = bm;
            tile[planeT + (3 + ry) * TWp + 4 + cx] = pavg;
            if (save) {
                ctx.pmax[bS + p0 + p] = bm;
                ctx.pavg[bS + p0 + p] = pavg;
                ctx.idx[bS + p0 + p] = bi;
            }
        }
    }
    stamp(5);
    cluster.sync();  // #2: every rank's own rows of the three planes are in its tile

    // ---- phase 4: halo rows through DSMEM, then a = sigmoid(conv7x7([pmax, pavg, m])) for the own rows
    stamp(6);
    if (nP > 0) cl_fetch_halo<NT>(cluster, tile, has_mask ? 3 : 2, gm, y0, rows, H, W);
    cluster.barrier_arrive();  // #3 (waited at the end): this CTA reads no remote shared memory any more
    __syncthreads();
    {
        float* cpart = mg;  // [3][nPmax]
        const int grp = tid / kGrp, gl = tid - grp * kGrp;
        const int spr = W / 4, nStrips = rows * spr;
        const float ispr = 1.0f / (float)spr;
        if (grp < 3) {
            const float* wv = wk + grp * kMaxK * kMaxK;
            for (int s = gl; s < nStrips; s += kGrp) {
                const int ry = fdiv_small(s, ispr), x0 = (s - ry * spr) * 4;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                strip_conv7(tile + grp * planeT + ry * TWp + x0, TWp, wv, acc);
                *reinterpret_cast<float4*>(cpart + (size_t)grp * nPmax + s * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            }
        }
        __syncthreads();
        for (int p = tid; p < nP; p += NT) {
            const float v = (cpart[p] + cpart[nPmax + p]) + cpart[2 * nPmax + p];
            const float a = sigmoidf_acc(v);
            aloc[p] = a;
            ctx.a[bS + p0 + p] = a;
        }
    }
    __syncthreads();

    // ---- phase 5 (T2): out = x * (A*a + B)
    stamp(7);
    if (act2 && !sh.gates_only()) {  // (gates-only calls stop here: s and a are in the context)
        float av[VEC];
        lds_f<VEC>(aloc + ul2 * VEC, av);
        const char* p = xbytes + (size_t)g2 * rowB + (size_t)ul2 * 16;
        char* q = reinterpret_cast<char*>(out + (size_t)b * C * S + p0) + (size_t)g2 * rowB + (size_t)ul2 * 16;
        const float2* ab = sAB + g2 * CG;
        int j = 0;
        for (; j + KB <= nj; j += KB) {
            uint4 raw[KB];
#pragma unroll
            for (int kc = 0; kc < KB; ++kc) raw[kc] = ldg128<kPolLast>(addw(p, kc * gstep), pol);
            float abv[2 * KB];
            lds_f<2 * KB>(reinterpret_cast<const float*>(ab + j), abv);
#pragma unroll
            for (int kc = 0; kc < KB; ++kc) {
                float v[VEC];
                unpack<T, VEC>(raw[kc], v);
#pragma unroll
                for (int i = 0; i < VEC; ++i) v[i] *= fmaf(abv[2 * kc], av[i], abv[2 * kc + 1]);
                __stcs(reinterpret_cast<uint4*>(addw(q, kc * gstep)), pack<T, VEC>(v));
            }
            p += (size_t)KB * gstep;
            q += (size_t)KB * gstep;
        }
        for (; j < nj; ++j) {
            float v[VEC];
            unpack<T, VEC>(ldg128<kPolLast>(p, pol), v);
            const float2 c2 = ab[j];
#pragma unroll
            for (int i = 0; i < VEC; ++i) v[i] *= fmaf(c2.x, av[i], c2.y);
            __stcs(reinterpret_cast<uint4*>(q), pack<T, VEC>(v));
            p += gstep;
            q += gstep;
        }
    }
    stamp(8);
    cluster.barrier_wait();  // #3
}

// ================================================================== backward
template <typename T>
__global__ void __launch_bounds__(kClNTB, MGA_CL_MINB_B) cl_bwd_kernel(const T* __restrict__ x, const T* __restrict__ g, const void* __restrict__ mask,
                                                                  int mdt, T* __restrict__ dx, void* __restrict__ dmask, Shape sh,
                                                                  mga_cbam_params prm, Ctx ctx, BwdScratch bs, const __grid_constant__ PlaneMaps maps,
                                                                  const ClBwdOff o, ClGeom gm) {
    constexpr int VEC = 16 / sizeof(T);
    constexpr int K = (VEC == 4) ? kClKB : kClKB16;
    constexpr int NT = kClNTB, NW = NT / 32;
    constexpr int kGrp = NT / 3;  // conv: threads per input plane
    constexpr int kStride = 3 * kMaxK * kMaxK + 1;
    extern __shared__ __align__(128) float clsm[];
    float* const csm = clsm;
    cgx::cluster_group cluster = cgx::this_cluster();
    const int CS = gm.CS;
    const int r = (int)cluster.block_rank();
    const int b = blockIdx.x / CS;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int C = sh.C, S = sh.S, Hd = sh.hidden, W = sh.W, H = sh.H;
    const float invW = 1.0f / (float)W;
    const int y0 = r * gm.rowsPer;
    const int rows = max(0, min(gm.rowsPer, H - y0));
    const int nP = rows * W, nU = nP / VEC, p0 = y0 * W;
    const int nPmax = gm.nPmax, nUmax = gm.nUmax, TWp = gm.TWp, planeT = gm.planeT;
    const int G = gm.G, CG = gm.CG;
    const bool has_mask = sh.has_mask();
    const bool multiply = !sh.samcam_add();
    // (shared-memory layout `o`: computed on the host)
    float* wsm = csm + o.wsm;
    double* redd = reinterpret_cast<double*>(csm + o.red);
    float* s_s = csm + o.s;
    float4* chA = reinterpret_cast<float4*>(csm + o.chA);
    float* chG = csm + o.chG;
    int* amx = reinterpret_cast<int*>(csm + o.amx);
    float* s_cM = csm + o.cM;
    float* s_dz = csm + o.dz;
    float* e_loc = csm + o.eloc;
    float* q_loc = csm + o.qloc;
    float* bin_loc = csm + o.binloc;
    float* s_dha = csm + o.dha;
    float* s_dhm = csm + o.dhm;
    float* s_q = csm + o.q;
    float* epart = csm + o.epart;
    float* qpart = csm + o.qpart;
    float* binpart = csm + o.binpart;
    double* gxpart = reinterpret_cast<double*>(csm + o.gxpart);
    float* aloc = csm + o.aloc;
    float* ae = csm + o.ae;
    float* mloc = csm + o.mloc;
    int* idxl = reinterpret_cast<int*>(csm + o.idx);
    float* pmx = csm + o.pmx;
    float* d0 = csm + o.d0;
    float* d1s = csm + o.d1s;
    float* d2 = csm + o.d2;
    float* dpre = csm + o.dpre;
    float* cat = csm + o.cat;
    float* tp = csm + o.tp;
    float* dwp = csm + o.dwp;
    float* binw = csm + o.binw;
    float* stage = csm + o.stage;
    const ClPol pol = cl_policies();
    const float k0 = ctx.consts[0], k1 = ctx.consts[1];
    const unsigned rowB = (unsigned)S * (unsigned)sizeof(T);

    const size_t bS = (size_t)b * S;
    // ---- phase 0: first (x,g) loads in flight, MLP weights (async), per-sample vectors and planes of the saved context
    stamp(0);
    if (gm.prefetch) {
        cl_prefetch_share(x + (size_t)b * C * S, (size_t)C * rowB, r, CS);
        cl_prefetch_share(g + (size_t)b * C * S, (size_t)C * rowB, r, CS);
    }
    const int LPT = gm.LPT, LPTp = LPT | 1, slots = gm.slots;
    const int slot = tid / LPT, ul = tid - slot * LPT;
    const size_t sbase = (size_t)b * C * S + p0;
    const char* xbytes = reinterpret_cast<const char*>(x + sbase);
    const char* gbytes = reinterpret_cast<const char*>(g + sbase);
    unsigned offk[K];
    int nk = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int u = ul + k * LPT;
        offk[k] = (unsigned)u * 16u;
        nk += (slot < slots && u < nU) ? 1 : 0;
    }
    uint4 xr[K], gr[K];  // the first channel of this thread's walk: in flight while the context is staged
    if (slot < C) {
        const char* xrow = xbytes + (size_t)slot * rowB;
        const char* grow = gbytes + (size_t)slot * rowB;
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (k < nk) {
                xr[k] = ldg128<kPolKeep>(addw(xrow, offk[k]), pol);
                gr[k] = ldg128<MGA_CL_GPOL>(addw(grow, offk[k]), pol);
            }
    }
    const float* w1p = prm.w1;
    const float* w2p = prm.w2;
    if (o.mlpw >= 0 && ((reinterpret_cast<uintptr_t>(prm.w1) | reinterpret_cast<uintptr_t>(prm.w2)) & 15) == 0) {
        float* mw = csm + o.mlpw;
        const int n4 = C * Hd / 4;
        for (int i = tid; i < n4; i += NT) {
            cp_async16(mw + 4 * i, prm.w1 + 4 * i);
            cp_async16(mw + C * Hd + 4 * i, prm.w2 + 4 * i);
        }
        cp_async_commit();
        w1p = mw;
        w2p = mw + C * Hd;
    }
    cl_load_weights7<NT>(prm.wsam, sh.k, true, wsm);
    for (int i = tid; i < planeT; i += NT) dpre[i] = 0.0f;
    for (int i = tid; i < NW * C; i += NT) binw[i] = 0.0f;
    for (int c = tid; c < C; c += NT) {
        const float s = ctx.s[b * C + c];
        s_s[c] = s;
        s_q[c] = multiply ? s : 1.0f;
    }
    // Saved planes of the CTA's own pixels (a, idx, pmax [, m]): contiguous nP * 4 bytes each -> asynchronous bulk copies
    // (cp.async.bulk, SASS UBLKCP) issued by one thread and counted on an mbarrier; no registers, no scoreboard.  In multiply
    // mode the zero-padded T1 copy `ae` of a is a second bulk copy; the padding and the constant planes are plain stores.
    uint64_t* const bars = reinterpret_cast<uint64_t*>(csm + o.bar);
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
        if (nP > 0) {
            const uint32_t nb = (uint32_t)nP * 4u;
            mbar_expect_tx(&bars[0], nb * (3u + (has_mask ? 1u : 0u) + (multiply ? 1u : 0u)));
            bulk_g2s(aloc, ctx.a + bS + p0, nb, &bars[0]);
            bulk_g2s(idxl, ctx.idx + bS + p0, nb, &bars[0]);
            bulk_g2s(pmx, ctx.pmax + bS + p0, nb, &bars[0]);
            if (has_mask) bulk_g2s(mloc, ctx.m + bS + p0, nb, &bars[0]);
            if (multiply) bulk_g2s(ae, ctx.a + bS + p0, nb, &bars[0]);
        }
    }
    {
        const int padP = gm.LPT * K * VEC;
        for (int p = tid; p < padP; p += NT) {
            d1s[p] = 0.0f;
            if (p >= nP) ae[p] = 0.0f;
            else if (!multiply) ae[p] = 1.0f;
        }
        if (!has_mask)
            for (int p = tid; p < nP; p += NT) mloc[p] = 0.0f;
    }
    __syncthreads();
    if (nP > 0) mbar_wait(&bars[0], 0);

    // ---- phase 1 (T1) over (x,g): T_p = sum_c g x q_c (per pixel), E_c = sum_p g x (a | 1) (per channel), sum g x
    stamp(1);
    double gxs = 0.0;  // sum g x over everything this thread touches: it enters d beta, which cancels heavily -> fp64 across channels
    {
        float tacc[K][VEC];
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int i = 0; i < VEC; ++i) tacc[k][i] = 0.0f;
        if (slot < slots) {
            for (int c = slot; c < C; c += slots) {
                if (c != slot) {
                    const char* xrow = xbytes + (size_t)c * rowB;
                    const char* grow = gbytes + (size_t)c * rowB;
#pragma unroll
                    for (int k = 0; k < K; ++k)
                        if (k < nk) {
                            xr[k] = ldg128<kPolKeep>(addw(xrow, offk[k]), pol);
                            gr[k] = ldg128<MGA_CL_GPOL>(addw(grow, offk[k]), pol);
                        }
                }
                const float q = s_q[c];
                float e = 0.0f, gxc = 0.0f;  // (sum g x: per channel first, then across channels -- two short fp32 chains instead of one long one)
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (k < nk) {
                        float xv[VEC], gv[VEC], av[VEC];
                        unpack<T, VEC>(xr[k], xv);
                        unpack<T, VEC>(gr[k], gv);
                        lds_f<VEC>(ae + (ul + k * LPT) * VEC, av);
#pragma unroll
                        for (int i = 0; i < VEC; ++i) {
                            const float gx = gv[i] * xv[i];
                            tacc[k][i] = fmaf(gx, q, tacc[k][i]);
                            e = fmaf(gx, av[i], e);
                            gxc += gx;
                        }
                    }
                }
                gxs += (double)gxc;
                stage[c * LPTp + ul] = e;
            }
            // per-pixel partials of this channel slot
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (k < nk) {
                    float* d = tp + (size_t)slot * nPmax + (ul + k * LPT) * VEC;
#pragma unroll
                    for (int i = 0; i < VEC; ++i) d[i] = tacc[k][i];
                }
            }
        }
    }
    __syncthreads();
    cl_stage_sum<NT>(stage, e_loc, C, LPT);
    double at_acc = 0.0;
    for (int p = tid; p < nP; p += NT) {
        float t = 0.0f;
        for (int j = 0; j < slots; ++j) t += tp[(size_t)j * nPmax + p];
        const float a = aloc[p];
        const int ry = fdiv_small(p, invW), cx = p - ry * W;
        dpre[(3 + ry) * TWp + 4 + cx] = k1 * t * a * (1.0f - a);
        at_acc += (double)a * (double)t;
    }
    stamp(2);
    cluster.sync();  // #1: every rank's own rows of dpre are in its tile

    // ---- phase 2: dpre halo through DSMEM; dcat = conv7x7^T(dpre); dWsam partials
    stamp(3);
    // the T partials are dead (every thread of the CTA is past the merge): their storage takes the padded [pmax, pavg, m]
    // tile of rows [y0-3, y0+rowsPer+3) x columns [-4, W+4) -- one tensor-map box per plane (SASS UTMALDG), out-of-image
    // elements zero-filled by the TMA unit (= the convolution's padding); it lands while the dpre halo / conv^T run.
    const bool tile_tma = gm.use_tma && rows > 0;
    if (tile_tma) {
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes of the old tenant before async-proxy writes
            const int np = has_mask ? 3 : 2;
            mbar_expect_tx(&bars[1], (uint32_t)(np * gm.tileRows * TWp * 4));
            for (int pl = 0; pl < np; ++pl) tma_load_3d(cat + pl * planeT, &maps.m[pl], -4, y0 - kMaxK / 2, b, &bars[1]);
        }
        if (!has_mask)
            for (int i = tid; i < planeT; i += NT) cat[2 * planeT + i] = 0.0f;
    } else if (rows > 0) {
        const float* const planes[3] = {ctx.pmax + bS, ctx.pavg + bS, has_mask ? ctx.m + bS : nullptr};
        cl_stage_three<NT>(cat, planeT, planes, y0 - kMaxK / 2, gm.tileRows, H, W, TWp);
    }
    if (nP > 0) cl_fetch_halo<NT>(cluster, dpre, 1, gm, y0, rows, H, W);
    __syncthreads();
    {
        const int grp = tid / kGrp, gl = tid - grp * kGrp;
        const int spr = W / 4, nStrips = rows * spr;
        const float ispr = 1.0f / (float)spr;
        const float invC = 1.0f / (float)C;
        if (grp < 3) {
            const float* wv = wsm + grp * kMaxK * kMaxK;
            float* dst = grp == 0 ? d0 : (grp == 1 ? d1s : d2);
            const float sc = grp == 1 ? invC : 1.0f;
            for (int s = gl; s < nStrips; s += kGrp) {
                const int ry = fdiv_small(s, ispr), x0 = (s - ry * spr) * 4;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                strip_conv7(dpre + ry * TWp + x0, TWp, wv, acc);
                *reinterpret_cast<float4*>(dst + s * 4) = make_float4(acc[0] * sc, acc[1] * sc, acc[2] * sc, acc[3] * sc);
            }
        }
        if (tile_tma) mbar_wait(&bars[1], 0);
        // conv2d_weight: a team of 12 threads owns one (plane, kernel row) pair, the 7 column taps live in registers
        constexpr int kTeam = 12;
        const int team = tid / kTeam, tl = tid - team * kTeam;
        if (team < 3 * kMaxK) {
            const int pl = team / kMaxK, i7 = team - pl * kMaxK;
            float acc[kMaxK] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int s = tl; s < nStrips; s += kTeam) {
                const int ry = fdiv_small(s, ispr), x0 = (s - ry * spr) * 4;
                const float4 d = *reinterpret_cast<const float4*>(dpre + (ry + kMaxK / 2) * TWp + x0 + 4);
                const float* cr = cat + pl * planeT + (ry + i7) * TWp + x0;
                const float4 a4 = *reinterpret_cast<const float4*>(cr);
                const float4 b4 = *reinterpret_cast<const float4*>(cr + 4);
                const float4 c4 = *reinterpret_cast<const float4*>(cr + 8);
                const float win[12] = {a4.x, a4.y, a4.z, a4.w, b4.x, b4.y, b4.z, b4.w, c4.x, c4.y, c4.z, c4.w};
                const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                for (int j = 0; j < kMaxK; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[j] = fmaf(win[1 + i + j], dv[i], acc[j]);
            }
#pragma unroll
            for (int j = 0; j < kMaxK; ++j) dwp[(team * kTeam + tl) * kMaxK + j] = acc[j];
        }
        __syncthreads();
        const int cta = b * CS + r;
        const int k = sh.k, koff = (kMaxK - k) / 2;
        if (tid < 3 * kMaxK * kMaxK) {
            const int tm = tid / kMaxK, j7 = tid - tm * kMaxK;  // tm = pl * 7 + i7
            const int pl = tm / kMaxK, ii = tm - pl * kMaxK - koff, jj = j7 - koff;
            if (ii >= 0 && ii < k && jj >= 0 && jj < k) {
                float t = 0.0f;
#pragma unroll
                for (int q = 0; q < kTeam; ++q) t += dwp[(tm * kTeam + q) * kMaxK + j7];
                bs.convpart[(size_t)cta * kStride + (pl * k + ii) * k + jj] = t;
            }
        }
        const double at = block_sum_d(at_acc, redd);
        if (tid == 0) bs.atpart[cta] = at;
    }
    __syncthreads();

    // ---- phase 3 (T1, multiply mode) over x: Q_c = sum_p x * dcat1/C ; the arg-max term of ds goes through pmax:
    //      sum_{p: idx_p = c} dcat0_p x_cp s_c = sum_{p: idx_p = c} dcat0_p pmax_p   (binned per channel, deterministic)
    stamp(4);
    if (multiply) {
        if (slot < slots) {
            for (int c = slot; c < C; c += slots) {
                const char* xrow = xbytes + (size_t)c * rowB;
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (k < nk) xr[k] = ldg128<kPolKeep>(addw(xrow, offk[k]), pol);
                float qv = 0.0f;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (k < nk) {
                        float xv[VEC], wv[VEC];
                        unpack<T, VEC>(xr[k], xv);
                        lds_f<VEC>(d1s + (ul + k * LPT) * VEC, wv);
#pragma unroll
                        for (int i = 0; i < VEC; ++i) qv = fmaf(xv[i], wv[i], qv);
                    }
                }
                stage[c * LPTp + ul] = qv;
            }
        }
        // warp-private bins: lanes with the same arg-max channel are summed in lane order by their lowest lane
        for (int base = w * 32; base < nP; base += NT) {
            const int p = base + lane;
            const bool ok = p < nP;
            const int key = ok ? idxl[p] : -1 - lane;
            const float v = ok ? d0[p] * pmx[p] : 0.0f;
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            float acc = 0.0f;
#pragma unroll 8
            for (int l = 0; l < 32; ++l) {
                const float t = __shfl_sync(0xffffffffu, v, l);
                if ((peers >> l) & 1u) acc += t;
            }
            if (ok && lane == __ffs(peers) - 1 && key >= 0 && key < C) binw[w * C + key] += acc;
            __syncwarp();
        }
        __syncthreads();
        cl_stage_sum<NT>(stage, q_loc, C, LPT);
    }
    __syncthreads();
    for (int c = tid; c < C; c += NT) {
        float t = 0.0f;
        if (multiply) {
#pragma unroll
            for (int j = 0; j < NW; ++j) t += binw[j * C + c];
        } else {
            q_loc[c] = 0.0f;
        }
        bin_loc[c] = t;
    }
    const double gx_cta = block_sum_d(gxs, redd);  // (barriers inside)
    const float invC_f = 1.0f / (float)C;
    for (int i = tid; i < C * CS; i += NT) {
        const int rr = fdiv_small(i, invC_f), c = i - rr * C;
        cluster.map_shared_rank(epart, rr)[r * C + c] = e_loc[c];
        cluster.map_shared_rank(qpart, rr)[r * C + c] = q_loc[c];
        cluster.map_shared_rank(binpart, rr)[r * C + c] = bin_loc[c];
    }
    if (tid < CS) cluster.map_shared_rank(gxpart, tid)[r] = gx_cta;
    stamp(5);
    cluster.sync();  // #2: per-channel partials of all ranks are local; no remote access after this point

    // ---- phase 4: ds -> dz -> MLP backward -> per-channel coefficients (recomputed by every CTA)
    stamp(6);
    cp_async_wait_all();  // staged MLP weights (made visible to the other threads by the barriers below)
    float kb = 0.0f;
    {
        double gx_tot = 0.0;
        for (int rr = 0; rr < CS; ++rr) gx_tot += gxpart[rr];
        double se_acc = 0.0;
        for (int c = tid; c < C; c += NT) {
            float es = 0.0f, qs = 0.0f, bn = 0.0f;
            for (int rr = 0; rr < CS; ++rr) { es += epart[rr * C + c]; qs += qpart[rr * C + c]; bn += binpart[rr * C + c]; }
            const float s = s_s[c];
            const float dz = fmaf(fmaf(k1, es, qs) * s, 1.0f - s, (1.0f - s) * bn);
            s_dz[c] = dz;
            if (r == 0) bs.dz[b * C + c] = dz;
            if (!multiply) se_acc += (double)s * (double)es;  // add mode: sum_c s_c sum_p g x
        }
        const double ssum = block_sum_d(se_acc, redd);
        if (r == 0 && tid == 0) bs.alphapart[b] = ssum - (sh.pyramid_multiply() ? 0.0 : gx_tot);
        __syncthreads();
        for (int j = w; j < Hd; j += NW) {
            float acc = 0.0f;
#pragma unroll 4
            for (int c = lane; c < C; c += 32) acc = fmaf(s_dz[c], w2p[(size_t)c * Hd + j], acc);
            acc = warp_sum(acc);
            if (lane == 0) {
                const float da = ctx.ha[b * Hd + j] > 0.0f ? acc : 0.0f;
                const float dm = ctx.hm[b * Hd + j] > 0.0f ? acc : 0.0f;
                s_dha[j] = da;
                s_dhm[j] = dm;
                if (r == 0) { bs.dha[b * Hd + j] = da; bs.dhm[b * Hd + j] = dm; }
            }
        }
        __syncthreads();
        const float use = has_mask ? ctx.use[b] : 0.0f;
        const float den = has_mask ? ctx.den[b] : 1.0f;
        const float pass = (has_mask && ctx.msum[b] >= sh.eps) ? 1.0f : 0.0f;  // clamp_min backward
        const float invS = 1.0f / (float)S;
        float kacc = 0.0f;
        for (int c = tid; c < C; c += NT) {
            float davg = 0.0f, dmx = 0.0f;
#pragma unroll 4
            for (int j = 0; j < Hd; ++j) {
                const float wv = w1p[(size_t)j * C + c];
                davg = fmaf(s_dha[j], wv, davg);
                dmx = fmaf(s_dhm[j], wv, dmx);
            }
            const int i = b * C + c;
            const int am = ctx.amax[i];
            const bool dead = has_mask && am < 0;
            const float cA = has_mask ? use * davg / den : 0.0f;
            const float cG = ((1.0f - use) * davg + (dead ? dmx : 0.0f)) * invS;
            const float cM = dead ? 0.0f : dmx;
            const float s = s_s[c];
            const int gi = cl_gm(c, G, CG);
            // dL/dout * gate: multiply: k0 + k1*s*a ; add: k0 + k1*(s + a)
            chA[gi] = make_float4(multiply ? k1 * s : k1, multiply ? k0 : fmaf(k1, s, k0), multiply ? s : 1.0f, cA);
            chG[gi] = cG;
            amx[c] = am;
            s_cM[c] = cM;
            kacc = fmaf(cA, ctx.apool[i] * pass, kacc);
        }
        kb = block_sum(kacc, reinterpret_cast<float*>(redd));
    }
    __syncthreads();

    // ---- phase 5 (T2) over (x,g): dx (streaming store) and R_p = sum_c cA_c x -> dmask
    stamp(7);
    const int g2 = tid / nUmax, ul2 = tid - g2 * nUmax;
    const bool act2 = g2 < G && ul2 < nU;
    const bool want_dmask = has_mask && dmask != nullptr;
    {
        float racc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) racc[i] = 0.0f;
        if (act2) {
            float av[VEC], wv[VEC], wdv[VEC], mv[VEC];
            int ix[VEC];
            lds_f<VEC>(aloc + ul2 * VEC, av);
            lds_f<VEC>(d1s + ul2 * VEC, wv);
            lds_f<VEC>(d0 + ul2 * VEC, wdv);
            lds_f<VEC>(mloc + ul2 * VEC, mv);
            lds_i<VEC>(idxl + ul2 * VEC, ix);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                wdv[i] += wv[i];
                // channel -> position in this thread's walk (g2, g2+G, ...); other groups never match
                const int d = ix[i] - g2;
                ix[i] = (d >= 0 && d % G == 0) ? d / G : -1;
            }
            constexpr int KB = MGA_CL_DXKB;
            const int nj = (C - g2 + G - 1) / G;
            const unsigned gstep = (unsigned)G * rowB;
            const size_t toff = (size_t)g2 * rowB + (size_t)ul2 * 16;
            const char* xp = xbytes + toff;
            const char* gp = gbytes + toff;
            char* op = reinterpret_cast<char*>(dx + sbase) + toff;
            const float4* ca = chA + g2 * CG;
            const float* cg = chG + g2 * CG;
            int j = 0;
            for (; j + KB <= nj; j += KB) {
                uint4 xr[KB], gr[KB];
#pragma unroll
                for (int kc = 0; kc < KB; ++kc) {
                    xr[kc] = ldg128<kPolLast>(addw(xp, kc * gstep), pol);
                    gr[kc] = ldg128<kPolLast>(addw(gp, kc * gstep), pol);
                }
                float cgv[KB];
                lds_f<KB>(cg + j, cgv);
#pragma unroll
                for (int kc = 0; kc < KB; ++kc) {
                    float xv[VEC], gv[VEC], ov[VEC];
                    unpack<T, VEC>(xr[kc], xv);
                    unpack<T, VEC>(gr[kc], gv);
                    const float4 c4 = ca[j + kc];
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const float gate = fmaf(c4.x, av[i], c4.y);
                        const float sel = (ix[i] == j + kc) ? wdv[i] : wv[i];
                        float v = fmaf(c4.w, mv[i], cgv[kc]);
                        v = fmaf(c4.z, sel, v);
                        ov[i] = fmaf(gv[i], gate, v);
                        racc[i] = fmaf(c4.w, xv[i], racc[i]);
                    }
                    __stcs(reinterpret_cast<uint4*>(addw(op, kc * gstep)), pack<T, VEC>(ov));
                }
                xp += (size_t)KB * gstep;
                gp += (size_t)KB * gstep;
                op += (size_t)KB * gstep;
            }
            for (; j < nj; ++j) {
                float xv[VEC], gv[VEC], ov[VEC];
                unpack<T, VEC>(ldg128<kPolLast>(xp, pol), xv);
                unpack<T, VEC>(ldg128<kPolLast>(gp, pol), gv);
                const float4 c4 = ca[j];
                const float cgq = cg[j];
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const float gate = fmaf(c4.x, av[i], c4.y);
                    const float sel = (ix[i] == j) ? wdv[i] : wv[i];
                    float v = fmaf(c4.w, mv[i], cgq);
                    v = fmaf(c4.z, sel, v);
                    ov[i] = fmaf(gv[i], gate, v);
                    racc[i] = fmaf(c4.w, xv[i], racc[i]);
                }
                __stcs(reinterpret_cast<uint4*>(op), pack<T, VEC>(ov));
                xp += gstep;
                gp += gstep;
                op += gstep;
            }
        }
        if (want_dmask && g2 < G && ul2 < nUmax) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) tp[(size_t)g2 * nPmax + ul2 * VEC + i] = racc[i];
        }
    }
    __syncthreads();  // dx of this CTA's pixel range is written (visible to the block): arg-max pixel fix-up below
    for (int c = tid; c < C; c += NT) {
        const int am = amx[c];
        const float cM = s_cM[c];
        if (am >= p0 && am < p0 + nP && cM != 0.0f) {  // dx[c][argmax pixel] += dMx_c (masked_cbam.py:116-117 backward)
            T* e = dx + ((size_t)b * C + c) * S + am;
            *e = from_f<T>(to_f<T>(__ldcg(e)) + cM);
        }
    }
    if (want_dmask) {
        for (int p = tid; p < nP; p += NT) {
            float rsum = 0.0f;
            for (int j = 0; j < G; ++j) rsum += tp[(size_t)j * nPmax + p];
            float dm = (rsum - kb) + d2[p];
            if (sh.sigmoid_mask()) {
                const float m = mloc[p];
                dm *= m * (1.0f - m);
            }
            if (sh.gate_clamp()) {
                const float raw = load_mask_any(mask, mdt, bS + p0 + p);
                if (!(raw >= 0.0f && raw <= 1.0f)) dm = 0.0f;
            }
            store_mask_any(dmask, mdt, bS + p0 + p, dm);
        }
    }
    stamp(8);
}

}  // namespace mga
