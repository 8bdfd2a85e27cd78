// Microbenchmark (round 2): how fast can a CTA stream its slice of a sample through a shared-memory ring that a DMA
// thread fills with 1-D bulk copies (cp.async.bulk, SASS UBLKCP) -- the data path of the ring kernels (cbam_ring.cuh).
//
// Geometry mimics the cluster kernels: CTA (b, r) owns pixels [r*nP, (r+1)*nP) of all C channel planes of sample b, i.e.
// C rows of nP*4 bytes at a stride of S*4 bytes.  A tile = CH channel rows; the ring has ST tiles.  `passes` passes over
// the slice (the first comes from HBM, the later ones from L2); with write=1 the last pass scales the tile in place and
// a store thread sends it to `out` with bulk shared->global copies.  idle_ns emulates the data-free phases between passes.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/ringbw.cu -o tools/ringbw.bin
//   ./tools/ringbw.bin B C H W CS NT ST CH passes write idle_ns
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, unsigned long long pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

struct P {
    const float* x; float* out; float* sink;
    int C, S, nP, CS, CH, ST, passes, write;
    unsigned idle_ns;
};

__global__ void __launch_bounds__(512, 2) ring_kernel(P p) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* full = reinterpret_cast<uint64_t*>(sm);
    uint64_t* empty = full + 16;
    uint64_t* ready = empty + 16;
    float* ring = reinterpret_cast<float*>(sm + 1024);
    const int NT = blockDim.x, NW = NT / 32, ncw = NW - 1;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x / p.CS, r = blockIdx.x % p.CS;
    const int tileF = p.CH * p.nP, nTiles = p.C / p.CH, total = nTiles * p.passes;
    const size_t base = (size_t)b * p.C * p.S + (size_t)r * p.nP;
    if (tid == 0) {
        for (int i = 0; i < p.ST; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, ncw); mbar_init(ready + i, ncw); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t rowB = (uint32_t)p.nP * 4u;
    if (w == 0) {
        if (lane == 0) {  // load DMA thread
            unsigned long long keep, last;
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(last));
            for (int i = 0; i < total; ++i) {
                const int st = i % p.ST, k = i / p.ST;
                if (k > 0) mbar_wait(empty + st, (k - 1) & 1);
                const int pass = i / nTiles, t = i - pass * nTiles;
                mbar_expect(full + st, rowB * p.CH);
                const float* src = p.x + base + (size_t)t * p.CH * p.S;
                float* dst = ring + (size_t)st * tileF;
                const unsigned long long pol = pass == p.passes - 1 ? last : keep;
                for (int c = 0; c < p.CH; ++c) bulk_g2s(dst + c * p.nP, src + (size_t)c * p.S, rowB, full + st, pol);
            }
        } else if (lane == 1 && p.write) {  // store DMA thread: tiles of the last pass, in order, one bulk group per tile
            const int first = nTiles * (p.passes - 1);
            for (int i = first; i < total; ++i) {
                const int st = i % p.ST;
                mbar_wait(ready + st, ((i - first) / p.ST) & 1);  // ready[] completes one phase per write-pass use of the slot
                const int t = i - first;
                float* dst = p.out + base + (size_t)t * p.CH * p.S;
                const float* src = ring + (size_t)st * tileF;
                for (int c = 0; c < p.CH; ++c) bulk_s2g(dst + (size_t)c * p.S, src + c * p.nP, rowB);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (i > first) {
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    mbar_arrive(empty + (i - 1) % p.ST, ncw);
                }
            }
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            mbar_arrive(empty + (total - 1) % p.ST, ncw);
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        return;
    }
    // consumers
    const int ct = tid - 32, NC = NT - 32;
    float acc = 0.f;
    for (int i = 0; i < total; ++i) {
        const int st = i % p.ST, k = i / p.ST;
        const int pass = i / nTiles;
        if (p.idle_ns && pass > 0 && i == pass * nTiles) {
            const unsigned long long t0 = gtime();
            while (gtime() - t0 < p.idle_ns) {}
        }
        mbar_wait(full + st, k & 1);
        float4* tile = reinterpret_cast<float4*>(ring + (size_t)st * tileF);
        const bool wr = p.write && pass == p.passes - 1;
        for (int u = ct; u < tileF / 4; u += NC) {
            float4 v = tile[u];
            acc += (v.x + v.y) + (v.z + v.w);
            if (wr) { v.x *= 2.f; v.y *= 2.f; v.z *= 2.f; v.w *= 2.f; tile[u] = v; }
        }
        if (wr) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(ready + st, 1);
        } else {
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + st, 1);
        }
    }
    if (acc == 123.456f) *p.sink = acc;
}

// the same work with 128-bit LDG/STG straight from registers (K units in flight per thread), for comparison
__global__ void __launch_bounds__(512, 2) ldg_kernel(P p) {
    const int NT = blockDim.x, tid = threadIdx.x;
    const int b = blockIdx.x / p.CS, r = blockIdx.x % p.CS;
    const size_t base = (size_t)b * p.C * p.S + (size_t)r * p.nP;
    const int nU = p.nP / 4, total = p.C * nU;
    float acc = 0.f;
    for (int pass = 0; pass < p.passes; ++pass) {
        if (p.idle_ns && pass > 0) {
            const unsigned long long t0 = gtime();
            while (gtime() - t0 < p.idle_ns) {}
        }
        const bool wr = p.write && pass == p.passes - 1;
        for (int i0 = tid; i0 < total; i0 += 4 * NT) {
            float4 v[4];
            size_t off[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i0 + j * NT;
                const int c = i / nU, u = i - c * nU;
                off[j] = base + (size_t)c * p.S + u * 4;
                if (i < total) v[j] = __ldcg(reinterpret_cast<const float4*>(p.x + off[j]));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (i0 + j * NT < total) {
                    acc += (v[j].x + v[j].y) + (v[j].z + v[j].w);
                    if (wr) { v[j].x *= 2.f; v[j].y *= 2.f; v[j].z *= 2.f; v[j].w *= 2.f; __stcs(reinterpret_cast<float4*>(p.out + off[j]), v[j]); }
                }
            }
        }
    }
    if (acc == 123.456f) *p.sink = acc;
}

int main(int argc, char** argv) {
    auto arg = [&](int i, int d) { return argc > i ? atoi(argv[i]) : d; };
    const int B = arg(1, 64), C = arg(2, 64), H = arg(3, 80), W = arg(4, 80), CS = arg(5, 8), NT = arg(6, 512), ST = arg(7, 4), CH = arg(8, 4);
    const int passes = arg(9, 3), write = arg(10, 1), idle = arg(11, 0);
    const int S = H * W, nP = S / CS;
    const size_t N = (size_t)B * C * S;
    const int nset = 3;
    float *x[nset], *o[nset], *sink;
    for (int i = 0; i < nset; ++i) { CK(cudaMalloc(&x[i], N * 4)); CK(cudaMalloc(&o[i], N * 4)); CK(cudaMemset(x[i], 0, N * 4)); CK(cudaMemset(o[i], 0, N * 4)); }
    CK(cudaMalloc(&sink, 4));
    const int smem = 1024 + ST * CH * nP * 4;
    CK(cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 0, occ2 = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ring_kernel, NT, smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, ldg_kernel, NT, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const double bytes = (double)N * 4 * (1 + (write ? 1 : 0));  // algorithmic: x once (+ out once)
    for (int variant = 0; variant < 2; ++variant) {
        float best = 1e9f, sum = 0.f;
        const int reps = 12;
        for (int it = 0; it < reps + 3; ++it) {
            P p{x[it % nset], o[it % nset], sink, C, S, nP, CS, CH, ST, passes, write, (unsigned)idle};
            CK(cudaEventRecord(e0));
            if (variant == 0) ring_kernel<<<B * CS, NT, smem>>>(p);
            else ldg_kernel<<<B * CS, NT>>>(p);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it >= 3) { best = ms < best ? ms : best; sum += ms; }
        }
        printf("%s B%d C%d %dx%d CS%d NT%d ST%d CH%d passes%d write%d idle%d smem%d occ%d: avg %.1f us best %.1f us -> %.0f GB/s algorithmic (avg)\n",
               variant == 0 ? "ring" : "ldg ", B, C, H, W, CS, NT, ST, CH, passes, write, idle, smem, variant == 0 ? occ : occ2, sum / reps * 1e3,
               best * 1e3, bytes / (sum / reps * 1e-3) * 1e-9);
    }
    return 0;
}
