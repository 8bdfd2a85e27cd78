// Microbenchmark: read bandwidth of a streaming sum kernel vs working-set size (L2-resident vs HBM),
// and copy (read+write) bandwidth.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/l2bw.cu -o /tmp/l2bw
#include <cstdio>
#include <cuda_runtime.h>
__global__ void rd(const float4* __restrict__ p, size_t n, float* out) {
    float acc = 0.f;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * st < n; i += 4 * st) {
        float4 a = p[i], b = p[i + st], c = p[i + 2 * st], d = p[i + 3 * st];
        acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w + d.x + d.y + d.z + d.w;
    }
    for (; i < n; i += st) { float4 a = p[i]; acc += a.x + a.y + a.z + a.w; }
    if (acc == 123.456f) *out = acc;
}
__global__ void cp(const float4* __restrict__ p, float4* __restrict__ q, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * st < n; i += 4 * st) {
        float4 a = p[i], b = p[i + st], c = p[i + 2 * st], d = p[i + 3 * st];
        q[i] = a; q[i + st] = b; q[i + 2 * st] = c; q[i + 3 * st] = d;
    }
    for (; i < n; i += st) q[i] = p[i];
}
int main() {
    const size_t maxb = 512ull << 20;
    float4 *a, *b; float* o;
    cudaMalloc(&a, maxb); cudaMalloc(&b, maxb); cudaMalloc(&o, 4);
    cudaMemset(a, 0, maxb); cudaMemset(b, 0, maxb);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (size_t mb : {8, 16, 24, 32, 48, 64, 96, 128, 256, 512}) {
        size_t n = (mb << 20) / 16;
        for (int grid : {148 * 8, 148 * 16}) {
            for (int w = 0; w < 3; ++w) rd<<<grid, 256>>>(a, n, o);
            cudaEventRecord(e0);
            const int reps = 20;
            for (int r = 0; r < reps; ++r) rd<<<grid, 256>>>(a, n, o);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("read  %4zu MB grid %5d: %7.1f GB/s (%.2f us/launch)\n", mb, grid, (double)(mb << 20) * reps / ms / 1e6, ms * 1e3 / reps);
        }
    }
    for (size_t mb : {16, 32, 64, 256}) {
        size_t n = (mb << 20) / 16;
        for (int w = 0; w < 3; ++w) cp<<<148 * 16, 256>>>(a, b, n);
        cudaEventRecord(e0);
        const int reps = 20;
        for (int r = 0; r < reps; ++r) cp<<<148 * 16, 256>>>(a, b, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("copy  %4zu MB: %7.1f GB/s r+w (%.2f us/launch)\n", mb, 2.0 * (mb << 20) * reps / ms / 1e6, ms * 1e3 / reps);
    }
    return 0;
}
