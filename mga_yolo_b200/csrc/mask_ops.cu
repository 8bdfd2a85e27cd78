// mask_ops.cu -- binary-mask downsample to pyramid resolution, bit-exact with what the
// reference gets from OpenCV (mga_yolo/utils/mask_utils.py:14-48,64-141).  Integer / byte
// work: one thread per output pixel, coalesced byte reads along W, no tensor cores.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace mga {

struct DsArgs {
    int B, H, W, nh, nw, stride, method;
    float thresh;
    int out_f32;   // write float32 instead of uint8
    int binarise;  // apply "> thresh" (downsample_mask 'area')
};

// (dst, src, weight) entries of OpenCV's computeResizeAreaTab for ONE destination index,
// visited in emission order.  Weights are computed in double and rounded to float, like cv2.
template <typename F>
__device__ __forceinline__ void area_entries(int d, int ssize, int dsize, F&& fn) {
    const double scale = (double)ssize / (double)dsize;
    const double f1 = d * scale, f2 = f1 + scale;
    const double cell = fmin(scale, (double)ssize - f1);
    int s1 = (int)ceil(f1), s2 = (int)floor(f2);
    s2 = min(s2, ssize - 1);
    s1 = min(s1, s2);
    if (s1 - f1 > 1e-3) fn(s1 - 1, (float)((s1 - f1) / cell));
    for (int sx = s1; sx < s2; ++sx) fn(sx, (float)(1.0 / cell));
    if (f2 - s2 > 1e-3) fn(s2, (float)(fmin(fmin(f2 - s2, 1.0), cell) / cell));
}

__device__ __forceinline__ uint8_t sat_u8_rne(float v) {
    const float r = rintf(v);  // round half to even, like cvRound
    return (uint8_t)fminf(fmaxf(r, 0.0f), 255.0f);
}

__global__ void __launch_bounds__(kBlock) mask_downsample_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst_u8,
                                                                 float* __restrict__ dst_f32, DsArgs a) {
    const size_t total = (size_t)a.B * a.nh * a.nw;
    const size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= total) return;
    const int ox = (int)(i % a.nw), oy = (int)((i / a.nw) % a.nh), b = (int)(i / ((size_t)a.nw * a.nh));
    const uint8_t* sp = src + (size_t)b * a.H * a.W;
    float fval = 0.0f;
    uint8_t uval = 0;
    if (a.method == MGA_DS_NEAREST) {
        const int sy = min((int)floor(oy * ((double)a.H / a.nh)), a.H - 1);
        const int sx = min((int)floor(ox * ((double)a.W / a.nw)), a.W - 1);
        uval = sp[(size_t)sy * a.W + sx];
        fval = (float)uval;
    } else if (a.method == MGA_DS_MAXPOOL || a.method == MGA_DS_AVGPOOL) {
        int cnt = 0, any = 0;
        const int y1 = min((oy + 1) * a.stride, a.H), x1 = min((ox + 1) * a.stride, a.W);
        for (int y = oy * a.stride; y < y1; ++y)
            for (int x = ox * a.stride; x < x1; ++x) {
                const int v = sp[(size_t)y * a.W + x];
                cnt += v;
                any = max(any, v);
            }
        if (a.method == MGA_DS_MAXPOOL) { uval = (uint8_t)any; fval = (float)any; }
        else { fval = __fdiv_rn((float)cnt, (float)(a.stride * a.stride)); uval = (uint8_t)(fval > 0.0f); }
    } else {  // INTER_AREA on uint8
        if (a.H % a.nh == 0 && a.W % a.nw == 0) {
            const int ky = a.H / a.nh, kx = a.W / a.nw;
            int sum = 0;
            for (int y = oy * ky; y < (oy + 1) * ky; ++y)
                for (int x = ox * kx; x < (ox + 1) * kx; ++x) sum += sp[(size_t)y * a.W + x];
            const float scale = __fdiv_rn(1.0f, (float)(ky * kx));
            uval = sat_u8_rne(__fmul_rn((float)sum, scale));
        } else {
            float acc = 0.0f;
            bool first = true;
            area_entries(oy, a.H, a.nh, [&](int sy, float beta) {
                float buf = 0.0f;
                area_entries(ox, a.W, a.nw, [&](int sx, float alpha) {
                    buf = __fadd_rn(buf, __fmul_rn((float)sp[(size_t)sy * a.W + sx], alpha));
                });
                const float t = __fmul_rn(beta, buf);
                acc = first ? t : __fadd_rn(acc, t);
                first = false;
            });
            uval = sat_u8_rne(acc);
        }
        if (a.binarise) uval = (uint8_t)((float)uval > a.thresh);
        fval = fminf(fmaxf((float)uval, 0.0f), 1.0f);
        if (a.binarise) fval = (float)uval;
    }
    if (a.out_f32) dst_f32[i] = fval;
    else dst_u8[i] = uval;
}

// 3x3 dilate (op = 0) / erode (op = 1) on uint8, neighbours outside the image ignored.
__global__ void __launch_bounds__(kBlock) morph3x3_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst_u8,
                                                          float* __restrict__ dst_f32, int B, int H, int W, int op) {
    const size_t total = (size_t)B * H * W;
    const size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const uint8_t* sp = src + (i - (size_t)y * W - x);
    int v = op ? 255 : 0;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            const int yy = y + dy, xx = x + dx;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const int s = sp[(size_t)yy * W + xx];
            v = op ? min(v, s) : max(v, s);
        }
    if (dst_f32) dst_f32[i] = (float)v;
    else dst_u8[i] = (uint8_t)v;
}

}  // namespace mga

using namespace mga;

extern "C" int mga_mask_downsample(const uint8_t* src, void* dst, void* tmp, int32_t B, int32_t H, int32_t W, int32_t stride,
                                   int32_t method, float thresh, int32_t close3x3, int32_t out_dtype, void* stream) {
    auto bad = [&](int code, const char* msg) { return fail(code, "mga_mask_downsample: %s", msg); };
    if (!src || !dst) return bad(MGA_ERR_ARG, "null pointer argument");
    if (B <= 0 || H <= 0 || W <= 0 || stride < 1) return bad(MGA_ERR_ARG, "bad mask shape / stride");
    if (method < MGA_DS_NEAREST || method > MGA_DS_AREA_RAW) return bad(MGA_ERR_ARG, "unknown downsample method");
    if (out_dtype != MGA_U8 && out_dtype != MGA_F32) return bad(MGA_ERR_ARG, "mask output dtype must be u8 or f32");
    if (close3x3 && !tmp) return bad(MGA_ERR_WORKSPACE, "close3x3 needs a tmp buffer of 2*B*nh*nw bytes");
    if (close3x3 && method == MGA_DS_AVGPOOL) return bad(MGA_ERR_ARG, "close3x3 is defined on binary outputs only");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DsArgs a;
    a.B = B; a.H = H; a.W = W; a.stride = stride; a.method = method; a.thresh = thresh;
    a.nh = (H + stride - 1) / stride;
    a.nw = (W + stride - 1) / stride;
    a.binarise = method == MGA_DS_AREA;
    const size_t total = (size_t)B * a.nh * a.nw;
    const int grid = (int)((total + kBlock - 1) / kBlock);
    if (!close3x3) {
        a.out_f32 = out_dtype == MGA_F32;
        mask_downsample_kernel<<<grid, kBlock, 0, st>>>(src, static_cast<uint8_t*>(dst), static_cast<float*>(dst), a);
    } else {
        // downsample -> dst(u8 view of tmp2) ... two byte planes ping-pong: tmp holds the raw map, then the dilation
        a.out_f32 = 0;
        uint8_t* t0 = static_cast<uint8_t*>(tmp);
        uint8_t* t1 = t0 + total;  // caller provides 2*total bytes
        mask_downsample_kernel<<<grid, kBlock, 0, st>>>(src, t0, nullptr, a);
        morph3x3_kernel<<<grid, kBlock, 0, st>>>(t0, t1, nullptr, B, a.nh, a.nw, 0);
        if (out_dtype == MGA_F32) morph3x3_kernel<<<grid, kBlock, 0, st>>>(t1, nullptr, static_cast<float*>(dst), B, a.nh, a.nw, 1);
        else morph3x3_kernel<<<grid, kBlock, 0, st>>>(t1, static_cast<uint8_t*>(dst), nullptr, B, a.nh, a.nw, 1);
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bad(MGA_ERR_CUDA, cudaGetErrorString(e));
    return MGA_OK;
}
