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


// ------------------------------------------------------------------ all three pyramid masks in ONE pass over the image
// One CTA per image.  The (H,W) mask is read once with 64-bit loads as 8x8 blocks: per block the pixel count and the
// top-left pixel.  16x16 and 32x32 blocks are sums of 2x2 children, so the strides 8/16/32 come out of one read:
//   nearest  = top-left pixel (H % s == 0: cv2.INTER_NEAREST picks src[i*s][j*s])
//   maxpool  = count > 0;   avgpool = count / s^2 (exact in fp32)
//   area     = rne(count / s^2) on uint8, then "> thresh"     (cv2.INTER_AREA on an integer factor = block mean)
// followed by the optional 3x3 close (dilate, erode; neighbours outside the map ignored) in shared memory.
// Needs H % 32 == 0, W % 32 == 0 (letterboxed YOLO inputs are) and (H/8)*(W/8) <= kMultiMaxBlocks.
constexpr int kMultiThreads = 512;
constexpr int kMultiMaxBlocks = 32768;

struct MultiArgs {
    int B, H, W, method, close3x3, out_f32, binarise;
    float thresh;
};

__device__ __forceinline__ void multi_morph(const uint8_t* src, uint8_t* dst, int h, int w, int op) {
    for (int i = threadIdx.x; i < h * w; i += kMultiThreads) {
        const int y = i / w, x = i - y * w;
        int v = op ? 255 : 0;
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int yy = y + dy, xx = x + dx;
                if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
                const int q = src[yy * w + xx];
                v = op ? min(v, q) : max(v, q);
            }
        dst[i] = (uint8_t)v;
    }
}

// Stage 1 of the two-stage form (mga_masks_multi_ws): one THREAD per 8x8 block over the whole batch -- the only pass over the masks,
// with enough CTAs to pull them at HBM speed (one CTA per image leaves 84 of 148 SMs idle at batch 64).  cnt8 / tl8: (B, H/8 * W/8) bytes.
__global__ void __launch_bounds__(256) masks_count8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ cnt8, uint8_t* __restrict__ tl8,
                                                           int B, int H, int W) {
    const int w8 = W / 8, n8 = (H / 8) * w8;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * n8) return;
    const int b = (int)(i / n8), j = (int)(i - (size_t)b * n8);
    const int by = j / w8, bx = j - by * w8;
    const uint8_t* p = src + (size_t)b * H * W + (size_t)(by * 8) * W + bx * 8;
    int cnt = 0;
    unsigned first = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(p + (size_t)r * W));
        cnt += __popc(v.x & 0x01010101u) + __popc(v.y & 0x01010101u);
        if (r == 0) first = v.x & 0xffu;
    }
    cnt8[i] = (uint8_t)cnt;
    tl8[i] = (uint8_t)(first != 0);
}

__global__ void __launch_bounds__(kMultiThreads) masks_multi_kernel(const uint8_t* __restrict__ src, void* __restrict__ d8, void* __restrict__ d16,
                                                                   void* __restrict__ d32, MultiArgs a, const uint8_t* __restrict__ pre_cnt8,
                                                                   const uint8_t* __restrict__ pre_tl8) {
    extern __shared__ __align__(16) uint8_t msm[];
    const int b = blockIdx.x;
    const int h8 = a.H / 8, w8 = a.W / 8, n8 = h8 * w8;
    const int h16 = h8 / 2, w16 = w8 / 2, n16 = h16 * w16;
    const int h32 = h16 / 2, w32 = w16 / 2, n32 = h32 * w32;
    uint8_t* c8 = msm;             // counts (<= 64 fits a byte; 16x16 and 32x32 counts need 16 bits)
    uint8_t* t8 = c8 + n8;         // top-left pixels
    uint16_t* c16 = reinterpret_cast<uint16_t*>(t8 + n8 + ((2 * n8) & 1));
    uint16_t* c32 = c16 + n16;
    uint8_t* wa = reinterpret_cast<uint8_t*>(c32 + n32);  // work planes for the close: [n8] x 2
    uint8_t* wb = wa + n8;
    const uint8_t* sp = src + (size_t)b * a.H * a.W;
    // two-stage form: the 8x8 block counts were made by masks_count8_kernel and gridDim.y == 3: this CTA derives ONE stride (blockIdx.y)
    const int only = pre_cnt8 != nullptr ? (int)blockIdx.y : -1;
    if (pre_cnt8 != nullptr) {
        for (int i = threadIdx.x; i < n8; i += kMultiThreads) {
            c8[i] = pre_cnt8[(size_t)b * n8 + i];
            t8[i] = pre_tl8[(size_t)b * n8 + i];
        }
    } else
    for (int i = threadIdx.x; i < n8; i += kMultiThreads) {
        const int by = i / w8, bx = i - by * w8;
        const uint8_t* p = sp + (size_t)(by * 8) * a.W + bx * 8;
        int cnt = 0;
        unsigned first = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(p + (size_t)r * a.W));
            // the mask is {0,1} bytes: the byte sum of a word is its popcount
            cnt += __popc(v.x & 0x01010101u) + __popc(v.y & 0x01010101u);
            if (r == 0) first = v.x & 0xffu;
        }
        c8[i] = (uint8_t)cnt;
        t8[i] = (uint8_t)(first != 0);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n16; i += kMultiThreads) {
        const int y = i / w16, x = i - y * w16;
        const int o = (2 * y) * w8 + 2 * x;
        c16[i] = (uint16_t)(c8[o] + c8[o + 1] + c8[o + w8] + c8[o + w8 + 1]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n32; i += kMultiThreads) {
        const int y = i / w32, x = i - y * w32;
        const int o = (2 * y) * w16 + 2 * x;
        c32[i] = (uint16_t)(c16[o] + c16[o + 1] + c16[o + w16] + c16[o + w16 + 1]);
    }
    __syncthreads();
    // per stride: raw map -> (close) -> store
    for (int lvl = 0; lvl < 3; ++lvl) {
        if (only >= 0 && lvl != only) continue;
        const int h = lvl == 0 ? h8 : (lvl == 1 ? h16 : h32), w = lvl == 0 ? w8 : (lvl == 1 ? w16 : w32), n = h * w;
        const int s = 8 << lvl;
        void* dst = lvl == 0 ? d8 : (lvl == 1 ? d16 : d32);
        uint8_t* du = static_cast<uint8_t*>(dst) + (size_t)b * n;
        float* df = static_cast<float*>(dst) + (size_t)b * n;
        const bool direct_f = a.out_f32 && !a.close3x3;
        // raw map into wa (counts / top-left of this level)
        for (int i = threadIdx.x; i < n; i += kMultiThreads) {
            int c;
            uint8_t tl;
            if (lvl == 0) { c = c8[i]; tl = t8[i]; }
            else {
                const int y = i / w, x = i - y * w;
                c = lvl == 1 ? c16[i] : c32[i];
                tl = t8[(y << lvl) * w8 + (x << lvl)];
            }
            const int s2 = s * s;
            uint8_t u;
            float f;
            if (a.method == MGA_DS_NEAREST) { u = tl; f = (float)u; }
            else if (a.method == MGA_DS_MAXPOOL) { u = (uint8_t)(c > 0); f = (float)u; }
            else if (a.method == MGA_DS_AVGPOOL) { f = __fdiv_rn((float)c, (float)s2); u = (uint8_t)(f > 0.0f); }
            else {
                u = sat_u8_rne(__fmul_rn((float)c, __fdiv_rn(1.0f, (float)s2)));
                if (a.binarise) { u = (uint8_t)((float)u > a.thresh); f = (float)u; }
                else f = fminf(fmaxf((float)u, 0.0f), 1.0f);
            }
            if (a.close3x3) wa[i] = u;
            else if (direct_f) df[i] = f;
            else du[i] = u;
        }
        if (a.close3x3) {
            __syncthreads();
            multi_morph(wa, wb, h, w, 0);
            __syncthreads();
            multi_morph(wb, wa, h, w, 1);
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += kMultiThreads) {
                if (a.out_f32) df[i] = (float)wa[i];
                else du[i] = wa[i];
            }
            __syncthreads();
        }
    }
}


// Stage 2 of the two-stage form, banded: grid (row bands, B, 3 strides), one CTA = kBandRows output rows of ONE stride of ONE image.
// The raw map of the band plus a 2-row halo comes from the stage-1 block counts (L2 hits: 6.4 KB per image at 640x640), the 3x3 close
// (dilate rows [r0-1, r1+1), erode rows [r0, r1); neighbours outside the MAP are ignored, exactly as in masks_multi_kernel) runs in
// shared memory.  64 images x (10 + 5 + 3) bands = 1152 CTAs instead of 192: the stage is no longer one latency chain per image.
constexpr int kBandRows = 8, kBandThreads = 256;
__device__ __forceinline__ void multi_raw(int c, uint8_t tl, int s2, const MultiArgs& a, uint8_t* u, float* f) {
    if (a.method == MGA_DS_NEAREST) { *u = tl; *f = (float)tl; }
    else if (a.method == MGA_DS_MAXPOOL) { *u = (uint8_t)(c > 0); *f = (float)*u; }
    else if (a.method == MGA_DS_AVGPOOL) { *f = __fdiv_rn((float)c, (float)s2); *u = (uint8_t)(*f > 0.0f); }
    else {
        *u = sat_u8_rne(__fmul_rn((float)c, __fdiv_rn(1.0f, (float)s2)));
        if (a.binarise) { *u = (uint8_t)((float)*u > a.thresh); *f = (float)*u; }
        else *f = fminf(fmaxf((float)*u, 0.0f), 1.0f);
    }
}
__global__ void __launch_bounds__(kBandThreads) masks_band_kernel(void* __restrict__ d8, void* __restrict__ d16, void* __restrict__ d32, MultiArgs a,
                                                                 const uint8_t* __restrict__ cnt8, const uint8_t* __restrict__ tl8) {
    extern __shared__ __align__(16) uint8_t bsm[];
    const int lvl = blockIdx.z, b = blockIdx.y;
    const int h8 = a.H / 8, w8 = a.W / 8, n8 = h8 * w8;
    const int h = h8 >> lvl, w = w8 >> lvl, n = h * w;
    const int r0 = blockIdx.x * kBandRows;
    if (r0 >= h) return;
    const int r1 = min(h, r0 + kBandRows);
    const int q0 = max(0, r0 - 2), q1 = min(h, r1 + 2);  // raw rows held
    uint8_t* raw = bsm;                                  // [(kBandRows + 4) * w]
    uint8_t* dil = raw + (kBandRows + 4) * w;            // [(kBandRows + 2) * w]
    const uint8_t* c8 = cnt8 + (size_t)b * n8;
    const uint8_t* t8 = tl8 + (size_t)b * n8;
    const int s = 8 << lvl, s2 = s * s, sub = 1 << lvl;
    void* dst = lvl == 0 ? d8 : (lvl == 1 ? d16 : d32);
    uint8_t* du = static_cast<uint8_t*>(dst) + (size_t)b * n;
    float* df = static_cast<float*>(dst) + (size_t)b * n;
    const float iw = 1.0f / (float)w;
    for (int i = threadIdx.x; i < (q1 - q0) * w; i += kBandThreads) {
        const int ry = (int)(((float)i + 0.5f) * iw), x = i - ry * w, y = q0 + ry;
        int c = 0;
        for (int dy = 0; dy < sub; ++dy)
            for (int dx = 0; dx < sub; ++dx) c += c8[((y << lvl) + dy) * w8 + (x << lvl) + dx];
        const uint8_t tl = t8[(y << lvl) * w8 + (x << lvl)];
        uint8_t u;
        float f;
        multi_raw(c, tl, s2, a, &u, &f);
        if (a.close3x3) raw[ry * w + x] = u;
        else if (y >= r0 && y < r1) {
            if (a.out_f32) df[y * w + x] = f;
            else du[y * w + x] = u;
        }
    }
    if (!a.close3x3) return;
    __syncthreads();
    const int e0 = max(0, r0 - 1), e1 = min(h, r1 + 1);  // dilated rows held
    for (int i = threadIdx.x; i < (e1 - e0) * w; i += kBandThreads) {
        const int ry = (int)(((float)i + 0.5f) * iw), x = i - ry * w, y = e0 + ry;
        int v = 0;
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int yy = y + dy, xx = x + dx;
                if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
                v = max(v, (int)raw[(yy - q0) * w + xx]);
            }
        dil[ry * w + x] = (uint8_t)v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (r1 - r0) * w; i += kBandThreads) {
        const int ry = (int)(((float)i + 0.5f) * iw), x = i - ry * w, y = r0 + ry;
        int v = 255;
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int yy = y + dy, xx = x + dx;
                if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
                v = min(v, (int)dil[(yy - e0) * w + xx]);
            }
        if (a.out_f32) df[y * w + x] = (float)v;
        else du[y * w + x] = (uint8_t)v;
    }
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

extern "C" int mga_masks_multi_ws(const uint8_t* src, void* dst8, void* dst16, void* dst32, void* tmp, int32_t B, int32_t H, int32_t W,
                                  int32_t method, float thresh, int32_t close3x3, int32_t out_dtype, void* stream);
extern "C" int mga_masks_multi(const uint8_t* src, void* dst8, void* dst16, void* dst32, int32_t B, int32_t H, int32_t W, int32_t method,
                               float thresh, int32_t close3x3, int32_t out_dtype, void* stream) {
    return mga_masks_multi_ws(src, dst8, dst16, dst32, nullptr, B, H, W, method, thresh, close3x3, out_dtype, stream);
}

extern "C" int mga_masks_multi_ws(const uint8_t* src, void* dst8, void* dst16, void* dst32, void* tmp, int32_t B, int32_t H, int32_t W,
                                  int32_t method, float thresh, int32_t close3x3, int32_t out_dtype, void* stream) {
    auto bad = [&](int code, const char* msg) { return fail(code, "mga_masks_multi: %s", msg); };
    if (!src || !dst8 || !dst16 || !dst32) return bad(MGA_ERR_ARG, "null pointer argument");
    if (B <= 0 || H <= 0 || W <= 0) return bad(MGA_ERR_ARG, "bad mask shape");
    if (method < MGA_DS_NEAREST || method > MGA_DS_AREA_RAW) return bad(MGA_ERR_ARG, "unknown downsample method");
    if (out_dtype != MGA_U8 && out_dtype != MGA_F32) return bad(MGA_ERR_ARG, "mask output dtype must be u8 or f32");
    if (close3x3 && method == MGA_DS_AVGPOOL) return bad(MGA_ERR_ARG, "close3x3 is defined on binary outputs only");
    if (H % 32 || W % 32 || (H / 8) * (W / 8) > kMultiMaxBlocks || (reinterpret_cast<uintptr_t>(src) & 7))
        return bad(MGA_ERR_UNSUPPORTED, "needs H % 32 == 0, W % 32 == 0, (H/8)*(W/8) <= 32768 and an 8-byte aligned source: use mga_mask_downsample per stride");
    MultiArgs a;
    a.B = B; a.H = H; a.W = W; a.method = method; a.close3x3 = close3x3; a.out_f32 = out_dtype == MGA_F32;
    a.binarise = method == MGA_DS_AREA; a.thresh = thresh;
    const int n8 = (H / 8) * (W / 8);
    const size_t smem = (size_t)2 * n8 + 2 + (size_t)2 * (n8 / 4 + n8 / 16) + (size_t)2 * n8 + 16;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(masks_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        configured = true;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* cnt8 = static_cast<uint8_t*>(tmp);
    uint8_t* tl8 = cnt8 ? cnt8 + (size_t)B * n8 : nullptr;
    if (cnt8) {
        const size_t total = (size_t)B * n8;
        masks_count8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, cnt8, tl8, B, H, W);
    }
    if (cnt8) {  // stage 2, banded: (row bands of the stride-8 map, image, stride)
        const int h8 = H / 8, w8 = W / 8;
        const size_t bsmem = (size_t)(2 * kBandRows + 6) * w8;
        masks_band_kernel<<<dim3((h8 + kBandRows - 1) / kBandRows, B, 3), kBandThreads, bsmem, st>>>(dst8, dst16, dst32, a, cnt8, tl8);
    } else {
        masks_multi_kernel<<<dim3(B, 1), kMultiThreads, smem, st>>>(src, dst8, dst16, dst32, a, nullptr, nullptr);
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bad(MGA_ERR_CUDA, cudaGetErrorString(e));
    return MGA_OK;
}
