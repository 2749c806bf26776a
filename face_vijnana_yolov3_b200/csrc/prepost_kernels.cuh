// Small kernels around the conv stack: activation unpack (debug / parity), letterbox (face_detection.py:657-690), Keras Adam.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_igemm_sm100.cuh"

namespace fvy {

// Debug / parity aid: stored activation -> dense NHWC fp32.
__global__ void unpack_kernel(OutDesc od, int batch, int H, int W, int C, float* __restrict__ dst) {
    const long long total = (long long)batch * H * W * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const long long n = t / H;
        float v;
        if (od.kind == OUT_HEAD_F32) {
            v = reinterpret_cast<const float*>(od.ptr)[((n * H + h) * W + w) * od.c_real + c];
        } else {
            long long row;
            if (od.kind == OUT_PADDED) row = n * od.dst_plane + (long long)(h + 1) * od.dst_w + (w + 1);
            else if (od.kind == OUT_PHASE) {
                const int hp = h + 1, wp = w + 1, ph = ((hp & 1) << 1) | (wp & 1);
                row = ((long long)ph * od.nmax + n) * od.dst_plane + (long long)(hp >> 1) * od.dst_w + (wp >> 1);
            } else row = n * od.dst_plane + (long long)(2 * h + 1) * od.dst_w + (2 * w + 1);
            v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(od.ptr)[row * od.pitch + od.choff + c]);
        }
        dst[i] = v;
    }
}

// ------------------------------------------------------------------------------------------ single convolution (fvy_conv_*)
// Dense NHWC fp32 -> the interior of a padded bf16 buffer (row pitch dst_w pixels, dst_plane pixels per image): 8 channels per thread.
__global__ void __launch_bounds__(256) pack_padded_kernel(const float* __restrict__ x, int batch, int H, int W, int C, int dst_w, int dst_plane,
                                                          __nv_bfloat16* __restrict__ dst) {
    const int cg = C >> 3;
    const long long total = (long long)batch * H * W * cg;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % cg);
        long long t = i / cg;
        const int w = (int)(t % W); t /= W;
        const int hh = (int)(t % H);
        const long long n = t / H;
        const float4* src = reinterpret_cast<const float4*>(x + (((n * H + hh) * W + w) * (long long)C + g * 8));
        const float4 a = src[0], b = src[1];
        __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
        uint4 o;
        o.x = *reinterpret_cast<unsigned*>(&p0); o.y = *reinterpret_cast<unsigned*>(&p1);
        o.z = *reinterpret_cast<unsigned*>(&p2); o.w = *reinterpret_cast<unsigned*>(&p3);
        *reinterpret_cast<uint4*>(dst + ((n * dst_plane + (long long)(hh + 1) * dst_w + (w + 1)) * C + g * 8)) = o;
    }
}
// The same into the 4-phase form a stride-2 consumer reads: padded pixel (hp, wp) = (y + 1, x + 1) of image n goes to phase
// (hp & 1, wp & 1), position (hp >> 1, wp >> 1) of planes with row pitch dst_w and dst_plane rows per image; the four phases are
// phase_rows rows apart.
__global__ void __launch_bounds__(256) pack_phase_kernel(const float* __restrict__ x, int batch, int H, int W, int C, int dst_w, int dst_plane,
                                                         long long phase_rows, __nv_bfloat16* __restrict__ dst) {
    const int cg = C >> 3;
    const long long total = (long long)batch * H * W * cg;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % cg);
        long long t = i / cg;
        const int w = (int)(t % W); t /= W;
        const int hh = (int)(t % H);
        const long long n = t / H;
        const float4* src = reinterpret_cast<const float4*>(x + (((n * H + hh) * W + w) * (long long)C + g * 8));
        const float4 a = src[0], b = src[1];
        __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
        uint4 o;
        o.x = *reinterpret_cast<unsigned*>(&p0); o.y = *reinterpret_cast<unsigned*>(&p1);
        o.z = *reinterpret_cast<unsigned*>(&p2); o.w = *reinterpret_cast<unsigned*>(&p3);
        const int hp = hh + 1, wp = w + 1;
        const long long row = (long long)(((hp & 1) << 1) | (wp & 1)) * phase_rows + n * dst_plane + (long long)(hp >> 1) * dst_w + (wp >> 1);
        *reinterpret_cast<uint4*>(dst + (row * C + g * 8)) = o;
    }
}
// torch Conv2d weight [Co][Ci][k][k] fp32 -> the plan's [cout_pad][taps][cin] bf16.  dgrad = 0: this convolution (the handle was created
// with cin = Ci, cout = Co).  dgrad = 1: the convolution that maps dY to dX (handle created with cin = Co, cout = Ci): weights
// transposed and the taps reversed (both filter axes flipped), dX = conv(dY, flip(W)^T) for a stride-1 'same' convolution.
// tap_perm: the plan of a stride-2 layer stores the column taps of a filter row in the order 0, 2, 1.
__global__ void __launch_bounds__(256) conv_weight_kernel(const float* __restrict__ w, int cin, int cout, int cout_pad, int taps, bool dgrad,
                                                          bool tap_perm, __nv_bfloat16* __restrict__ dst) {
    const long long total = (long long)cout_pad * taps * cin;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % cin);
        int t = (int)((i / cin) % taps);
        const int o = (int)(i / ((long long)cin * taps));
        if (tap_perm) { const int q = t % 3; t = t - q + (q == 0 ? 0 : (q == 1 ? 2 : 1)); }      // stored position -> column tap
        float v = 0.f;
        if (o < cout) v = dgrad ? w[((long long)ci * cout + o) * taps + (taps - 1 - t)] : w[((long long)o * cin + ci) * taps + t];
        dst[i] = __float2bfloat16_rn(v);
    }
}

// ------------------------------------------------------------------------------------------ letterbox (face_detection.py:657-690)
// cv.resize(image / 255, (w_p, h_p), INTER_CUBIC) + zero border, restated operation for operation (OpenCV resizeGeneric_ for
// CV_64F: interpolateCubic with A = -0.75 in float, HResizeCubic then VResizeCubic accumulating in double left to right,
// indices clamped to the image).  One thread per output pixel, three channels; -fmad=false keeps every product and sum
// separately rounded as the C++ reference computes them.
__device__ __forceinline__ void cubic_tab(int d, double scale, int src, int idx[4], float c[4]) {
    const float f0 = (float)(((double)d + 0.5) * scale - 0.5);
    const int s = (int)floorf(f0);
    const float x = f0 - (float)s;
    const float A = -0.75f;
    c[0] = ((A * (x + 1.f) - 5.f * A) * (x + 1.f) + 8.f * A) * (x + 1.f) - 4.f * A;
    c[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
    c[2] = ((A + 2.f) * (1.f - x) - (A + 3.f)) * (1.f - x) * (1.f - x) + 1.f;
    c[3] = 1.f - c[0] - c[1] - c[2];
#pragma unroll
    for (int k = 0; k < 4; ++k) idx[k] = min(max(s + k - 1, 0), src - 1);
}
__global__ void letterbox_u8_kernel(const unsigned char* __restrict__ src, int src_h, int src_w, int w_p, int h_p, int pad_t, int pad_l,
                                    int net_h, int net_w, float* __restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= net_w) return;
    float* o = dst + ((size_t)y * net_w + x) * 3;
    const int dx = x - pad_l, dy = y - pad_t;
    if (dx < 0 || dx >= w_p || dy < 0 || dy >= h_p) { o[0] = o[1] = o[2] = 0.f; return; }
    int xi[4], yi[4];
    float xa[4], ya[4];
    cubic_tab(dx, 1.0 / ((double)w_p / (double)src_w), src_w, xi, xa);
    cubic_tab(dy, 1.0 / ((double)h_p / (double)src_h), src_h, yi, ya);
    double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned char* row = src + (size_t)yi[k] * src_w * 3;
        double r[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double v = ((double)row[xi[0] * 3 + c] / 255.0) * (double)xa[0];
            v = v + ((double)row[xi[1] * 3 + c] / 255.0) * (double)xa[1];
            v = v + ((double)row[xi[2] * 3 + c] / 255.0) * (double)xa[2];
            v = v + ((double)row[xi[3] * 3 + c] / 255.0) * (double)xa[3];
            r[c] = v * (double)ya[k];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] = k == 0 ? r[c] : acc[c] + r[c];
    }
    o[0] = (float)acc[0]; o[1] = (float)acc[1]; o[2] = (float)acc[2];
}

// Keras Adam over a flat bucket: 16 bytes of each of p, g, m, v per thread and iteration (HBM-bound: 28 B per parameter).
// Separately rounded operations (the library is built with -fmad=false) so that the update equals the torch-op restatement.
__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, long long n, float lr_t, float b1, float b2, float eps,
                                                        float gs) {
    const long long n4 = n >> 2;
    const float c1 = 1.0f - b1, c2 = 1.0f - b2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float* pa = &pp.x; float* ma = &mm.x; float* va = &vv.x; const float* ga = &gg.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = ga[k] * gs;
            ma[k] = ma[k] * b1 + c1 * gk;
            va[k] = va[k] * b2 + c2 * (gk * gk);
            pa[k] = pa[k] - (lr_t * ma[k]) / (sqrtf(va[k]) + eps);
        }
        reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {          // tail
        const long long i = (n4 << 2) + threadIdx.x;
        const float gk = g[i] * gs;
        m[i] = m[i] * b1 + c1 * gk;
        v[i] = v[i] * b2 + c2 * (gk * gk);
        p[i] = p[i] - (lr_t * m[i]) / (sqrtf(v[i]) + eps);
    }
}

}  // namespace fvy
