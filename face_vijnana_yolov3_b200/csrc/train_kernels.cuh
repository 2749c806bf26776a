// Training-step kernels (first slice): BatchNormalization in TRAINING mode (batch statistics, Keras momentum
// 0.99 = torch momentum 0.01, eps 1e-3; reference layers: src/space/yolov3_detect.py:212, trained through
// src/space/face_detection.py:361-381, :602-630) fused with the LeakyReLU(0.1) that follows every one of them (:213), forward
// and backward, over NHWC fp32 activations.  HBM-bound: forward reads x twice and writes y (12 B per element), backward reads
// x and dy twice and writes dx (20 B per element).  Sums are accumulated in double (E[x^2] - mean^2 cancels badly in float).
//   forward : mean_c = sum x / M, var_c = sum x^2 / M - mean^2 (biased, used to normalise), y = leaky(gamma (x - mean) invstd + beta),
//             running_mean += momentum (mean - running_mean), running_var += momentum (var M / (M - 1) - running_var)
//   backward: dz = dy * (z > 0 ? 1 : slope), dbeta = sum dz, dgamma = sum dz xhat,
//             dx = gamma invstd (dz - dbeta / M - xhat dgamma / M)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fvy {

constexpr int kBnThreads = 256;      // 32 channel quads (128 channels) x 8 row lanes

// mode 0: sum x, sum x^2;  mode 1: sum dz, sum dz * xhat
template <int MODE>
__global__ void __launch_bounds__(kBnThreads) bn_sums_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long rows, int C,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const float* __restrict__ mean, const float* __restrict__ invstd, float slope,
                                                             double* __restrict__ ws /*[2][C]*/) {
    __shared__ double sh[2][8][128];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.y * 128 + tx * 4;
    double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    if (c < C) {
        float g[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0}, m[4] = {0, 0, 0, 0}, is[4] = {0, 0, 0, 0};
        if (MODE == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { g[k] = gamma[c + k]; b[k] = beta[c + k]; m[k] = mean[c + k]; is[k] = invstd[c + k]; }
        }
        for (long long r = (long long)blockIdx.x * 8 + ty; r < rows; r += (long long)gridDim.x * 8) {
            const float4 v = *reinterpret_cast<const float4*>(x + r * C + c);
            const float xv[4] = {v.x, v.y, v.z, v.w};
            if (MODE == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { s[k] += (double)xv[k]; q[k] += (double)xv[k] * (double)xv[k]; }
            } else {
                const float4 d = *reinterpret_cast<const float4*>(dy + r * C + c);
                const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float xh = (xv[k] - m[k]) * is[k];
                    const float z = g[k] * xh + b[k];
                    const float dz = z > 0.f ? dv[k] : dv[k] * slope;
                    s[k] += (double)dz; q[k] += (double)dz * (double)xh;
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { sh[0][ty][tx * 4 + k] = s[k]; sh[1][ty][tx * 4 + k] = q[k]; }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int cc = blockIdx.y * 128 + threadIdx.x;
        if (cc < C) {
            double a = 0, b2 = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { a += sh[0][j][threadIdx.x]; b2 += sh[1][j][threadIdx.x]; }
            atomicAdd(ws + cc, a);
            atomicAdd(ws + C + cc, b2);
        }
    }
}

__global__ void bn_fwd_finalize_kernel(const double* __restrict__ ws, long long rows, int C, float eps, float momentum, float* __restrict__ running_mean,
                                       float* __restrict__ running_var, float* __restrict__ save_mean, float* __restrict__ save_invstd) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double M = (double)rows;
    const double mean = ws[c] / M;
    double var = ws[C + c] / M - mean * mean;
    if (var < 0) var = 0;
    save_mean[c] = (float)mean;
    save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mean);
    if (running_var) running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * (rows > 1 ? var * M / (M - 1.0) : var));
}

__global__ void __launch_bounds__(256) bn_fwd_apply_kernel(const float* __restrict__ x, long long n4, int C, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd, float slope, float* __restrict__ y) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)((i * 4) % C);
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        const float xv[4] = {v.x, v.y, v.z, v.w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float z = gamma[c + k] * ((xv[k] - mean[c + k]) * invstd[c + k]) + beta[c + k];
            o[k] = z > 0.f ? z : z * slope;
        }
        reinterpret_cast<float4*>(y)[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

__global__ void bn_bwd_finalize_kernel(const double* __restrict__ ws, int C, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    dbeta[c] = (float)ws[c];
    dgamma[c] = (float)ws[C + c];
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long n4, long long rows, int C,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ mean, const float* __restrict__ invstd, float slope,
                                                           const double* __restrict__ ws, float* __restrict__ dx) {
    const float inv_m = (float)(1.0 / (double)rows);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)((i * 4) % C);
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        const float4 d = reinterpret_cast<const float4*>(dy)[i];
        const float xv[4] = {v.x, v.y, v.z, v.w}, dv[4] = {d.x, d.y, d.z, d.w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float is = invstd[c + k], g = gamma[c + k];
            const float xh = (xv[k] - mean[c + k]) * is;
            const float z = g * xh + beta[c + k];
            const float dz = z > 0.f ? dv[k] : dv[k] * slope;
            o[k] = g * is * (dz - (float)ws[c + k] * inv_m - xh * ((float)ws[C + c + k] * inv_m));
        }
        reinterpret_cast<float4*>(dx)[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

}  // namespace fvy
