// microbench.cuh — measures the two roofline denominators this path is bound by and that
// MEASURED_PEAKS.json does not hold: the NON-FUSED fp64 instruction rate (DADD/DMUL; bit-parity
// with Java forbids DFMA) and shared-memory LDS.128 bandwidth.  Timed with CUDA events.
#pragma once
#include <cuda_runtime.h>

namespace sdpb {

constexpr int kMbChains = 8;

// Each thread: kMbChains independent chains of (mul, add) pairs — the same instruction mix and
// dependency shape as the tiled kernel's accumulators.  Compiled with -fmad=false: no DFMA.
__global__ void __launch_bounds__(256) mb_fp64_nofma(double* out, int iters, double a, double b) {
    double x[kMbChains];
#pragma unroll
    for (int c = 0; c < kMbChains; c++) x[c] = 1.0 + 1e-9 * (threadIdx.x + c);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < kMbChains; c++) {
            x[c] = x[c] * a;
            x[c] = x[c] + b;
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < kMbChains; c++) s += x[c];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) mb_fp64_fma(double* out, int iters, double a, double b) {
    double x[kMbChains];
#pragma unroll
    for (int c = 0; c < kMbChains; c++) x[c] = 1.0 + 1e-9 * (threadIdx.x + c);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < kMbChains; c++) {
            x[c] = __fma_rn(x[c], a, b);
            x[c] = __fma_rn(x[c], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < kMbChains; c++) s += x[c];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Conflict-free 16-byte shared-memory loads, 4 independent per iteration; XOR keeps them live.
__global__ void __launch_bounds__(256) mb_lds128(double* out, int iters) {
    __shared__ ulonglong2 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_ulonglong2(i, ~(unsigned long long)i);
    __syncthreads();
    unsigned long long sx = 0, sy = 0;
    int p = threadIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const ulonglong2 v = buf[(p + k * 256) & 1023];
            sx ^= v.x;
            sy ^= v.y;
        }
        p = (p + 1) & 1023;
    }
    if ((sx ^ sy) == 0x123456789abcull) out[blockIdx.x * blockDim.x + threadIdx.x] = (double)sx;
}

}  // namespace sdpb
