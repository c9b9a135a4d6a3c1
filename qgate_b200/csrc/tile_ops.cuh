/*
 * tile_ops.cuh — small device helpers of the fused-pass kernel (kernels_tma.cu).  `a` is the
 * thread's register file of 2^K amplitudes of the current stage.
 *
 * Arithmetic being reproduced: qgate/simulator/src/CPUQubitProcessor.cpp:307-362
 * (o0 = m00 q0 + m01 q1, o1 = m10 q0 + m11 q1 on every pair whose control bits are 1).
 */
#pragma once
#include <cuda_runtime.h>

#include "kernels.h"

#define QGB_K64 3
#define QGB_K32 4

namespace qgb {
namespace {

template <typename real> struct Cplx;
template <> struct Cplx<float> { typedef float2 type; };
template <> struct Cplx<double> { typedef double2 type; };

template <typename real> __device__ __forceinline__ uint32_t swz(uint32_t e) {
    return tile_swizzle(e, sizeof(real) == 4);
}

/* the same factor (dr, di), per thread, on every selected register */
template <typename real, int K>
__device__ __forceinline__ void apply_phase(typename Cplx<real>::type (&a)[1 << K], real dr, real di,
                                            uint32_t regmask) {
#pragma unroll
    for (int r = 0; r < (1 << K); ++r) {
        if (regmask & (1u << r)) {
            const real qr = a[r].x, qi = a[r].y;
            a[r].x = dr * qr - di * qi;
            a[r].y = dr * qi + di * qr;
        }
    }
}

} // namespace
} // namespace qgb
