/*
 * tile_ops.cuh — per-op device bodies shared by the two fused-pass kernels (kernels_tile.cu:
 * cp.async staging, kernels_tma.cu: TMA tensor-map staging).  `a` is the thread's register
 * file of 2^K amplitudes of the current stage.
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

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

template <int K> __device__ __forceinline__ uint32_t reg_offset(int r, const uint32_t (&rb)[K]) {
    uint32_t off = 0;
#pragma unroll
    for (int j = 0; j < K; ++j)
        if (r & (1 << j)) off |= rb[j];
    return off;
}

/* ---- per-op bodies; `a` is the thread's register file of 2^K amplitudes -----------------
 * Every predicate inside an op body is UNIFORM (the same for all threads: it depends on the
 * op and the register index only); the per-thread part of the control predicate is tested
 * once, before the body.  This keeps the amplitudes in fixed registers across the op loop:
 * with per-amplitude divergent predicates ptxas copied the whole register file (32 MOVs)
 * on every op (profiles/r1a_tile_f64_summary.md). */

template <typename real, int K, int J, bool ALL>
__device__ __forceinline__ void apply_gen_pairs(typename Cplx<real>::type (&a)[1 << K], const real *m,
                                                uint32_t regmask) {
    const real m00r = m[0], m00i = m[1], m01r = m[2], m01i = m[3];
    const real m10r = m[4], m10i = m[5], m11r = m[6], m11i = m[7];
#pragma unroll
    for (int r0 = 0; r0 < (1 << K); ++r0) {
        if (r0 & (1 << J)) continue;
        const int r1 = r0 | (1 << J);
        if (ALL || (regmask & (1u << r0))) {
            const real q0r = a[r0].x, q0i = a[r0].y, q1r = a[r1].x, q1i = a[r1].y;
            a[r0].x = m00r * q0r - m00i * q0i + m01r * q1r - m01i * q1i;
            a[r0].y = m00r * q0i + m00i * q0r + m01r * q1i + m01i * q1r;
            a[r1].x = m10r * q0r - m10i * q0i + m11r * q1r - m11i * q1i;
            a[r1].y = m10r * q0i + m10i * q0r + m11r * q1i + m11i * q1r;
        }
    }
}

/* 2x2 on register bit J.  No register-bit controls (regmask all ones, the usual case): one
 * straight-line block of 2^(K-1) independent pairs for the scheduler to interleave. */
template <typename real, int K, int J>
__device__ __forceinline__ void apply_gen(typename Cplx<real>::type (&a)[1 << K], const real *m,
                                          uint32_t regmask) {
    if (regmask == (1u << (1 << K)) - 1u)
        apply_gen_pairs<real, K, J, true>(a, m, regmask);
    else
        apply_gen_pairs<real, K, J, false>(a, m, regmask);
}

/* multiplexed by a register bit: the pairs of `regsel` take m1, the others m (uniform choice) */
template <typename real, int K, int J>
__device__ __forceinline__ void apply_gen_mux_reg(typename Cplx<real>::type (&a)[1 << K], const real *m,
                                                  const real *m1, uint32_t regsel) {
#pragma unroll
    for (int r0 = 0; r0 < (1 << K); ++r0) {
        if (r0 & (1 << J)) continue;
        const int r1 = r0 | (1 << J);
        const real *mm = (regsel & (1u << r0)) ? m1 : m;
        const real q0r = a[r0].x, q0i = a[r0].y, q1r = a[r1].x, q1i = a[r1].y;
        a[r0].x = mm[0] * q0r - mm[1] * q0i + mm[2] * q1r - mm[3] * q1i;
        a[r0].y = mm[0] * q0i + mm[1] * q0r + mm[2] * q1i + mm[3] * q1r;
        a[r1].x = mm[4] * q0r - mm[5] * q0i + mm[6] * q1r - mm[7] * q1i;
        a[r1].y = mm[4] * q0i + mm[5] * q0r + mm[6] * q1i + mm[7] * q1r;
    }
}

template <typename real, int K, int J>
__device__ __forceinline__ void apply_swap(typename Cplx<real>::type (&a)[1 << K], uint32_t regmask,
                                           bool active) {
#pragma unroll
    for (int r0 = 0; r0 < (1 << K); ++r0) {
        if (r0 & (1 << J)) continue;
        const int r1 = r0 | (1 << J);
        if (regmask & (1u << r0)) {
            const typename Cplx<real>::type t0 = a[r0], t1 = a[r1];
            a[r0].x = active ? t1.x : t0.x;
            a[r0].y = active ? t1.y : t0.y;
            a[r1].x = active ? t0.x : t1.x;
            a[r1].y = active ? t0.y : t1.y;
        }
    }
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

/* One op.  There is NO thread-divergent branch in here: a thread whose thread-bit controls
 * are not satisfied runs the same instructions with the identity substituted (matrix,
 * factor or exchange), so the op loop stays uniform and its parameters stay in uniform
 * registers. */
template <typename real, int K>
__device__ __forceinline__ void apply_op(typename Cplx<real>::type (&a)[1 << K], const Op<real> &op,
                                         uint64_t base, uint32_t ebase) {
    /* controls outside the tile are the same for the whole CTA */
    if ((base & op.ctrl_out) != op.ctrl_out) return;
    const bool active = (ebase & op.cmt) == op.cmt;
    const uint32_t regmask = op.regmask;
    const uint32_t arm = op.arm;
    if (arm & (ARM_GEN(0) | ARM_GEN(1) | ARM_GEN(2) | ARM_GEN(3))) {
        if (arm & ARM_MUX_REG) {
            const uint32_t regsel = op.regsel;
            if (arm & ARM_GEN(0)) apply_gen_mux_reg<real, K, 0>(a, op.m, op.m1, regsel);
            if (K > 1 && (arm & ARM_GEN(1))) apply_gen_mux_reg<real, K, (K > 1 ? 1 : 0)>(a, op.m, op.m1, regsel);
            if (K > 2 && (arm & ARM_GEN(2))) apply_gen_mux_reg<real, K, (K > 2 ? 2 : 0)>(a, op.m, op.m1, regsel);
            if (K > 3 && (arm & ARM_GEN(3))) apply_gen_mux_reg<real, K, (K > 3 ? 3 : 0)>(a, op.m, op.m1, regsel);
        } else if (op.cmt == 0 && !(arm & ARM_MUX_THR)) {
            /* no thread-bit controls (the usual case): the matrix stays in uniform registers;
             * a multiplexer outside the tile picks the matrix once per CTA */
            const real *mm = op.m;
            if ((arm & ARM_MUX_OUT) && ((base >> op.mux_out) & 1ull)) mm = op.m1;
            if (arm & ARM_GEN(0)) apply_gen<real, K, 0>(a, mm, regmask);
            if (K > 1 && (arm & ARM_GEN(1))) apply_gen<real, K, (K > 1 ? 1 : 0)>(a, mm, regmask);
            if (K > 2 && (arm & ARM_GEN(2))) apply_gen<real, K, (K > 2 ? 2 : 0)>(a, mm, regmask);
            if (K > 3 && (arm & ARM_GEN(3))) apply_gen<real, K, (K > 3 ? 3 : 0)>(a, mm, regmask);
        } else {
            /* per-thread matrix: multiplexed by a thread bit, or the identity for the threads whose
             * thread-bit controls are not satisfied */
            real m[8];
            if (arm & ARM_MUX_THR) {
                const bool one = (ebase & op.tsel) != 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) m[i] = one ? op.m1[i] : op.m[i];
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) m[i] = active ? op.m[i] : ((i == 0 || i == 6) ? (real)1 : (real)0);
            }
            if (arm & ARM_GEN(0)) apply_gen_pairs<real, K, 0, false>(a, m, regmask);
            if (K > 1 && (arm & ARM_GEN(1))) apply_gen_pairs<real, K, (K > 1 ? 1 : 0), false>(a, m, regmask);
            if (K > 2 && (arm & ARM_GEN(2))) apply_gen_pairs<real, K, (K > 2 ? 2 : 0), false>(a, m, regmask);
            if (K > 3 && (arm & ARM_GEN(3))) apply_gen_pairs<real, K, (K > 3 ? 3 : 0), false>(a, m, regmask);
        }
    } else if (arm & (ARM_SWAP(0) | ARM_SWAP(1) | ARM_SWAP(2) | ARM_SWAP(3))) {
        if (arm & ARM_SWAP(0)) apply_swap<real, K, 0>(a, regmask, active);
        if (K > 1 && (arm & ARM_SWAP(1))) apply_swap<real, K, (K > 1 ? 1 : 0)>(a, regmask, active);
        if (K > 2 && (arm & ARM_SWAP(2))) apply_swap<real, K, (K > 2 ? 2 : 0)>(a, regmask, active);
        if (K > 3 && (arm & ARM_SWAP(3))) apply_swap<real, K, (K > 3 ? 3 : 0)>(a, regmask, active);
    } else {
        const real d0r = active ? op.m[0] : (real)1, d0i = active ? op.m[1] : (real)0;
        const real d1r = active ? op.m[2] : (real)1, d1i = active ? op.m[3] : (real)0;
        if (arm & ARM_DIAG_REG) {
            /* target on a register bit: d1 on the registers of regsel, d0 on the others */
            const uint32_t regsel = op.regsel;
            apply_phase<real, K>(a, d0r, d0i, regmask & ~regsel);
            apply_phase<real, K>(a, d1r, d1i, regmask & regsel);
        } else {
            /* target on a thread bit (OP_DIAG), outside the tile (OP_DIAG_OUT), or a phase */
            bool one = (ebase & op.tsel) != 0;
            if (op.kind == OP_DIAG_OUT) one = (base >> op.bit) & 1ull;
            apply_phase<real, K>(a, one ? d1r : d0r, one ? d1i : d0i, regmask);
        }
    }
}

} // namespace
} // namespace qgb
