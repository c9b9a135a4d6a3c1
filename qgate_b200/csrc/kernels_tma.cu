/*
 * kernels_tma.cu — the fused gate pass with TMA staging (sm_100a).
 *
 *   tma_pass_kernel   same pass program and op bodies as tile_pass_kernel (kernels_tile.cu),
 *                     but a tile travels HBM -> shared memory -> HBM as ONE tensor-map copy
 *                     each way (cp.async.bulk.tensor, UTMALDG / UTMASTG), issued by one
 *                     thread and tracked by mbarriers / bulk groups.  The per-thread address
 *                     arithmetic of the cp.async version (about a third of its instruction
 *                     stream, profiles/r1g_tile_pass.md) disappears, and a ring of NBUF tile
 *                     buffers keeps a load, the gate stages and a store in flight at once.
 *
 * The tile is the sub-cube of T tile lanes (program.h).  The state vector is described to the
 * TMA unit as a 5-dimensional tensor: dimension 0 is one 128-byte row (the low row_lanes
 * lanes), dimensions 1..4 are the planner's lane groups [tile lanes][other lanes]
 * (PassProgram::grp_*): the box takes the tile lanes of every group, the tile number supplies
 * the coordinates of the other lanes.  CU_TENSOR_MAP_SWIZZLE_128B gives exactly the
 * bank-conflict-free layout the stage code expects (program.h: tile_swizzle).
 *
 * Reference semantics: qgate/simulator/src/CPUQubitProcessor.cpp:307-362; the incumbent device
 * code is DeviceProcPrimitives.cu:198-249 (one launch and one state sweep per gate).
 */
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>

#include "kernels.h"
#include "tile_ops.cuh"

/* build.py compiles this file twice (the complex128 and the complex64 instantiations in parallel):
 * QGB_TMA_PART = 0 -> complex128 + the shared state and setters, 1 -> complex64.  Undefined: everything. */
#if !defined(QGB_TMA_PART)
#define QGB_TMA_F64 1
#define QGB_TMA_F32 1
#define QGB_TMA_SHARED 1
#elif QGB_TMA_PART == 0
#define QGB_TMA_F64 1
#define QGB_TMA_SHARED 1
#else
#define QGB_TMA_F32 1
#endif

namespace qgb {

namespace {

struct TmaGeometry {
    int32_t n_groups;
    int32_t shift[QGB_MAX_GROUPS];   /* first tile-number bit consumed by the group            */
    uint32_t mask[QGB_MAX_GROUPS];   /* (1 << grp_r) - 1                                       */
    int32_t tbits[QGB_MAX_GROUPS];   /* coordinate = rest bits << tbits                        */
    int32_t base_shift[QGB_MAX_GROUPS]; /* lane of the group's first non-tile lane            */
    unsigned long long *phase;       /* QGB_PHASE_TIMING=1: cycles per phase, summed over warp 0 of */
                                     /* every CTA (diagnostic build only, see tma_pass_phase_report) */
    int32_t l2_hint;                 /* 1: loads, 2: stores, 3: both carry an L2 evict_first policy  */
    const void *fan_tiles;           /* phase fans: per fan two tables of TILE factors (device memory, */
                                     /* built by fan_tile_table_kernel before the pass), indexed by the  */
                                     /* low fan_lo bits of the tile number and by the rest; nullptr when */
                                     /* no fan of the pass has a term outside the tile                   */
    int32_t fan_lo, fan_hi;
    int32_t zero_input;              /* 1: the state is |0...0> and has not been written yet: no tile is */
                                     /* loaded, the first stage starts from zeros (amplitude 0 = 1)      */
    int32_t debug_mode;              /* measurement only (option debug_pass_mode): 1 = the stages   */
                                     /* load and store their registers but skip the ops, 2 = no     */
                                     /* stages at all (the tile only travels in and out)             */
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
/* barrier among the first `count` threads of the CTA (a multiple of 32): the consumer warps */
__device__ __forceinline__ void named_sync(uint32_t count) {
    asm volatile("bar.sync 1, %0;\n" ::"r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void *dst, const CUtensorMap *map, uint64_t *bar, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n" ::"r"(smem_u32(dst)),
        "l"(map), "r"(0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap *map, const void *src, int c1, int c2, int c3,
                                             int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];\n" ::"l"(map),
                 "r"(0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_5d_hint(void *dst, const CUtensorMap *map, uint64_t *bar, int c1, int c2,
                                                 int c3, int c4, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3, %4, %5, %6}], [%7], %8;\n" ::"r"(smem_u32(dst)),
        "l"(map), "r"(0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_store_5d_hint(const CUtensorMap *map, const void *src, int c1, int c2, int c3,
                                                  int c4, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2, %3, %4, %5}], [%6], %7;\n" ::"l"(map),
                 "r"(0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(src)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

/* ---- op bodies of this kernel ----------------------------------------------------------
 * One specialised body per Op::code, entered through a switch on a CTA-uniform value.  The 2x2
 * matrices of the pass live in shared memory (copied there once per CTA): a body fetches its
 * matrix with 128-bit shared loads through a pointer that may differ per thread (gate
 * multiplexed by a thread bit), so there is no per-thread select of 8 values; thread-bit
 * controls are a branch around the body (a thread either runs it or skips it, nothing else
 * diverges); register-bit controls and register-bit multiplexers are resolved at compile
 * time or by uniform tests.  Arithmetic: CPUQubitProcessor.cpp:307-362. */

template <typename real> struct Mat8;
template <> struct Mat8<double> {
    double v[8];
    /* diagonal factor `which` (0: d0 = m[0..1], 1: d1 = m[2..3]) */
    __device__ __forceinline__ void factor(int which, double &re, double &im) const {
        re = v[2 * which], im = v[2 * which + 1];
    }
    __device__ __forceinline__ void load(const double *p) {
        const double2 *q = reinterpret_cast<const double2 *>(p);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double2 t = q[i];
            v[2 * i] = t.x;
            v[2 * i + 1] = t.y;
        }
    }
};
/* complex64: one amplitude (re, im) is one 64-bit register and the arithmetic is the packed
 * FFMA2 of sm_100 (fma.rn.f32x2): m * q = (mr, mr) * (x, y) + (-mi, mi) * (y, x) — ptxas folds the
 * broadcast and the half swap into operand modifiers (R.F32, .F32x2.LO_HI), so a 2x2 on a pair
 * is 8 instructions instead of 16.  A matrix is stored as 12 floats: the four real parts, then
 * (-mi, mi) of the four entries. */
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 swap2(u64 v) {
    float lo, hi;
    upk2(v, lo, hi);
    return pk2(hi, lo);
}
__device__ __forceinline__ u64 as_u64(const float2 &q) { return pk2(q.x, q.y); }
__device__ __forceinline__ float2 as_f2(u64 v) {
    float2 q;
    upk2(v, q.x, q.y);
    return q;
}

template <> struct Mat8<float> {
    float r[4]; /* real parts of m00, m01, m10, m11 */
    u64 ni[4];  /* (-im, im) of the same              */
    /* diagonal factor `which`: d0 = (m[0], m[1]) -> r[0], ni[0]; d1 = (m[2], m[3]) -> r[1], ni[1] */
    __device__ __forceinline__ void factor(int which, float &re, float &im) const {
        float neg;
        re = r[which];
        upk2(ni[which], neg, im);
    }
    __device__ __forceinline__ void load(const float *p) {
        const float4 *q = reinterpret_cast<const float4 *>(p);
        const float4 t0 = q[0], t1 = q[1], t2 = q[2];
        r[0] = t0.x, r[1] = t0.y, r[2] = t0.z, r[3] = t0.w;
        ni[0] = pk2(t1.x, t1.y), ni[1] = pk2(t1.z, t1.w), ni[2] = pk2(t2.x, t2.y), ni[3] = pk2(t2.z, t2.w);
    }
};

template <typename real> struct MatLayout;
template <> struct MatLayout<double> {
    static constexpr int kStride = 8; /* reals per matrix in shared memory */
    /* value k (0..7: re, im of m00, m01, m10, m11) of a matrix -> what is stored at slot i */
    __device__ static double slot(const double *m, int i) { return m[i]; }
    __device__ static void factor(const double *blk, int which, double &re, double &im) {
        re = blk[2 * which], im = blk[2 * which + 1];
    }
};
template <> struct MatLayout<float> {
    static constexpr int kStride = 12;
    __device__ static float slot(const float *m, int i) {
        if (i < 4) return m[2 * i];
        const int e = (i - 4) >> 1;
        return ((i - 4) & 1) ? m[2 * e + 1] : -m[2 * e + 1];
    }
    /* diagonal factor `which` (0: d0 = m[0..1], 1: d1 = m[2..3]) */
    __device__ static void factor(const float *blk, int which, float &re, float &im) {
        re = blk[which], im = blk[5 + 2 * which];
    }
};

__device__ __forceinline__ void pair_2x2(double2 &x0, double2 &x1, const Mat8<double> &mt) {
    const double(&m)[8] = mt.v;
    const double q0r = x0.x, q0i = x0.y, q1r = x1.x, q1i = x1.y;
    x0.x = m[0] * q0r - m[1] * q0i + m[2] * q1r - m[3] * q1i;
    x0.y = m[0] * q0i + m[1] * q0r + m[2] * q1i + m[3] * q1r;
    x1.x = m[4] * q0r - m[5] * q0i + m[6] * q1r - m[7] * q1i;
    x1.y = m[4] * q0i + m[5] * q0r + m[6] * q1i + m[7] * q1r;
}

__device__ __forceinline__ void pair_2x2(float2 &x0, float2 &x1, const Mat8<float> &mt) {
    const u64 q0 = as_u64(x0), q1 = as_u64(x1), q0s = swap2(q0), q1s = swap2(q1);
    u64 o0 = mul2(pk2(mt.r[0], mt.r[0]), q0);
    u64 o1 = mul2(pk2(mt.r[2], mt.r[2]), q0);
    o0 = fma2(mt.ni[0], q0s, o0);
    o1 = fma2(mt.ni[2], q0s, o1);
    o0 = fma2(pk2(mt.r[1], mt.r[1]), q1, o0);
    o1 = fma2(pk2(mt.r[3], mt.r[3]), q1, o1);
    o0 = fma2(mt.ni[1], q1s, o0);
    o1 = fma2(mt.ni[3], q1s, o1);
    x0 = as_f2(o0);
    x1 = as_f2(o1);
}

/* a[r] *= (dr + i di) on the registers of `regmask` */
template <int K>
__device__ __forceinline__ void lean_phase(double2 (&a)[1 << K], double dr, double di, uint32_t regmask) {
    apply_phase<double, K>(a, dr, di, regmask);
}
template <int K>
__device__ __forceinline__ void lean_phase(float2 (&a)[1 << K], float dr, float di, uint32_t regmask) {
    const u64 rr = pk2(dr, dr), ni = pk2(-di, di);
#pragma unroll
    for (int r = 0; r < (1 << K); ++r) {
        if (regmask & (1u << r)) {
            const u64 q = as_u64(a[r]);
            a[r] = as_f2(fma2(ni, swap2(q), mul2(rr, q)));
        }
    }
}

/* MODE 0: every pair; 1: pairs allowed by regmask (uniform); 2 / 3: pairs whose register bit J2 is
 * 0 / 1 (compile time) */
template <typename real, int K, int J, int MODE, int J2>
__device__ __forceinline__ void lean_gen(typename Cplx<real>::type (&a)[1 << K], const Mat8<real> &m, uint32_t regmask) {
#pragma unroll
    for (int r0 = 0; r0 < (1 << K); ++r0) {
        if (r0 & (1 << J)) continue;
        if (MODE == 2 && (r0 & (1 << J2))) continue;
        if (MODE == 3 && !(r0 & (1 << J2))) continue;
        if (MODE == 1 && !(regmask & (1u << r0))) continue;
        pair_2x2(a[r0], a[r0 | (1 << J)], m);
    }
}

template <typename real, int K, int J, int J2>
__device__ __forceinline__ void lean_regmux(typename Cplx<real>::type (&a)[1 << K], const Mat8<real> &m0,
                                            const real *mats) {
    if (J == J2) return; /* never planned */
    Mat8<real> m1;
    if (sizeof(real) == 4) m1.load(mats + MatLayout<real>::kStride); /* in flight under the first half */
    lean_gen<real, K, J, 2, (J == J2 ? 0 : J2)>(a, m0, 0u);
    if (sizeof(real) == 8) m1.load(mats + MatLayout<real>::kStride); /* complex128: no registers to spare */
    lean_gen<real, K, J, 3, (J == J2 ? 0 : J2)>(a, m1, 0u);
}

template <typename real, int K, int J>
__device__ __forceinline__ void lean_swap(typename Cplx<real>::type (&a)[1 << K], uint32_t regmask) {
#pragma unroll
    for (int r0 = 0; r0 < (1 << K); ++r0) {
        if (r0 & (1 << J)) continue;
        if (regmask & (1u << r0)) {
            const typename Cplx<real>::type t = a[r0];
            a[r0] = a[r0 | (1 << J)];
            a[r0 | (1 << J)] = t;
        }
    }
}

#define QGB_J(j) ((j) < K ? (j) : 0)

/* ---- OP_SHEAR: M = phi P^s N, N = [[1,x],[0,1]] [[1,0],[g,1]] [[1,y],[0,1]] (program.h) -----------
 * q0 += y q1;  q1 += g q0;  q0 += x q1 — twelve FMAs per pair, every one of them in place (no
 * temporaries, no register moves at the end of the op body), against sixteen plus ~30 moves for
 * the direct 2x2.  Coefficient block: y, g, x as (re, im), then two unused reals. */
template <typename real> struct Sh6;
template <> struct Sh6<double> {
    double yr, yi, gr, gi, xr, xi;
    __device__ __forceinline__ void load(const double *p) {
        const double2 *q = reinterpret_cast<const double2 *>(p);
        const double2 t0 = q[0], t1 = q[1], t2 = q[2];
        yr = t0.x, yi = t0.y, gr = t1.x, gi = t1.y, xr = t2.x, xi = t2.y;
    }
};
template <> struct Sh6<float> {
    u64 yr, yn, gr, gn, xr, xn; /* (re, re) and (-im, im) of y, g, x */
    __device__ __forceinline__ void load(const float *p) {
        const float4 *q = reinterpret_cast<const float4 *>(p);
        const float4 t0 = q[0], t1 = q[1], t2 = q[2]; /* MatLayout<float>: 4 real parts, then (-im, im) x 4 */
        yr = pk2(t0.x, t0.x), gr = pk2(t0.y, t0.y), xr = pk2(t0.z, t0.z);
        yn = pk2(t1.x, t1.y), gn = pk2(t1.z, t1.w), xn = pk2(t2.x, t2.y);
    }
};

__device__ __forceinline__ void pair_shear(double2 &q0, double2 &q1, const Sh6<double> &c) {
    q0.x = fma(c.yr, q1.x, q0.x);
    q0.y = fma(c.yr, q1.y, q0.y);
    q0.x = fma(-c.yi, q1.y, q0.x);
    q0.y = fma(c.yi, q1.x, q0.y);
    q1.x = fma(c.gr, q0.x, q1.x);
    q1.y = fma(c.gr, q0.y, q1.y);
    q1.x = fma(-c.gi, q0.y, q1.x);
    q1.y = fma(c.gi, q0.x, q1.y);
    q0.x = fma(c.xr, q1.x, q0.x);
    q0.y = fma(c.xr, q1.y, q0.y);
    q0.x = fma(-c.xi, q1.y, q0.x);
    q0.y = fma(c.xi, q1.x, q0.y);
}
__device__ __forceinline__ void pair_shear(float2 &x0, float2 &x1, const Sh6<float> &c) {
    u64 q0 = as_u64(x0), q1 = as_u64(x1);
    q0 = fma2(c.yr, q1, q0);
    q0 = fma2(c.yn, swap2(q1), q0);
    q1 = fma2(c.gr, q0, q1);
    q1 = fma2(c.gn, swap2(q0), q1);
    q0 = fma2(c.xr, q1, q0);
    q0 = fma2(c.xn, swap2(q1), q0);
    x0 = as_f2(q0);
    x1 = as_f2(q1);
}

/* MODE as lean_gen */
template <typename real, int K, int J, int MODE, int J2>
__device__ __forceinline__ void lean_shear(typename Cplx<real>::type (&a)[1 << K], const Sh6<real> &c, uint32_t regmask) {
#pragma unroll
    for (int r0 = 0; r0 < (1 << K); ++r0) {
        if (r0 & (1 << J)) continue;
        if (MODE == 2 && (r0 & (1 << J2))) continue;
        if (MODE == 3 && !(r0 & (1 << J2))) continue;
        if (MODE == 1 && !(regmask & (1u << r0))) continue;
        pair_shear(a[r0], a[r0 | (1 << J)], c);
    }
}

template <typename real, int K, int J, int J2>
__device__ __forceinline__ void lean_shear_regmux(typename Cplx<real>::type (&a)[1 << K], const Sh6<real> &c0,
                                                  const real *mats) {
    if (J == J2) return; /* never planned */
    Sh6<real> c1;
    c1.load(mats + MatLayout<real>::kStride); /* in flight under the first half */
    lean_shear<real, K, J, 2, (J == J2 ? 0 : J2)>(a, c0, 0u);
    lean_shear<real, K, J, 3, (J == J2 ? 0 : J2)>(a, c1, 0u);
}

/* the sheared ops of a pass (kept apart from lean_apply_op: their coefficient block is shorter) */
template <typename real, int K>
__device__ __forceinline__ void lean_apply_shear(typename Cplx<real>::type (&a)[1 << K], const Op<real> &op,
                                                 const real *mo, const real *mm) {
    Sh6<real> c;
    c.load(mm);
    switch (op.code) {
    case OPC_SHEAR(0): lean_shear<real, K, 0, 0, 0>(a, c, 0u); break;
    case OPC_SHEAR(1): lean_shear<real, K, QGB_J(1), 0, 0>(a, c, 0u); break;
    case OPC_SHEAR(2): lean_shear<real, K, QGB_J(2), 0, 0>(a, c, 0u); break;
    case OPC_SHEAR(3): if (K > 3) lean_shear<real, K, QGB_J(3), 0, 0>(a, c, 0u); break;
    case OPC_SHEAR_MASKED(0): lean_shear<real, K, 0, 1, 0>(a, c, op.regmask); break;
    case OPC_SHEAR_MASKED(1): lean_shear<real, K, QGB_J(1), 1, 0>(a, c, op.regmask); break;
    case OPC_SHEAR_MASKED(2): lean_shear<real, K, QGB_J(2), 1, 0>(a, c, op.regmask); break;
    case OPC_SHEAR_MASKED(3): if (K > 3) lean_shear<real, K, QGB_J(3), 1, 0>(a, c, op.regmask); break;
    case OPC_SHEAR_REGMUX(0, 1): lean_shear_regmux<real, K, 0, QGB_J(1)>(a, c, mo); break;
    case OPC_SHEAR_REGMUX(0, 2): lean_shear_regmux<real, K, 0, QGB_J(2)>(a, c, mo); break;
    case OPC_SHEAR_REGMUX(0, 3): if (K > 3) lean_shear_regmux<real, K, 0, QGB_J(3)>(a, c, mo); break;
    case OPC_SHEAR_REGMUX(1, 0): lean_shear_regmux<real, K, QGB_J(1), 0>(a, c, mo); break;
    case OPC_SHEAR_REGMUX(1, 2): lean_shear_regmux<real, K, QGB_J(1), QGB_J(2)>(a, c, mo); break;
    case OPC_SHEAR_REGMUX(1, 3): if (K > 3) lean_shear_regmux<real, K, QGB_J(1), QGB_J(3)>(a, c, mo); break;
    case OPC_SHEAR_REGMUX(2, 0): lean_shear_regmux<real, K, QGB_J(2), 0>(a, c, mo); break;
    case OPC_SHEAR_REGMUX(2, 1): lean_shear_regmux<real, K, QGB_J(2), QGB_J(1)>(a, c, mo); break;
    case OPC_SHEAR_REGMUX(2, 3): if (K > 3) lean_shear_regmux<real, K, QGB_J(2), QGB_J(3)>(a, c, mo); break;
    case OPC_SHEAR_REGMUX(3, 0): if (K > 3) lean_shear_regmux<real, K, QGB_J(3), 0>(a, c, mo); break;
    case OPC_SHEAR_REGMUX(3, 1): if (K > 3) lean_shear_regmux<real, K, QGB_J(3), QGB_J(1)>(a, c, mo); break;
    case OPC_SHEAR_REGMUX(3, 2): if (K > 3) lean_shear_regmux<real, K, QGB_J(3), QGB_J(2)>(a, c, mo); break;
    default: break;
    }
}

/* ---- OP_FAN: register r is multiplied by f * prod of reg[b] over the set bits b of r --------------
 * complex128: a Gray-code walk over the registers with a running factor (one complex product per
 * step, by reg[b] when bit b turns on, by its conjugate = inverse when it turns off), straight-line
 * code specialised on the hub predicate — no per-register branches, so no register moves at their
 * join points (the masked phase loop spends a third of its issue slots on them), and only four live
 * registers of factor state.  Measured against independent per-register factors from a 16-entry
 * constant table (32 more constant loads per fan and thread, 16% of them missing the constant
 * cache): the walk is 20% faster (profiles/r2m).  MODE 0: every register, 1: the registers of
 * regmask (uniform tests), 2: the registers whose bit J equals POL. */
template <int K, int MODE, int J, int POL, typename FanT>
__device__ __forceinline__ void fan_walk(double2 (&a)[1 << K], double fr, double fi, const FanT &fn, uint32_t regmask) {
    constexpr int NF = MODE == 2 ? K - 1 : K; /* free register bits */
#pragma unroll
    for (int i = 0; i < (1 << NF); ++i) {
        const int gray = i ^ (i >> 1);
        if (i > 0) {
            int fb = 0;
            while (!((i >> fb) & 1)) ++fb; /* the free bit that changes at this step */
            const int pb = (MODE == 2 && fb >= J) ? fb + 1 : fb;
            const double tr = fn.reg[pb][0], ti = ((gray >> fb) & 1) ? fn.reg[pb][1] : -fn.reg[pb][1];
            const double nr = fr * tr - fi * ti;
            fi = fr * ti + fi * tr;
            fr = nr;
        }
        int r = 0;
#pragma unroll
        for (int fb = 0; fb < NF; ++fb) {
            const int pb = (MODE == 2 && fb >= J) ? fb + 1 : fb;
            r |= ((gray >> fb) & 1) << pb;
        }
        if (MODE == 2) r |= POL << J;
        if (MODE != 1 || ((regmask >> r) & 1u)) {
            const double qr = a[r].x, qi = a[r].y;
            a[r].x = fr * qr - fi * qi;
            a[r].y = fr * qi + fi * qr;
        }
    }
}
template <int K, typename FanT>
__device__ __forceinline__ void lean_apply_fan(double2 (&a)[1 << K], int code, double fr, double fi, const FanT &fn,
                                               uint32_t regmask) {
    switch (code) {
    case OPC_FAN_ALL: fan_walk<K, 0, 0, 0>(a, fr, fi, fn, 0u); break;
    case OPC_FAN_REG(0, 0): fan_walk<K, 2, 0, 0>(a, fr, fi, fn, 0u); break;
    case OPC_FAN_REG(0, 1): fan_walk<K, 2, 0, 1>(a, fr, fi, fn, 0u); break;
    case OPC_FAN_REG(1, 0): fan_walk<K, 2, QGB_J(1), 0>(a, fr, fi, fn, 0u); break;
    case OPC_FAN_REG(1, 1): fan_walk<K, 2, QGB_J(1), 1>(a, fr, fi, fn, 0u); break;
    case OPC_FAN_REG(2, 0): fan_walk<K, 2, QGB_J(2), 0>(a, fr, fi, fn, 0u); break;
    case OPC_FAN_REG(2, 1): fan_walk<K, 2, QGB_J(2), 1>(a, fr, fi, fn, 0u); break;
    case OPC_FAN_REG(3, 0): if (K > 3) fan_walk<K, 2, QGB_J(3), 0>(a, fr, fi, fn, 0u); break;
    case OPC_FAN_REG(3, 1): if (K > 3) fan_walk<K, 2, QGB_J(3), 1>(a, fr, fi, fn, 0u); break;
    default: fan_walk<K, 1, 0, 0>(a, fr, fi, fn, regmask); break;
    }
}
/* complex64: the packed phase multiply is two instructions per amplitude, a per-register factor would
 * cost more in packing than it saves: the thread factor on the registers of regmask, then one phase
 * per register bit that carries a term */
template <int K, typename FanT>
__device__ __forceinline__ void lean_apply_fan(float2 (&a)[1 << K], int code, float fr, float fi, const FanT &fn,
                                               uint32_t regmask) {
    (void)code;
    lean_phase<K>(a, fr, fi, regmask);
    const uint32_t bit_set[4] = {0xaaaau, 0xccccu, 0xf0f0u, 0xff00u};
#pragma unroll
    for (int b = 0; b < K; ++b)
        if ((fn.n_reg >> b) & 1) lean_phase<K>(a, (float)fn.reg[b][0], (float)fn.reg[b][1], regmask & bit_set[b]);
}

/* One op on the registers of a thread that takes part in it (`mm` = the op's matrix or factor
 * already selected for this thread and tile, `mo` = the op's [m | m1] block).  The matrix is
 * fetched BEFORE the dispatch so the shared-memory latency hides under the switch. */
template <typename real, int K, bool DIRECT>
__device__ __forceinline__ void lean_apply_op(typename Cplx<real>::type (&a)[1 << K], const Op<real> &op,
                                              const real *mo, const real *mm) {
    Mat8<real> m;
    m.load(mm);
    if (DIRECT) { /* direct 2x2 bodies: compiled into the variant that serves passes with OP_GEN ops */
    switch (op.code) {
    case OPC_GEN(0): lean_gen<real, K, 0, 0, 0>(a, m, 0u); break;
    case OPC_GEN(1): lean_gen<real, K, QGB_J(1), 0, 0>(a, m, 0u); break;
    case OPC_GEN(2): lean_gen<real, K, QGB_J(2), 0, 0>(a, m, 0u); break;
    case OPC_GEN(3): if (K > 3) lean_gen<real, K, QGB_J(3), 0, 0>(a, m, 0u); break;
    case OPC_GEN_MASKED(0): lean_gen<real, K, 0, 1, 0>(a, m, op.regmask); break;
    case OPC_GEN_MASKED(1): lean_gen<real, K, QGB_J(1), 1, 0>(a, m, op.regmask); break;
    case OPC_GEN_MASKED(2): lean_gen<real, K, QGB_J(2), 1, 0>(a, m, op.regmask); break;
    case OPC_GEN_MASKED(3): if (K > 3) lean_gen<real, K, QGB_J(3), 1, 0>(a, m, op.regmask); break;
    case OPC_GEN_REGMUX(0, 1): lean_regmux<real, K, 0, QGB_J(1)>(a, m, mo); break;
    case OPC_GEN_REGMUX(0, 2): lean_regmux<real, K, 0, QGB_J(2)>(a, m, mo); break;
    case OPC_GEN_REGMUX(0, 3): if (K > 3) lean_regmux<real, K, 0, QGB_J(3)>(a, m, mo); break;
    case OPC_GEN_REGMUX(1, 0): lean_regmux<real, K, QGB_J(1), 0>(a, m, mo); break;
    case OPC_GEN_REGMUX(1, 2): lean_regmux<real, K, QGB_J(1), QGB_J(2)>(a, m, mo); break;
    case OPC_GEN_REGMUX(1, 3): if (K > 3) lean_regmux<real, K, QGB_J(1), QGB_J(3)>(a, m, mo); break;
    case OPC_GEN_REGMUX(2, 0): lean_regmux<real, K, QGB_J(2), 0>(a, m, mo); break;
    case OPC_GEN_REGMUX(2, 1): lean_regmux<real, K, QGB_J(2), QGB_J(1)>(a, m, mo); break;
    case OPC_GEN_REGMUX(2, 3): if (K > 3) lean_regmux<real, K, QGB_J(2), QGB_J(3)>(a, m, mo); break;
    case OPC_GEN_REGMUX(3, 0): if (K > 3) lean_regmux<real, K, QGB_J(3), 0>(a, m, mo); break;
    case OPC_GEN_REGMUX(3, 1): if (K > 3) lean_regmux<real, K, QGB_J(3), QGB_J(1)>(a, m, mo); break;
    case OPC_GEN_REGMUX(3, 2): if (K > 3) lean_regmux<real, K, QGB_J(3), QGB_J(2)>(a, m, mo); break;
    default: break;
    }
    }
    switch (op.code) {
    case OPC_SWAP(0): lean_swap<real, K, 0>(a, op.regmask); break;
    case OPC_SWAP(1): lean_swap<real, K, QGB_J(1)>(a, op.regmask); break;
    case OPC_SWAP(2): lean_swap<real, K, QGB_J(2)>(a, op.regmask); break;
    case OPC_SWAP(3): if (K > 3) lean_swap<real, K, QGB_J(3)>(a, op.regmask); break;
    case OPC_DIAG_REG: {
        const uint32_t regmask = op.regmask, regsel = op.regsel;
        real d0r, d0i, d1r, d1i;
        m.factor(0, d0r, d0i); /* (of the block the thread / tile parity selected: parity diagonals) */
        m.factor(1, d1r, d1i);
        lean_phase<K>(a, d0r, d0i, regmask & ~regsel);
        lean_phase<K>(a, d1r, d1i, regmask & regsel);
        break;
    }
    case OPC_DIAG_THR: {
        real dr, di;
        m.factor(0, dr, di); /* the selected block starts with the selected factor */
        lean_phase<K>(a, dr, di, op.regmask);
        break;
    }
    default: break;
    }
}

/* Shared memory: [NBUF tiles, 1024-byte aligned][2 NBUF mbarriers][matrices][per stage, per thread:
 * base slot].
 *
 * WS (warp specialised, option tma_ws=1, tiles of >= 32 threads): the CTA has one extra PRODUCER
 * warp that owns all TMA traffic — it stores a tile as soon as the consumers hand it over (`done`
 * mbarrier), waits for the store to drain the buffer and refills it — so no computing thread
 * waits for a store (thread 0 of the unspecialised kernel spends 27% of its time there, phase
 * timing build).  Consumers synchronise among themselves on a named barrier.  Measured 5-7%
 * SLOWER than the unspecialised kernel (the other resident CTAs already cover that wait, the
 * extra warp costs registers): kept as an option, off by default. */
template <typename real, int K, int NT, int MINB, int NBUF, bool WS, bool DIRECT, bool FANS>
__global__ void __launch_bounds__(NT, MINB)
tma_pass_kernel(const __grid_constant__ PassProgram<real> prog, const __grid_constant__ CUtensorMap tmap,
                const __grid_constant__ TmaGeometry geo) {
    typedef typename Cplx<real>::type cplx;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int T = prog.T;
    const uint32_t tile_bytes = (uint32_t)sizeof(cplx) << T;
    /* SWIZZLE_128B needs 1024-byte aligned tile buffers */
    unsigned char *tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full = reinterpret_cast<uint64_t *>(tiles + NBUF * tile_bytes);
    uint64_t *done = full + NBUF; /* WS: consumers -> producer, "tile finished, writes fenced" */
    constexpr int MS = MatLayout<real>::kStride;
    real *mats = reinterpret_cast<real *>(full + 2 * NBUF); /* 16-byte aligned: [op][m, m1][MS] */
    uint32_t *lut = reinterpret_cast<uint32_t *>(mats + 2 * MS * prog.n_ops);
    const uint32_t tid = threadIdx.x;
    const uint32_t nthr = WS ? blockDim.x - 32u : blockDim.x; /* computing threads == 2^(T-K) */
    const bool producer = WS && tid >= nthr;
    /* phase fans (OP_FAN): per fan two tables of thread factors — one indexed by the low FLO bits of
     * the thread number, one by the rest — and the factor of the tile being worked on / the next one */
    const int FLO = (T - K + 1) >> 1, FHI = T - K - FLO;
    const int fan_tab_size = (1 << FLO) + (1 << FHI);
    cplx *fan_tab = reinterpret_cast<cplx *>(lut + (size_t)prog.n_stages * nthr); /* (a multiple of 128 bytes in) */
    cplx *fan_out = fan_tab + prog.n_fans * fan_tab_size;                        /* [2][n_fans] */

    /* the pass's matrices, once per CTA */
    for (int i = tid; !producer && i < 2 * MS * prog.n_ops; i += nthr) {
        const int o = i / (2 * MS), w = i - o * (2 * MS);
        const Op<real> &op = prog.op[o];
        mats[i] = w >= MS ? MatLayout<real>::slot(op.m1, w - MS) : MatLayout<real>::slot(op.m, w);
    }

    /* per stage: byte offset (inside a tile buffer) of this thread's base element */
    for (int s = 0; !producer && s < prog.n_stages; ++s) {
        const Stage &st = prog.stage[s];
        uint32_t ebase = 0;
        for (int i = 0; i < T - K; ++i) ebase |= ((tid >> i) & 1u) << st.W[i];
        /* low 16 bits: element index (op predicates), high bits: swizzled byte offset / 8 */
        lut[s * nthr + tid] = ebase | ((swz<real>(ebase) * (uint32_t)sizeof(cplx) / 8u) << 16);
    }
    /* Which ops this thread takes part in (thread-bit controls) and where it takes the second
     * matrix / factor (multiplexer or diagonal target on a thread bit): one bit per op, fixed
     * for the whole pass because a thread's elements do not depend on the tile. */
    uint32_t act = 0, sel_thr = 0; /* n_ops <= 32 */
    for (int s = 0; !producer && s < prog.n_stages; ++s) {
        const Stage &st = prog.stage[s];
        uint32_t ebase = 0;
        for (int i = 0; i < T - K; ++i) ebase |= ((tid >> i) & 1u) << st.W[i];
        for (int o = st.op_begin; o < st.op_end; ++o) {
            const Op<real> &op = prog.op[o];
            if ((ebase & op.cmt) == op.cmt) act |= 1u << o;
            if (__popc(ebase & op.tsel) & 1) sel_thr |= 1u << o; /* one bit, or the parity of several */
        }
    }
    const uint64_t n_tiles = 1ull << (prog.n_lanes - T);
    const uint64_t stride = gridDim.x;
    /* state-vector index of the origin of tile number t */
    auto tile_base = [&](uint64_t t) {
        uint64_t base = 0;
#pragma unroll
        for (int d = 0; d < QGB_MAX_GROUPS; ++d)
            base |= (uint64_t)(((uint32_t)(t >> geo.shift[d])) & geo.mask[d]) << geo.base_shift[d];
        return base;
    };
    /* fan f's factor for tile number t: its terms on lanes outside the tile, from the two tables
     * fan_tile_table_kernel built for this pass (one entry per value of the low / high tile-number bits) */
    auto fan_tile_factor = [&](int f, uint64_t t) {
        cplx out;
        out.x = (real)1, out.y = (real)0;
        if (geo.fan_tiles) {
            const cplx *tab = reinterpret_cast<const cplx *>(geo.fan_tiles) + (size_t)f * ((1u << geo.fan_lo) + (1u << geo.fan_hi));
            const cplx lo = tab[t & ((1ull << geo.fan_lo) - 1ull)], hi = tab[(1ull << geo.fan_lo) + (t >> geo.fan_lo)];
            out.x = lo.x * hi.x - lo.y * hi.y;
            out.y = lo.x * hi.y + lo.y * hi.x;
        }
        return out;
    };
    for (int i = tid; FANS && !producer && i < prog.n_fans * fan_tab_size; i += nthr) {
        const int f = i / fan_tab_size, e = i - f * fan_tab_size;
        const uint32_t x = e < (1 << FLO) ? (uint32_t)e : ((uint32_t)(e - (1 << FLO)) << FLO); /* thread-number bits */
        const auto &fn = prog.fan[f];
        double pr = 1., pi = 0.;
        if (e < (1 << FLO)) pr = fn.base[0], pi = fn.base[1]; /* (the factor common to the whole fan) */
        for (int k = fn.first; k < fn.first + fn.n_thr; ++k) {
            const auto &tm = prog.fan_term[k];
            if ((x >> tm.bit) & 1u) {
                const double r = pr * tm.re - pi * tm.im;
                pi = pr * tm.im + pi * tm.re;
                pr = r;
            }
        }
        fan_tab[i].x = (real)pr, fan_tab[i].y = (real)pi;
    }
    if (FANS && !producer && (int)tid < prog.n_fans && blockIdx.x < n_tiles) fan_out[tid] = fan_tile_factor((int)tid, blockIdx.x);
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(&full[b], 1);
            mbar_init(&done[b], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    /* tile number -> tensor coordinates of the tile's origin */
    auto issue_load = [&](uint64_t t, int b) {
        int c[QGB_MAX_GROUPS];
#pragma unroll
        for (int d = 0; d < QGB_MAX_GROUPS; ++d)
            c[d] = (int)((((uint32_t)(t >> geo.shift[d])) & geo.mask[d]) << geo.tbits[d]);
        mbar_expect_tx(&full[b], tile_bytes);
        if (geo.l2_hint & 1)
            tma_load_5d_hint(tiles + (size_t)b * tile_bytes, &tmap, &full[b], c[0], c[1], c[2], c[3], l2_evict_first_policy());
        else
            tma_load_5d(tiles + (size_t)b * tile_bytes, &tmap, &full[b], c[0], c[1], c[2], c[3]);
    };
    auto issue_store = [&](uint64_t t, int b) {
        int c[QGB_MAX_GROUPS];
#pragma unroll
        for (int d = 0; d < QGB_MAX_GROUPS; ++d)
            c[d] = (int)((((uint32_t)(t >> geo.shift[d])) & geo.mask[d]) << geo.tbits[d]);
        if (geo.l2_hint & 2)
            tma_store_5d_hint(&tmap, tiles + (size_t)b * tile_bytes, c[0], c[1], c[2], c[3], l2_evict_first_policy());
        else
            tma_store_5d(&tmap, tiles + (size_t)b * tile_bytes, c[0], c[1], c[2], c[3]);
        bulk_commit();
    };

    if (WS) {
        if (producer) {
            if (tid == nthr) { /* one lane drives the TMA unit */
                uint64_t t_load = blockIdx.x;
                for (int j = 0; j < NBUF && t_load < n_tiles && !geo.zero_input; ++j, t_load += stride) issue_load(t_load, j);
                for (int j = 0; j < NBUF && geo.zero_input; ++j) mbar_arrive(&full[j]); /* nothing to load: the buffers are free */
                int pb = 0;
                uint32_t pparity = 0;
                for (uint64_t t = blockIdx.x; t < n_tiles; t += stride) {
                    mbar_wait(&done[pb], pparity); /* the consumers are through with tile t */
                    issue_store(t, pb);
                    if (geo.zero_input) {
                        bulk_wait_read<0>(); /* the store has drained the buffer: the consumers may refill it */
                        mbar_arrive(&full[pb]);
                    } else if (t_load < n_tiles) {
                        bulk_wait_read<0>(); /* the store has drained the buffer: refill it */
                        issue_load(t_load, pb);
                        t_load += stride;
                    }
                    if (++pb == NBUF) {
                        pb = 0;
                        pparity ^= 1u;
                    }
                }
                bulk_wait_read<0>(); /* shared memory stays alive until the last store has read it */
            }
            return;
        }
    } else {
        /* prologue: the first tile */
        if (tid == 0 && blockIdx.x < n_tiles && !geo.zero_input) issue_load(blockIdx.x, 0);
    }

    int b = 0;            /* buffer of the tile being worked on */
    uint32_t parity = 0;  /* phase of full[b] this tile completes */
    int fan_cur = 0;      /* half of fan_out that holds this tile's factors */
#ifdef QGB_PHASE_TIMING
    long long ph_wait = 0, ph_load = 0, ph_ops = 0, ph_store = 0, ph_tail = 0, ph_t;
#define PH_MARK(acc) do { const long long now_ = clock64(); acc += now_ - ph_t; ph_t = now_; } while (0)
    ph_t = clock64();
#else
#define PH_MARK(acc) do { } while (0)
#endif
    for (uint64_t t = blockIdx.x; t < n_tiles; t += stride) {
        /* prefetch the next tile into buffer b + 1.  That buffer held the tile stored NBUF - 1
         * iterations ago, so at most the newest NBUF - 2 stores may still be reading shared
         * memory (NBUF == 2: wait for the store just issued; NBUF == 3: load, stages and store
         * of three consecutive tiles overlap without a wait) */
        if (!WS && tid == 0) {
            const uint64_t tn = t + stride;
            if (tn < n_tiles) {
                bulk_wait_read<NBUF - 2>();
                if (!geo.zero_input) issue_load(tn, b + 1 == NBUF ? 0 : b + 1);
            }
        }
        /* index of the tile's origin -> the ops this tile skips (controls outside the tile) and
         * the ops that take their second matrix / factor (multiplexer or diagonal target
         * outside the tile): CTA-uniform bit masks over the ops */
        const uint64_t base = tile_base(t);
        /* the fans' factors of the NEXT tile, one thread per fan: published by the barrier that ends
         * this tile (the buffer written here was last read in the previous tile) */
        if (FANS && (int)tid < prog.n_fans && t + stride < n_tiles)
            fan_out[(fan_cur ^ 1) * prog.n_fans + tid] = fan_tile_factor((int)tid, t + stride);
        uint32_t eff = act, sel = sel_thr;
        for (int i = 0; i < prog.n_out; ++i) {
            const uint64_t cm = prog.out[i].ctrl_mask;
            const int o = prog.out[i].op;
            if ((base & cm) != cm) eff &= ~(1u << o);
            if (__popcll(base & prog.out[i].sel_mask) & 1) sel ^= 1u << o;
        }

        PH_MARK(ph_tail);
        if (WS || !geo.zero_input) mbar_wait(&full[b], parity);
        PH_MARK(ph_wait);
        unsigned char *buf = tiles + (size_t)b * tile_bytes;
        bool fresh = geo.zero_input != 0; /* the tile is |0...0>'s: zeros, amplitude 0 of the state = 1 */

        for (int s = 0; s < prog.n_stages && geo.debug_mode < 2; ++s) {
            const Stage &st = prog.stage[s];
            if (st.op_begin == st.op_end) continue;
            const uint32_t packed = lut[s * nthr + tid];
            const uint32_t sbyte = (packed >> 16) << 3;

            /* slot of register r: base slot ^ XOR of the per-bit constants over the bits of r.  The
             * constants are CTA-uniform, the slots are recomputed at the store (a table of 2^K
             * slots kept across the ops costs 2^K registers and spilled) */
            const uint32_t x0 = st.xb[0], x1 = st.xb[1], x2 = st.xb[2], x3 = K > 3 ? st.xb[3] : 0u;
#define QGB_SLOT(base, r) ((base) ^ (((r) & 1) ? x0 : 0u) ^ (((r) & 2) ? x1 : 0u) ^ (((r) & 4) ? x2 : 0u) ^ (((r) & 8) ? x3 : 0u))
            cplx a[1 << K];
#pragma unroll
            for (int r = 0; r < (1 << K); ++r) a[r] = *reinterpret_cast<const cplx *>(buf + QGB_SLOT(sbyte, r));
            if (fresh) { /* (uniform) the first stage of a tile of a state that was never written */
#pragma unroll
                for (int r = 0; r < (1 << K); ++r) a[r].x = (real)0, a[r].y = (real)0;
                if (t == 0 && tid == 0) a[0].x = (real)1; /* thread 0 holds tile element 0 in register 0 */
                fresh = false;
            }
            PH_MARK(ph_load);

            if (geo.debug_mode == 0) {
                uint32_t bit = 1u << st.op_begin;
                const real *mo = mats + 2 * MS * st.op_begin;
                for (int o = st.op_begin; o < st.op_end; ++o, bit <<= 1, mo += 2 * MS) {
                    if (!(eff & bit)) continue;
                    const Op<real> &op = prog.op[o];
                    const real *mm = mo + ((sel & bit) ? MS : 0);
                    if (FANS && op.kind == OP_FAN) {
                        /* one factor per thread and tile (thread tables x tile table), then the register
                         * bits' factors, on the registers the hub predicate admits */
                        const cplx *tab = fan_tab + op.bit * fan_tab_size;
                        const cplx f0 = tab[tid & ((1u << FLO) - 1u)], f1 = tab[(1 << FLO) + (tid >> FLO)];
                        const cplx f2 = fan_out[fan_cur * prog.n_fans + op.bit];
                        const real gr = f0.x * f1.x - f0.y * f1.y, gi = f0.x * f1.y + f0.y * f1.x;
                        lean_apply_fan<K>(a, op.code, gr * f2.x - gi * f2.y, gr * f2.y + gi * f2.x, prog.fan[op.bit], op.regmask);
                    } else if (op.code >= OPC_SHEAR(0)) {
                        lean_apply_shear<real, K>(a, op, mo, mm);
                    } else {
                        lean_apply_op<real, K, DIRECT>(a, op, mo, mm);
                    }
                }
            }

            PH_MARK(ph_ops);
            /* registers relabelled by the stage's shears go to the slots they now stand for */
            const uint32_t sstore = sbyte ^ st.store_xor;
#pragma unroll
            for (int r = 0; r < (1 << K); ++r) *reinterpret_cast<cplx *>(buf + QGB_SLOT(sstore, r)) = a[r];
#undef QGB_SLOT
            if (s + 1 < prog.n_stages) {
                if (st.warp_local)
                    __syncwarp(); /* the next stage reads only what this warp wrote (planner.cpp) */
                else if (WS)
                    named_sync(nthr);
                else
                    __syncthreads();
            }
            PH_MARK(ph_store);
        }
        /* generic-proxy writes -> visible to the async proxy, then the tile goes back to HBM */
        fence_proxy_async();
        if (WS) {
            named_sync(nthr);
            if (tid == 0) mbar_arrive(&done[b]); /* the producer warp stores it */
        } else {
            __syncthreads();
            if (tid == 0) issue_store(t, b);
        }

        if (++b == NBUF) {
            b = 0;
            parity ^= 1u;
        }
        if (FANS) fan_cur ^= 1;
    }
    if (!WS && tid == 0) bulk_wait_read<0>();
#ifdef QGB_PHASE_TIMING
    PH_MARK(ph_tail);
    if (tid == 0 && geo.phase) {
        atomicAdd(&geo.phase[0], (unsigned long long)ph_wait);
        atomicAdd(&geo.phase[1], (unsigned long long)ph_load);
        atomicAdd(&geo.phase[2], (unsigned long long)ph_ops);
        atomicAdd(&geo.phase[3], (unsigned long long)ph_store);
        atomicAdd(&geo.phase[4], (unsigned long long)ph_tail);
    }
#endif
}

/* Tile factors of the pass's phase fans.  Tile number t addresses the lanes outside the tile (bit i of
 * t <-> PassProgram::rest_lane[i]); a fan's factor for a tile is the product of its outside terms whose
 * lane is 1 in the tile's origin, and splits into a factor of the low `lo` bits of t and one of the
 * rest: out[f][e] for e < 2^lo, out[f][2^lo + e] for the high part.  One thread per entry; the term
 * loop is uniform (constant-bank broadcasts), products in double, rounded once. */
template <typename real>
__global__ void __launch_bounds__(256)
fan_tile_table_kernel(const __grid_constant__ PassProgram<real> prog, int lo, int hi, typename Cplx<real>::type *out) {
    const uint32_t per_fan = (1u << lo) + (1u << hi);
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per_fan * (uint32_t)prog.n_fans) return;
    const int f = (int)(idx / per_fan);
    const uint32_t e = idx - (uint32_t)f * per_fan;
    const bool high = e >= (1u << lo);
    const uint32_t bits = high ? e - (1u << lo) : e;
    const int first = high ? lo : 0, count = high ? hi : lo;
    uint64_t base = 0;
    for (int i = 0; i < count; ++i) base |= (uint64_t)((bits >> i) & 1u) << prog.rest_lane[first + i];
    const auto &fn = prog.fan[f];
    double pr = 1., pi = 0.;
    for (int k = fn.first + fn.n_thr; k < fn.first + fn.n_thr + fn.n_out; ++k) {
        const auto &tm = prog.fan_term[k];
        if ((base >> tm.bit) & 1ull) {
            const double r = pr * tm.re - pi * tm.im;
            pi = pr * tm.im + pi * tm.re;
            pr = r;
        }
    }
    out[idx].x = (real)pr;
    out[idx].y = (real)pi;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

} // namespace

/* state shared by the two halves of this file (one definition, in the QGB_TMA_SHARED half) */
namespace tma_state {
extern void *g_fan_tiles; /* scratch of the tile-factor tables (passes run in stream order) */
extern size_t g_fan_tiles_bytes;
extern void *g_encode_fn;
extern unsigned long long *g_phase;
extern int g_tma_ws; /* measured 5-7% slower than the unspecialised kernel on B200 (profiles/r1z): off */
extern int g_tma_l2_hint;
extern int g_tma_debug_mode;
extern int g_tma_sm_count;
extern int g_tma_max_smem;
#ifdef QGB_TMA_SHARED
void *g_fan_tiles = nullptr;
size_t g_fan_tiles_bytes = 0;
void *g_encode_fn = nullptr;
unsigned long long *g_phase = nullptr;
int g_tma_ws = 0;
int g_tma_l2_hint = 0;
int g_tma_debug_mode = 0;
int g_tma_sm_count = 148;
int g_tma_max_smem = 48 * 1024;
#endif
} // namespace tma_state
using namespace tma_state;

namespace {

cudaError_t resolve_encode() {
    if (g_encode_fn) return cudaSuccess;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t rc = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (rc != cudaSuccess) return rc;
    if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
    g_encode_fn = fn;
    return cudaSuccess;
}

template <typename real>
cudaError_t encode_map(const PassProgram<real> &prog, void *amp, CUtensorMap *map, TmaGeometry *geo) {
    const uint64_t elem = 2 * sizeof(real);            /* bytes of one amplitude */
    const uint64_t scalars_per_row = 128 / sizeof(real); /* tensor elements are scalars (re, im separate) */
    cuuint64_t gdim[5];
    cuuint64_t gstride[4];
    cuuint32_t box[5], estride[5] = {1, 1, 1, 1, 1};
    gdim[0] = scalars_per_row;
    box[0] = (cuuint32_t)scalars_per_row;
    const uint64_t total_bytes = elem << prog.n_lanes;
    int consumed = 0;
    geo->n_groups = prog.n_groups;
    geo->phase = g_phase;
    geo->l2_hint = g_tma_l2_hint;
    geo->debug_mode = g_tma_debug_mode;
    geo->fan_tiles = nullptr;
    geo->fan_lo = geo->fan_hi = 0;
    geo->zero_input = 0;
    for (int d = 0; d < QGB_MAX_GROUPS; ++d) {
        if (d < prog.n_groups) {
            const int s = prog.grp_start[d], t = prog.grp_t[d], r = prog.grp_r[d];
            gdim[d + 1] = 1ull << (t + r);
            box[d + 1] = 1u << t;
            gstride[d] = elem << s;
            geo->shift[d] = consumed;
            geo->mask[d] = r >= 32 ? 0xffffffffu : ((1u << r) - 1u);
            geo->tbits[d] = t;
            geo->base_shift[d] = s + t;
            consumed += r;
        } else {
            /* padding dimension of extent 1 (keeps one rank-5 instruction for every tile shape) */
            gdim[d + 1] = 1;
            box[d + 1] = 1;
            gstride[d] = total_bytes < (1ull << 39) ? (total_bytes >= 128 ? total_bytes : 128) : (1ull << 39);
            geo->shift[d] = 0;
            geo->mask[d] = 0;
            geo->tbits[d] = 0;
            geo->base_shift[d] = 0;
        }
    }
    if (consumed != prog.n_lanes - prog.T) return cudaErrorInvalidValue;
    const CUtensorMapDataType dt = sizeof(real) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult res = reinterpret_cast<EncodeTiledFn>(g_encode_fn)(map, dt, 5, amp, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return res == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <typename real, int K, int NT, int MINB, int NBUF, bool WS, bool DIRECT, bool FANS>
cudaError_t launch_tma_variant2(const PassProgram<real> &prog, const CUtensorMap &map, const TmaGeometry &geo,
                                size_t smem, cudaStream_t stream) {
    static int configured = 0;
    auto kernel = tma_pass_kernel<real, K, NT, MINB, NBUF, WS, DIRECT, FANS>;
    if (!configured) {
        cudaError_t rc = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_tma_max_smem);
        if (rc != cudaSuccess) return rc;
        configured = 1;
    }
    const unsigned nthr = (1u << (prog.T - K)) + (WS ? 32u : 0u); /* + the producer warp */
    int per_sm = 0;
    cudaError_t rc = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, (int)nthr, smem);
    if (rc != cudaSuccess) return rc;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    const uint64_t n_tiles = 1ull << (prog.n_lanes - prog.T);
    const uint64_t resident = (uint64_t)g_tma_sm_count * per_sm; /* persistent: one wave */
    const unsigned nblocks = (unsigned)(n_tiles < resident ? n_tiles : resident);
    kernel<<<nblocks, nthr, smem, stream>>>(prog, map, geo);
    return cudaGetLastError();
}

/* passes without direct 2x2 ops (every dense gate sheared) run the variant compiled without those bodies */
template <typename real, int K, int NT, int MINB, int NBUF, bool WS>
cudaError_t launch_tma_variant(const PassProgram<real> &prog, const CUtensorMap &map, const TmaGeometry &geo,
                               size_t smem, cudaStream_t stream) {
    /* (and the fan bodies only into the one that serves passes with phase fans) */
    if (prog.n_fans > 0 && prog.n_direct == 0) return launch_tma_variant2<real, K, NT, MINB, NBUF, WS, false, true>(prog, map, geo, smem, stream);
    if (prog.n_fans > 0) return launch_tma_variant2<real, K, NT, MINB, NBUF, WS, true, true>(prog, map, geo, smem, stream);
    if (prog.n_direct == 0) return launch_tma_variant2<real, K, NT, MINB, NBUF, WS, false, false>(prog, map, geo, smem, stream);
    return launch_tma_variant2<real, K, NT, MINB, NBUF, WS, true, false>(prog, map, geo, smem, stream);
}

template <typename real, int K>
cudaError_t launch_tma_by_shape(const PassProgram<real> &prog, void *amp, int prec, int n_buf, int min_ctas,
                                cudaStream_t stream, int zero_input) {
    cudaError_t rc = resolve_encode();
    if (rc != cudaSuccess) return rc;
    CUtensorMap map;
    TmaGeometry geo;
    rc = encode_map<real>(prog, amp, &map, &geo);
    if (rc != cudaSuccess) return rc;
    geo.zero_input = zero_input ? 1 : 0;
    {
        int n_out_terms = 0;
        for (int f = 0; f < prog.n_fans; ++f) n_out_terms += prog.fan[f].n_out;
        if (n_out_terms > 0) {
            typedef typename Cplx<real>::type cplx;
            const int nb = prog.n_lanes - prog.T;
            geo.fan_lo = (nb + 1) / 2;
            geo.fan_hi = nb - geo.fan_lo;
            const size_t entries = (size_t)prog.n_fans * (((size_t)1 << geo.fan_lo) + ((size_t)1 << geo.fan_hi));
            if (entries * sizeof(cplx) > g_fan_tiles_bytes) {
                /* (grows a few times per process at most; freeing synchronises with the passes in flight) */
                if (g_fan_tiles) cudaFree(g_fan_tiles);
                g_fan_tiles = nullptr;
                g_fan_tiles_bytes = 0;
                const size_t want = entries * sizeof(cplx) * 2;
                rc = cudaMalloc(&g_fan_tiles, want);
                if (rc != cudaSuccess) return rc;
                g_fan_tiles_bytes = want;
            }
            fan_tile_table_kernel<real><<<(unsigned)((entries + 255) / 256), 256, 0, stream>>>(
                prog, geo.fan_lo, geo.fan_hi, reinterpret_cast<cplx *>(g_fan_tiles));
            rc = cudaGetLastError();
            if (rc != cudaSuccess) return rc;
            geo.fan_tiles = g_fan_tiles;
        }
    }
    const size_t smem = tma_pass_smem_bytes(prec, prog.T, prog.K, prog.n_stages, n_buf, prog.n_ops, prog.n_fans);
    const int nthr = 1 << (prog.T - prog.K);
    (void)min_ctas;
    if (!g_tma_ws) { /* A/B switch: the unspecialised kernel for every shape */
        if (n_buf >= 3) {
            if (nthr <= 256) return launch_tma_variant<real, K, 256, 2, 3, false>(prog, map, geo, smem, stream);
            return launch_tma_variant<real, K, 1024, 1, 3, false>(prog, map, geo, smem, stream);
        }
        if (nthr <= 256) return launch_tma_variant<real, K, 256, 2, 2, false>(prog, map, geo, smem, stream);
        return launch_tma_variant<real, K, 1024, 1, 2, false>(prog, map, geo, smem, stream);
    }
    /* tiles of 32..512 computing threads: warp-specialised (one more warp per CTA) */
    if (n_buf >= 3) {
        if (nthr < 32) return launch_tma_variant<real, K, 256, 2, 3, false>(prog, map, geo, smem, stream);
        if (nthr <= 128) return launch_tma_variant<real, K, 160, 3, 3, true>(prog, map, geo, smem, stream);
        if (nthr <= 256) return launch_tma_variant<real, K, 288, 2, 3, true>(prog, map, geo, smem, stream);
        if (nthr <= 512) return launch_tma_variant<real, K, 544, 1, 3, true>(prog, map, geo, smem, stream);
        return launch_tma_variant<real, K, 1024, 1, 3, false>(prog, map, geo, smem, stream);
    }
    if (nthr < 32) return launch_tma_variant<real, K, 256, 2, 2, false>(prog, map, geo, smem, stream);
    if (nthr <= 128) return launch_tma_variant<real, K, 160, 3, 2, true>(prog, map, geo, smem, stream);
    if (nthr <= 256) return launch_tma_variant<real, K, 288, 2, 2, true>(prog, map, geo, smem, stream);
    if (nthr <= 512) return launch_tma_variant<real, K, 544, 1, 2, true>(prog, map, geo, smem, stream);
    return launch_tma_variant<real, K, 1024, 1, 2, false>(prog, map, geo, smem, stream);
}

} // namespace

#ifdef QGB_TMA_SHARED
size_t tma_pass_smem_bytes(int prec, int T, int K, int n_stages, int n_buf, int n_ops, int n_fans) {
    const size_t elem = prec == 1 ? 16 : 8;
    size_t tiles = n_buf * (elem << T);
    tiles = (tiles + 1023) & ~(size_t)1023;
    /* phase fans: two tables of thread factors and two tile factors each */
    const int flo = (T - K + 1) >> 1, fhi = T - K - flo;
    const size_t fans = (size_t)n_fans * elem * ((size_t)(1 << flo) + (size_t)(1 << fhi) + 2);
    /* + mbarriers, the matrices of up to QGB_MAX_OPS ops, the per-stage thread table, alignment slack */
    return tiles + 16 * n_buf + (prec == 1 ? 128 : 96) * (size_t)n_ops + sizeof(uint32_t) * ((size_t)n_stages << (T - K)) + fans + 1024;
}

void tma_pass_set_warp_specialised(int on) { g_tma_ws = on ? 1 : 0; }
void tma_pass_set_l2_hint(int mode) { g_tma_l2_hint = mode & 3; }
void tma_pass_set_debug_mode(int mode) { g_tma_debug_mode = mode; }

/* Diagnostic (compile with -DQGB_PHASE_TIMING): where warp 0 of every CTA spends its cycles. */
void tma_pass_phase_report() {
#ifdef QGB_PHASE_TIMING
    if (!g_phase) return;
    unsigned long long h[5];
    cudaMemcpy(h, g_phase, sizeof(h), cudaMemcpyDeviceToHost);
    const double tot = (double)(h[0] + h[1] + h[2] + h[3] + h[4]);
    if (tot > 0)
        std::fprintf(stderr, "[qgb phase] wait-for-tile %.1f%%  stage loads %.1f%%  ops %.1f%%  stage stores+barrier %.1f%%  fence/store/issue %.1f%%\n",
                     100. * h[0] / tot, 100. * h[1] / tot, 100. * h[2] / tot, 100. * h[3] / tot, 100. * h[4] / tot);
    cudaMemset(g_phase, 0, sizeof(h));
#endif
}

void tma_pass_release() {
    if (g_fan_tiles) cudaFree(g_fan_tiles);
    g_fan_tiles = nullptr;
    g_fan_tiles_bytes = 0;
}

cudaError_t tma_pass_configure(int max_smem_optin, int sm_count) {
#ifdef QGB_PHASE_TIMING
    if (!g_phase) {
        cudaMalloc(&g_phase, 5 * sizeof(unsigned long long));
        cudaMemset(g_phase, 0, 5 * sizeof(unsigned long long));
    }
#endif
    if (sm_count > 0) g_tma_sm_count = sm_count;
    g_tma_max_smem = max_smem_optin;
    return cudaSuccess;
}
#endif /* QGB_TMA_SHARED */

#ifdef QGB_TMA_F64
template <>
cudaError_t launch_tma_pass<double>(const PassProgram<double> &prog, void *amp, int n_buf, int min_ctas,
                                    cudaStream_t stream, int zero_input) {
    if ((prog.K != 3 && prog.K != 4) || prog.T < prog.K || prog.T - prog.K > 10 || prog.n_groups < 1 || prog.n_ops > 32)
        return cudaErrorInvalidValue;
    /* 4 register bits: 16 complex128 per thread, CTAs of half as many threads (up to 128 registers) */
    if (prog.K == 4) return launch_tma_by_shape<double, 4>(prog, amp, 1, n_buf >= 3 ? 3 : 2, 2, stream, zero_input);
    return launch_tma_by_shape<double, 3>(prog, amp, 1, n_buf >= 3 ? 3 : 2, min_ctas, stream, zero_input);
}
#endif

#ifdef QGB_TMA_F32
template <>
cudaError_t launch_tma_pass<float>(const PassProgram<float> &prog, void *amp, int n_buf, int min_ctas,
                                   cudaStream_t stream, int zero_input) {
    if (prog.K != QGB_K32 || prog.T < prog.K || prog.T - prog.K > 10 || prog.n_groups < 1 || prog.n_ops > 32)
        return cudaErrorInvalidValue;
    return launch_tma_by_shape<float, QGB_K32>(prog, amp, 2, n_buf >= 3 ? 3 : 2, min_ctas, stream, zero_input);
}
#endif

} // namespace qgb
