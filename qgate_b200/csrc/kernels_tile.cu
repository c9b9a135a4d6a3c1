/*
 * kernels_tile.cu — gate application kernels (sm_100a).
 *
 *   tile_pass_kernel   the fused pass: one CTA stages a 2^T-amplitude tile in shared
 *                      memory, applies every gate of the pass, writes the tile back.
 *                      One read + one write of the state vector per PASS, not per gate.
 *   simple_gate_kernel one (multi-)controlled 2x2 per launch, one amplitude pair per
 *                      thread; used for state vectors too small for a tile.
 *
 * Arithmetic being reproduced: qgate/simulator/src/CPUQubitProcessor.cpp:307-362
 * (o0 = m00 q0 + m01 q1, o1 = m10 q0 + m11 q1 on every pair whose control bits are 1),
 * matrix cast to the state precision first (CPUQubitProcessor.cpp:312).  The incumbent
 * device code this replaces is DeviceProcPrimitives.cu:198-249 (one pair per thread, one
 * launch per gate, 12 KB of lookup tables per controlled gate).
 *
 * Shared-memory layout: tile element e lives at byte address (e * sizeof(complex)) with
 * bits [6:4] XORed by bits [9:7] — the 128-byte swizzle TMA tensor maps produce — so that
 * any choice of register bits leaves the threads of a quarter/half warp on distinct banks
 * (the planner orders the thread bits accordingly, planner.cpp: choose_thread_bits).
 */
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

/* ---- the fused pass ------------------------------------------------------------------
 * Persistent CTAs: the grid is (#SMs x resident CTAs) and every CTA walks over tiles
 * blockIdx.x, blockIdx.x + gridDim.x, ...  What depends only on the pass — the global
 * offsets of the tile's contiguous runs and, per stage and thread, the thread's base
 * element — is computed once per CTA into shared
 * memory (profiles/r1b: recomputing it per stage per tile was 37% of all instructions). */

template <typename real, int K, int NT, int MINB, int NBUF>
__global__ void __launch_bounds__(NT, MINB)
tile_pass_kernel(const __grid_constant__ PassProgram<real> prog,
                 typename Cplx<real>::type *__restrict__ amp) {
    typedef typename Cplx<real>::type cplx;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int T = prog.T, L = prog.L;
    cplx *tile = reinterpret_cast<cplx *>(smem_raw); /* NBUF tile buffers */
    uint64_t *choff = reinterpret_cast<uint64_t *>(smem_raw + NBUF * (sizeof(cplx) << T));
    uint16_t *lut = reinterpret_cast<uint16_t *>(choff + (1u << (T - L)));
    const uint32_t tid = threadIdx.x, nthr = blockDim.x; /* nthr == 2^(T-K) */

    /* global offset of each run of 2^L contiguous amplitudes of the tile */
    {
        const uint32_t nchunks = 1u << (T - L);
        for (uint32_t c = tid; c < nchunks; c += nthr) {
            uint64_t off = 0;
            for (int b = 0; b < T - L; ++b) off |= (uint64_t)((c >> b) & 1u) << prog.tile_lane[L + b];
            choff[c] = off;
        }
    }
    /* per stage: this thread's base element */
    for (int s = 0; s < prog.n_stages; ++s) {
        const Stage &st = prog.stage[s];
        uint32_t ebase = 0;
        for (int i = 0; i < T - K; ++i) ebase |= ((tid >> i) & 1u) << st.W[i];
        lut[s * nthr + tid] = (uint16_t)ebase;
    }
    __syncthreads();
    const uint32_t lowmask = (1u << L) - 1u;
    const int nrest = prog.n_lanes - T;
    const uint64_t n_tiles = 1ull << nrest;

    /* index of a tile's origin: tile number bits deposited on the non-tile lanes */
    auto tile_base = [&](uint64_t t) {
        uint64_t base = 0;
        for (int i = 0; i < nrest; ++i) base |= ((t >> i) & 1ull) << prog.rest_lane[i];
        return base;
    };
    /* global -> shared without register staging (cp.async, 16 bytes per request, coalesced
     * along the low lanes); the copy of tile i+1 runs under the stages of tile i */
    auto prefetch = [&](uint64_t t, cplx *dst) {
        const uint64_t base = tile_base(t);
        if (sizeof(real) == 8) {
#pragma unroll
            for (int i = 0; i < (1 << K); ++i) {
                const uint32_t e = tid + i * nthr;
                const uint64_t g = base | choff[e >> L] | (e & lowmask);
                cp_async16(&dst[swz<real>(e)], &amp[g]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < (1 << K) / 2; ++i) {
                const uint32_t e = 2u * (tid + i * nthr);
                const uint64_t g = base | choff[e >> L] | (e & lowmask);
                cp_async16(&dst[swz<real>(e)], &amp[g]);
            }
        }
    };

    cplx *buf = tile;                     /* tile being worked on */
    cplx *other = tile + (NBUF == 2 ? (1u << T) : 0u); /* tile in flight (NBUF == 2) */
    if (NBUF == 2) {
        if (blockIdx.x < n_tiles) prefetch(blockIdx.x, buf);
        cp_async_commit();
    }

    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        if (NBUF == 2) {
            if (t + gridDim.x < n_tiles) prefetch(t + gridDim.x, other);
            cp_async_commit();
            cp_async_wait<1>(); /* everything but the newest group has landed: this tile is here */
        } else {
            prefetch(t, buf);   /* one buffer: the other resident CTAs cover the latency */
            cp_async_commit();
            cp_async_wait<0>();
        }
        __syncthreads();
        const uint64_t base = tile_base(t);

        for (int s = 0; s < prog.n_stages; ++s) {
            const Stage &st = prog.stage[s];
            if (st.op_begin == st.op_end) continue;
            const uint32_t ebase = lut[s * nthr + tid];
            const uint32_t sbase = swz<real>(ebase);

            cplx a[1 << K];
#pragma unroll
            for (int r = 0; r < (1 << K); ++r) a[r] = buf[sbase ^ st.sro[r]];

            for (int o = st.op_begin; o < st.op_end; ++o) apply_op<real, K>(a, prog.op[o], base, ebase);

#pragma unroll
            for (int r = 0; r < (1 << K); ++r) buf[sbase ^ st.sro[r]] = a[r];
            __syncthreads();
        }

        /* shared -> global */
        if (sizeof(real) == 8) {
#pragma unroll
            for (int i = 0; i < (1 << K); ++i) {
                const uint32_t e = tid + i * nthr;
                const uint64_t g = base | choff[e >> L] | (e & lowmask);
                __stcs(&amp[g], buf[swz<real>(e)]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < (1 << K) / 2; ++i) {
                const uint32_t e = 2u * (tid + i * nthr);
                const uint64_t g = base | choff[e >> L] | (e & lowmask);
                __stcs(reinterpret_cast<float4 *>(&amp[g]),
                       *reinterpret_cast<const float4 *>(&buf[swz<real>(e)]));
            }
        }
        __syncthreads(); /* the next prefetch into this buffer overwrites what these stores read */
        if (NBUF == 2) {
            cplx *tmp = buf;
            buf = other;
            other = tmp;
        }
    }
    cp_async_wait<0>();
}

/* ---- one gate per launch (small state vectors) ------------------------------------------ */

template <typename real>
struct Mat2 {
    real m[8];
};

template <typename real>
__global__ void __launch_bounds__(256)
simple_gate_kernel(typename Cplx<real>::type *__restrict__ amp, uint64_t n_pairs, SortedBits skip,
                   uint64_t target_bit, uint64_t ctrl_mask, Mat2<real> mat) {
    /* `skip` lists the target, the controls (deposited as 1 through ctrl_mask) and the lanes
     * that must be 0 (a multiplexed gate's low branch): those stay 0 after the insertion */
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_pairs) return;
    /* insert a 0 at every skipped position (target and controls), ascending */
    uint64_t idx = gid;
    for (int i = 0; i < skip.n; ++i) {
        const uint64_t lo = idx & ((1ull << skip.pos[i]) - 1ull);
        idx = ((idx - lo) << 1) | lo;
    }
    const uint64_t i0 = idx | ctrl_mask, i1 = i0 | target_bit;
    const typename Cplx<real>::type q0 = amp[i0], q1 = amp[i1];
    const real *m = mat.m;
    typename Cplx<real>::type o0, o1;
    o0.x = m[0] * q0.x - m[1] * q0.y + m[2] * q1.x - m[3] * q1.y;
    o0.y = m[0] * q0.y + m[1] * q0.x + m[2] * q1.y + m[3] * q1.x;
    o1.x = m[4] * q0.x - m[5] * q0.y + m[6] * q1.x - m[7] * q1.y;
    o1.y = m[4] * q0.y + m[5] * q0.x + m[6] * q1.y + m[7] * q1.x;
    amp[i0] = o0;
    amp[i1] = o1;
}

int g_sm_count = 148;
int g_max_smem = 48 * 1024;

/* Launch bounds: every variant is capped at 64 registers per thread (NT x MINB = 1024 threads per
 * SM).  With more registers available ptxas preloads the 2x2 matrices into vector registers
 * (LDC + moves) instead of reading them from uniform registers (profiles/r1d). */
template <typename real, int K, int NT, int MINB, int NBUF>
cudaError_t launch_variant(const PassProgram<real> &prog, void *amp, size_t smem, cudaStream_t stream) {
    static int configured = 0;
    auto kernel = tile_pass_kernel<real, K, NT, MINB, NBUF>;
    if (!configured) {
        cudaError_t rc = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem);
        if (rc != cudaSuccess) return rc;
        configured = 1;
    }
    const unsigned nthr = 1u << (prog.T - K);
    int per_sm = 0;
    cudaError_t rc = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, (int)nthr, smem);
    if (rc != cudaSuccess) return rc;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    const uint64_t n_tiles = 1ull << (prog.n_lanes - prog.T);
    const uint64_t resident = (uint64_t)g_sm_count * per_sm; /* persistent: one wave */
    const unsigned nblocks = (unsigned)(n_tiles < resident ? n_tiles : resident);
    kernel<<<nblocks, nthr, smem, stream>>>(prog, reinterpret_cast<typename Cplx<real>::type *>(amp));
    return cudaGetLastError();
}

template <typename real, int K>
cudaError_t launch_by_shape(const PassProgram<real> &prog, void *amp, int prec, int n_buf, cudaStream_t stream) {
    const size_t smem = tile_pass_smem_bytes(prec, prog.T, prog.L, prog.n_stages, n_buf);
    const int nthr = 1 << (prog.T - prog.K);
    if (n_buf == 2) {
        if (nthr <= 256) return launch_variant<real, K, 256, 4, 2>(prog, amp, smem, stream);
        if (nthr <= 512) return launch_variant<real, K, 512, 2, 2>(prog, amp, smem, stream);
        return launch_variant<real, K, 1024, 1, 2>(prog, amp, smem, stream);
    }
    if (nthr <= 256) return launch_variant<real, K, 256, 4, 1>(prog, amp, smem, stream);
    if (nthr <= 512) return launch_variant<real, K, 512, 2, 1>(prog, amp, smem, stream);
    return launch_variant<real, K, 1024, 1, 1>(prog, amp, smem, stream);
}

} // namespace

size_t tile_pass_smem_bytes(int prec, int T, int L, int n_stages, int n_buf) {
    const size_t elem = prec == 1 ? 16 : 8;
    const int K = prec == 1 ? QGB_K64 : QGB_K32;
    return n_buf * (elem << T) + (sizeof(uint64_t) << (T - L)) +
           sizeof(uint16_t) * ((size_t)n_stages << (T - K));
}

cudaError_t tile_pass_configure(int max_smem_optin, int sm_count) {
    if (sm_count > 0) g_sm_count = sm_count;
    g_max_smem = max_smem_optin;
    return cudaSuccess;
}

template <>
cudaError_t launch_tile_pass<double>(const PassProgram<double> &prog, void *amp, int n_buf, cudaStream_t stream) {
    if (prog.K != QGB_K64 || prog.T < prog.K || prog.T - prog.K > 10) return cudaErrorInvalidValue;
    return launch_by_shape<double, QGB_K64>(prog, amp, 1, n_buf == 2 ? 2 : 1, stream);
}

template <>
cudaError_t launch_tile_pass<float>(const PassProgram<float> &prog, void *amp, int n_buf, cudaStream_t stream) {
    if (prog.K != QGB_K32 || prog.T < prog.K || prog.T - prog.K > 10 || prog.L < 1)
        return cudaErrorInvalidValue;
    return launch_by_shape<float, QGB_K32>(prog, amp, 2, n_buf == 2 ? 2 : 1, stream);
}

cudaError_t launch_simple_gate(int prec, void *amp, int n_lanes, const double *mat8, int target,
                               uint64_t ctrl_mask, uint64_t zero_mask, cudaStream_t stream) {
    SortedBits skip;
    skip.n = 0;
    const uint64_t touched = ctrl_mask | zero_mask | (1ull << target);
    for (int lane = 0; lane < n_lanes; ++lane)
        if (touched & (1ull << lane)) skip.pos[skip.n++] = (int8_t)lane;
    const uint64_t n_pairs = 1ull << (n_lanes - skip.n);
    const unsigned nthr = 256;
    const unsigned nblocks = (unsigned)((n_pairs + nthr - 1) / nthr);
    if (prec == 1) {
        Mat2<double> m;
        for (int i = 0; i < 8; ++i) m.m[i] = mat8[i];
        simple_gate_kernel<double><<<nblocks, nthr, 0, stream>>>(
            reinterpret_cast<double2 *>(amp), n_pairs, skip, 1ull << target, ctrl_mask, m);
    } else {
        Mat2<float> m;
        for (int i = 0; i < 8; ++i) m.m[i] = (float)mat8[i];
        simple_gate_kernel<float><<<nblocks, nthr, 0, stream>>>(
            reinterpret_cast<float2 *>(amp), n_pairs, skip, 1ull << target, ctrl_mask, m);
    }
    return cudaGetLastError();
}

} // namespace qgb
