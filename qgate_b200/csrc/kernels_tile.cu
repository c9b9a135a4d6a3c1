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
#include "tile_ops.cuh"

namespace qgb {

namespace {

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
