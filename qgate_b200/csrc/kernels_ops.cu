/*
 * kernels_ops.cu — everything on the hot path that is not a gate (sm_100a):
 * probability reduction, measurement collapse, separate / reset / join, amplitude and
 * probability readout, marginal probabilities, sampling-pool scan and search.
 *
 * Each kernel names the reference routine whose result it reproduces.  Where the reference
 * CPU code does a short fixed sequence of multiplications and additions (join products,
 * readout products, |a|^2) the kernels use non-contracted arithmetic (__fmul_rn / __dmul_rn
 * ...) in the same order, so those values are bit-identical to the reference's; reductions
 * accumulate in double with a fixed tree (deterministic, but a different order from the
 * reference's per-worker partial sums, CPUQubitProcessor.cpp:96-123).
 */
#include <cuda_runtime.h>

#include "kernels.h"

namespace qgb {

namespace {

template <typename real> struct Cplx;
template <> struct Cplx<float> { typedef float2 type; };
template <> struct Cplx<double> { typedef double2 type; };

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

/* std::complex product as the reference's compiler evaluates it: (ac - bd, ad + bc) */
template <typename C> __device__ __forceinline__ C cmul_exact(C a, C b) {
    C o;
    o.x = sub_rn(mul_rn(a.x, b.x), mul_rn(a.y, b.y));
    o.y = add_rn(mul_rn(a.x, b.y), mul_rn(a.y, b.x));
    return o;
}
template <typename C> __device__ __forceinline__ auto abs2_exact(C a) -> decltype(a.x) {
    return add_rn(mul_rn(a.x, a.x), mul_rn(a.y, a.y));
}

__device__ __forceinline__ uint64_t insert_zero(uint64_t idx, int pos) {
    const uint64_t lo = idx & ((1ull << pos) - 1ull);
    return ((idx - lo) << 1) | lo;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

/* sum over the block, fixed tree; valid in thread 0 */
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double warp_part[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads(); /* protect warp_part from a previous use */
    if (lane == 0) warp_part[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? warp_part[threadIdx.x] : 0.;
    if (w == 0) v = warp_sum(v);
    return v;
}

/* ---- one (multi-)controlled 2x2 per launch, one amplitude pair per thread ------------------
 * The path of state vectors too small for the fused pass (dynamic qubit grouping creates many)
 * and of the options fuse = 0 / exact = 1.  EXACT: the arithmetic of CPUQubitProcessor.cpp:316-324
 * operation by operation — std::complex products (ac - bd, ad + bc) rounded term by term, then the
 * complex sum, nothing contracted into FMAs — so the amplitudes are bit-identical to the
 * reference CPU runtime's. */
template <typename real> struct Mat2 { real m[8]; };

template <typename real, bool EXACT>
__global__ void __launch_bounds__(256)
simple_gate_kernel(typename Cplx<real>::type *__restrict__ amp, uint64_t n_pairs, const SortedBits skip,
                   uint64_t target_bit, uint64_t ctrl_mask, const Mat2<real> mat) {
    typedef typename Cplx<real>::type cplx;
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_pairs) return;
    uint64_t idx = gid; /* a 0 at every skipped position (target, controls, zero-controls), ascending */
    for (int i = 0; i < skip.n; ++i) idx = insert_zero(idx, skip.pos[i]);
    const uint64_t i0 = idx | ctrl_mask, i1 = i0 | target_bit;
    const cplx q0 = amp[i0], q1 = amp[i1];
    const real *m = mat.m;
    cplx o0, o1;
    if (EXACT) {
        cplx m00, m01, m10, m11;
        m00.x = m[0], m00.y = m[1], m01.x = m[2], m01.y = m[3];
        m10.x = m[4], m10.y = m[5], m11.x = m[6], m11.y = m[7];
        const cplx a0 = cmul_exact(m00, q0), b0 = cmul_exact(m01, q1);
        const cplx a1 = cmul_exact(m10, q0), b1 = cmul_exact(m11, q1);
        o0.x = add_rn(a0.x, b0.x), o0.y = add_rn(a0.y, b0.y);
        o1.x = add_rn(a1.x, b1.x), o1.y = add_rn(a1.y, b1.y);
    } else {
        o0.x = m[0] * q0.x - m[1] * q0.y + m[2] * q1.x - m[3] * q1.y;
        o0.y = m[0] * q0.y + m[1] * q0.x + m[2] * q1.y + m[3] * q1.x;
        o1.x = m[4] * q0.x - m[5] * q0.y + m[6] * q1.x - m[7] * q1.y;
        o1.y = m[4] * q0.y + m[5] * q0.x + m[6] * q1.y + m[7] * q1.x;
    }
    amp[i0] = o0;
    amp[i1] = o1;
}

/* parity diagonal on a state too small for the fused pass: a *= d0 (even number of `parity` lanes set)
 * or d1 (odd), where the control bits are set */
template <typename real>
__global__ void __launch_bounds__(256)
simple_parity_kernel(typename Cplx<real>::type *__restrict__ amp, uint64_t n_amps, uint64_t parity, uint64_t ctrl_mask,
                     real d0r, real d0i, real d1r, real d1i) {
    typedef typename Cplx<real>::type cplx;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_amps || (i & ctrl_mask) != ctrl_mask) return;
    cplx d;
    const bool odd = __popcll(i & parity) & 1;
    d.x = odd ? d1r : d0r;
    d.y = odd ? d1i : d0i;
    amp[i] = cmul_exact(d, amp[i]);
}

/* ---- reset to a basis state (CPUQubitProcessor.cpp:61-73) ----------------------------- */
template <typename real>
__global__ void set_one_kernel(typename Cplx<real>::type *amp, uint64_t one_at) {
    amp[one_at].x = (real)1;
    amp[one_at].y = (real)0;
}

/* ---- P(lane == 0) (CPUQubitProcessor.cpp:75-125; incumbent DeviceSum.cuh:7-39) ----------- */
template <typename real>
__global__ void __launch_bounds__(256)
prob0_kernel(const typename Cplx<real>::type *__restrict__ amp, uint64_t n_half, int lane,
             double *__restrict__ partials) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t lowmask = (1ull << lane) - 1ull;
    double acc0 = 0., acc1 = 0., acc2 = 0., acc3 = 0.;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n_half; i += 4 * stride) {
        typename Cplx<real>::type v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t j = i + u * stride;
            v[u] = amp[((j & ~lowmask) << 1) | (j & lowmask)];
        }
        acc0 += (double)v[0].x * v[0].x + (double)v[0].y * v[0].y;
        acc1 += (double)v[1].x * v[1].x + (double)v[1].y * v[1].y;
        acc2 += (double)v[2].x * v[2].x + (double)v[2].y * v[2].y;
        acc3 += (double)v[3].x * v[3].x + (double)v[3].y * v[3].y;
    }
    for (; i < n_half; i += stride) {
        const typename Cplx<real>::type v = amp[((i & ~lowmask) << 1) | (i & lowmask)];
        acc0 += (double)v.x * v.x + (double)v.y * v.y;
    }
    const double s = block_sum((acc0 + acc1) + (acc2 + acc3));
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__global__ void __launch_bounds__(1024)
final_sum_kernel(const double *__restrict__ partials, int n, double *__restrict__ result) {
    double acc = 0.;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];
    const double s = block_sum(acc);
    if (threadIdx.x == 0) *result = s;
}

/* ---- measurement collapse (CPUQubitProcessor.cpp:217-246) --------------------------------- */
template <typename real>
__global__ void __launch_bounds__(256)
decohere_kernel(typename Cplx<real>::type *__restrict__ amp, uint64_t n_amps, int lane, int value,
                real norm) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride) {
        typename Cplx<real>::type v = amp[i];
        if ((int)((i >> lane) & 1ull) == value) {
            v.x = mul_rn(v.x, norm);
            v.y = mul_rn(v.y, norm);
        } else {
            v.x = (real)0;
            v.y = (real)0;
        }
        amp[i] = v;
    }
}

/* ---- collapse + remove the lane (CPUQubitProcessor.cpp:248-284) ---------------------------- */
template <typename real>
__global__ void __launch_bounds__(256)
decohere_separate_kernel(typename Cplx<real>::type *__restrict__ dst,
                         const typename Cplx<real>::type *__restrict__ src, uint64_t n_dst, int lane,
                         int value, real norm) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t bit = (uint64_t)value << lane;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_dst; i += stride) {
        typename Cplx<real>::type v = src[insert_zero(i, lane) | bit];
        v.x = mul_rn(norm, v.x);
        v.y = mul_rn(norm, v.y);
        dst[i] = v;
    }
}

/* ---- reset of a lane measured as 1 (CPUQubitProcessor.cpp:286-305) -------------------------- */
template <typename real>
__global__ void __launch_bounds__(256)
apply_reset_kernel(typename Cplx<real>::type *__restrict__ amp, uint64_t n_half, int lane) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    typename Cplx<real>::type zero;
    zero.x = (real)0;
    zero.y = (real)0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_half; i += stride) {
        const uint64_t i0 = insert_zero(i, lane), i1 = i0 | (1ull << lane);
        amp[i0] = amp[i1];
        amp[i1] = zero;
    }
}

/* ---- join = Kronecker product + zero padding (CPUQubitProcessor.cpp:129-213) --------------- */
template <typename real>
__global__ void __launch_bounds__(256)
join_kernel(typename Cplx<real>::type *__restrict__ dst, uint64_t n_dst, uint64_t n_product,
            uint64_t index_offset, const __grid_constant__ JoinParams jp) {
    typedef typename Cplx<real>::type cplx;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; d < n_dst; d += stride) {
        const uint64_t i = index_offset + d; /* index in the whole product (shards: offset != 0) */
        cplx v;
        if (i < n_product) {
            const int last = jp.n_src - 1;
            v = reinterpret_cast<const cplx *>(jp.src[last])[(i >> jp.shift[last]) &
                                                             ((1ull << jp.n_lanes[last]) - 1ull)];
            /* the reference multiplies from the low-lane end: s_k * (s_{k+1} * (...)) */
            for (int k = last - 1; k >= 0; --k) {
                const cplx s = reinterpret_cast<const cplx *>(jp.src[k])[(i >> jp.shift[k]) &
                                                                         ((1ull << jp.n_lanes[k]) - 1ull)];
                v = cmul_exact(s, v);
            }
        } else {
            v.x = (real)0;
            v.y = (real)0;
        }
        dst[d] = v;
    }
}

/* ---- amplitude / probability readout (CPUQubitsStatesGetter.cpp:34-157) ------------------- */
__device__ __forceinline__ uint64_t ext_to_local(const LaneTable &t, uint64_t ext) {
    uint64_t local = 0;
    for (int l = 0; l < t.n_lanes; ++l) local |= ((ext >> t.ext[l]) & 1ull) << l;
    return local;
}

template <typename real, int MATHOP>
__global__ void __launch_bounds__(256)
get_states_kernel(void *__restrict__ out, const __grid_constant__ GatherParams gp, uint64_t empty_mask,
                  int64_t first, int64_t count, int64_t start, int64_t step) {
    typedef typename Cplx<real>::type cplx;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride) {
        const uint64_t ext = (uint64_t)(start + step * (first + j));
        const bool empty = (ext & empty_mask) != 0;
        if (MATHOP == 0) {
            cplx v;
            v.x = empty ? (real)0 : (real)1;
            v.y = (real)0;
            if (!empty)
                for (int q = 0; q < gp.n_qstates; ++q) {
                    const cplx s = reinterpret_cast<const cplx *>(gp.qs[q].amp)[ext_to_local(gp.qs[q], ext)];
                    v = cmul_exact(v, s);
                }
            reinterpret_cast<cplx *>(out)[j] = v;
        } else {
            real v = empty ? (real)0 : (real)1;
            if (!empty)
                for (int q = 0; q < gp.n_qstates; ++q) {
                    const cplx s = reinterpret_cast<const cplx *>(gp.qs[q].amp)[ext_to_local(gp.qs[q], ext)];
                    v = mul_rn(v, abs2_exact(s));
                }
            reinterpret_cast<real *>(out)[j] = v;
        }
    }
}

/* ---- marginal probabilities, hidden lanes in the LSBs (CPUQubitsStatesGetter.cpp:159-254) -- */
/* ACC = double: the engine's default (hidden lanes summed in double).  ACC = real: the reference
 * CPU runtime's own accumulation, value for value (CPUQubitsStatesGetter.cpp:186-201: a running
 * sum in `real`, started at 0, over 2^n_hidden consecutive indices) — option pool_compat_workers. */
template <typename real, typename ACC>
__global__ void __launch_bounds__(256)
prob_array_kernel(double *__restrict__ out, const __grid_constant__ GatherParams gp, int n_hidden,
                  int64_t first, int64_t count) {
    typedef typename Cplx<real>::type cplx;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint64_t n_sum = 1ull << n_hidden;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride) {
        const uint64_t d = (uint64_t)(first + j);
        ACC sum = (ACC)0;
        for (uint64_t h = 0; h < n_sum; ++h) {
            const uint64_t ext = (d << n_hidden) | h;
            real v = (real)1;
            for (int q = 0; q < gp.n_qstates; ++q) {
                const cplx s = reinterpret_cast<const cplx *>(gp.qs[q].amp)[ext_to_local(gp.qs[q], ext)];
                v = mul_rn(v, abs2_exact(s));
            }
            sum = add_rn(sum, (ACC)v);
        }
        out[j] = (double)sum;
    }
}

/* out[j] = sum of 2^log2_group consecutive inputs (sequential, in ACC: double, or the state
 * precision of the reference's second and later reduction levels, CPUQubitsStatesGetter.cpp:224-231) */
template <typename ACC>
__global__ void __launch_bounds__(256)
reduce_groups_kernel(double *__restrict__ out, const double *__restrict__ in, int log2_group, int64_t count) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t g = 1ll << log2_group;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride) {
        ACC sum = (ACC)0;
        for (int64_t h = 0; h < g; ++h) sum = add_rn(sum, (ACC)in[j * g + h]);
        out[j] = (double)sum;
    }
}

/* ---- sampling pool, reference-compatible scan (CPUSamplingPool.cpp:13-47) -------------------
 * The reference scans sequentially, in the state precision V, one span per worker thread
 * (Parallel.h:12-23: spans of ceil(N / W) rounded up to 16), prefix-sums the W span totals in V,
 * then adds the offset of the span and multiplies by V(1) / total.  A floating-point running sum
 * cannot be re-associated without changing its roundings, so this mode keeps the order: ONE warp
 * per span, the lanes move 32 x 16 values at a time between HBM and shared memory (coalesced),
 * lane 0 runs the sum.  Values are kept as doubles that hold V values exactly.  Used only with
 * option pool_compat_workers = W (index-exact parity with a W-worker reference); the default
 * pool is the parallel float64 scan below. */
#define COMPAT_CHUNK 2048
template <typename V>
__global__ void __launch_bounds__(32)
compat_scan_kernel(double *__restrict__ prob, int64_t n, int64_t span, double *__restrict__ span_total) {
    __shared__ double buf[COMPAT_CHUNK];
    const int64_t begin = span * blockIdx.x < n ? span * blockIdx.x : n;
    const int64_t end = begin + span < n ? begin + span : n;
    const int lane = threadIdx.x;
    V v = (V)0;
    for (int64_t base = begin; base < end; base += COMPAT_CHUNK) {
        const int64_t left = end - base;
        const int len = left < COMPAT_CHUNK ? (int)left : COMPAT_CHUNK;
        for (int i = lane; i < len; i += 32) buf[i] = prob[base + i];
        __syncwarp();
        if (lane == 0) {
#pragma unroll 8
            for (int i = 0; i < len; ++i) {
                v = add_rn(v, (V)buf[i]);
                buf[i] = (double)v;
            }
        }
        __syncwarp();
        for (int i = lane; i < len; i += 32) prob[base + i] = buf[i];
        __syncwarp();
    }
    if (lane == 0) span_total[blockIdx.x] = (double)v;
}

/* prob[idx] = (prob[idx] + offset of its span) * norm, in V; span 0 is only multiplied */
template <typename V>
__global__ void __launch_bounds__(256)
compat_apply_kernel(double *__restrict__ prob, int64_t n, int64_t span, const double *__restrict__ span_offset,
                    double norm) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t t = i / span;
        V p = (V)prob[i];
        if (t > 0) p = add_rn(p, (V)span_offset[t - 1]);
        prob[i] = (double)mul_rn(p, (V)norm);
    }
}

template <typename real>
__global__ void __launch_bounds__(256)
cast_from_double_kernel(real *__restrict__ out, const double *__restrict__ in, int64_t count) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride)
        out[j] = (real)in[j];
}

/* ---- sampling pool: inclusive scan (CPUSamplingPool.cpp:8-63) ------------------------------
 * Three phases over blocks of 4096 entries, all sums in double with a FIXED combination tree
 * (deterministic): (1) block sums, (2) exclusive scan of the block sums + grand total, (3) the
 * inclusive scan inside every block, + block offset, x 1 / total (the reference's order: scan, then
 * multiply, CPUSamplingPool.cpp:36).  Phases 1 and 3 share block_scan, so a block's total is the
 * same number in both and the cumulative array is monotonic across block boundaries.
 *
 * The source of phases 1 and 3 is either the marginal probability vector (doubles) or — pools over
 * ALL lanes of one state vector in index order, the 30-qubit Grover case — the amplitudes
 * themselves: |a|^2 is recomputed in both phases and the 2^n-entry probability vector is never
 * written (40 instead of 56 bytes of HBM traffic per complex128 amplitude).
 *
 * Memory access: a block moves its 4096 entries with coalesced loads through shared memory (entry
 * i at slot i + i / 16: the 16 consecutive entries of a thread then lie in distinct banks). */
#define SCAN_THREADS 256
#define SCAN_PER_THREAD 16
#define SCAN_BLOCK (SCAN_THREADS * SCAN_PER_THREAD)
#define SCAN_SLOT(i) ((i) + ((i) >> 4))
#define SCAN_SMEM (SCAN_BLOCK + SCAN_BLOCK / 16)

struct SrcProb {
    const double *p;
    __device__ __forceinline__ double at(int64_t i) const { return p[i]; }
};
template <typename real> struct SrcAmp {
    const typename Cplx<real>::type *a;
    /* the value prob_array_kernel stores for one qstates without hidden lanes */
    __device__ __forceinline__ double at(int64_t i) const { return (double)abs2_exact(a[i]); }
};

/* exclusive prefix of `v` over the block in thread order + the block total; the same shuffle tree
 * whatever the data (fixed order of additions).  `part` holds 8 doubles of shared memory. */
__device__ __forceinline__ double block_scan(double v, double *part, double &total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    __syncthreads(); /* protect `part` from a previous use */
    if (lane == 31) part[w] = incl;
    __syncthreads();
    double before = 0., all = 0.;
#pragma unroll
    for (int k = 0; k < SCAN_THREADS / 32; ++k) {
        const double pk = part[k];
        if (k < w) before += pk;
        all += pk;
    }
    total = all;
    return before + (incl - v);
}

/* the block's entries -> shared memory (coalesced), then this thread's 16 consecutive ones */
template <typename SRC>
__device__ __forceinline__ void scan_load(const SRC &src, int64_t n, double *stage, double (&v)[SCAN_PER_THREAD]) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK;
#pragma unroll
    for (int u = 0; u < SCAN_PER_THREAD; ++u) {
        const int i = u * SCAN_THREADS + (int)threadIdx.x;
        stage[SCAN_SLOT(i)] = base + i < n ? src.at(base + i) : 0.;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < SCAN_PER_THREAD; ++u) v[u] = stage[SCAN_SLOT((int)threadIdx.x * SCAN_PER_THREAD + u)];
}

template <typename SRC>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_phase1_kernel(const SRC src, int64_t n, double *__restrict__ block_sums) {
    __shared__ double stage[SCAN_SMEM];
    __shared__ double part[SCAN_THREADS / 32];
    double v[SCAN_PER_THREAD];
    scan_load(src, n, stage, v);
    double acc = 0.;
#pragma unroll
    for (int u = 0; u < SCAN_PER_THREAD; ++u) acc += v[u];
    double total;
    block_scan(acc, part, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

/* exclusive scan of the block sums in place (sequential per thread chunk, block_scan over the
 * chunk totals: deterministic), grand total to *total */
__global__ void __launch_bounds__(SCAN_THREADS)
scan_phase2_kernel(double *__restrict__ block_sums, int64_t n_blocks, double *__restrict__ total) {
    __shared__ double part[SCAN_THREADS / 32];
    const int64_t per = (n_blocks + blockDim.x - 1) / blockDim.x;
    const int64_t begin = (int64_t)threadIdx.x * per;
    const int64_t end = begin + per < n_blocks ? begin + per : n_blocks;
    double acc = 0.;
    for (int64_t i = begin; i < end; ++i) acc += block_sums[i];
    double all;
    acc = block_scan(acc, part, all);
    if (threadIdx.x == 0) *total = all;
    for (int64_t i = begin; i < end; ++i) {
        const double c = block_sums[i];
        block_sums[i] = acc;
        acc += c;
    }
}

template <typename SRC>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_phase3_kernel(const SRC src, double *__restrict__ cum, int64_t n, const double *__restrict__ block_sums,
                   double global_offset, double norm) {
    __shared__ double stage[SCAN_SMEM];
    __shared__ double part[SCAN_THREADS / 32];
    double v[SCAN_PER_THREAD];
    scan_load(src, n, stage, v);
    double acc = 0.;
#pragma unroll
    for (int u = 0; u < SCAN_PER_THREAD; ++u) {
        acc += v[u];
        v[u] = acc;
    }
    double total;
    const double before = block_scan(acc, part, total);
    /* global_offset: total of the lower ranks' shards (0 on one GPU, which keeps the sum exact);
     * norm = 1 / sum, then cum *= norm (CPUSamplingPool.cpp:36) */
    const double offset = global_offset + (block_sums[blockIdx.x] + before);
    __syncthreads(); /* every thread has read its entries: the staging area carries the results out */
#pragma unroll
    for (int u = 0; u < SCAN_PER_THREAD; ++u)
        stage[SCAN_SLOT((int)threadIdx.x * SCAN_PER_THREAD + u)] = (offset + v[u]) * norm;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK;
#pragma unroll
    for (int u = 0; u < SCAN_PER_THREAD; ++u) {
        const int i = u * SCAN_THREADS + (int)threadIdx.x;
        if (base + i < n) cum[base + i] = stage[SCAN_SLOT(i)];
    }
}

/* ---- sampling pool: upper_bound search + empty-lane deposit (CPUSamplingPool.cpp:71-81) ---- */
__global__ void __launch_bounds__(256)
sample_kernel(const double *__restrict__ cum, int64_t n_states, const double *__restrict__ rnd,
              int64_t *__restrict__ obs, int n_samples, SortedBits empty, int fp32) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_samples) return;
    double r = rnd[i];
    if (fp32) r = (double)(float)r;
    int64_t lo = 0, hi = n_states;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (cum[mid] > r)
            hi = mid;
        else
            lo = mid + 1;
    }
    if (lo > n_states - 1) lo = n_states - 1;
    uint64_t v = (uint64_t)lo;
    for (int k = 0; k < empty.n; ++k) v = insert_zero(v, empty.pos[k]);
    obs[i] = (int64_t)v;
}

/* ---- Simulator.sample for terminal measurements (simulator.py:85-119) ------
 * The reference re-runs the whole circuit per shot; its Measure ops draw one uniform number each and
 * answer 0 if r < P(qubit = 0 | earlier outcomes of the shot).  With all measurements at the end,
 * those conditional probabilities are ratios of partial sums of ONE marginal distribution, i.e.
 * differences of the pool's cumulative array: shot s walks the binary tree over the n_bits pool
 * lanes from the most significant one (the first measured qubit) down, one draw per level. */
__global__ void __launch_bounds__(256)
sample_sequential_kernel(const double *__restrict__ cum, int n_bits, const double *__restrict__ rnd,
                         int64_t *__restrict__ obs, int n_shots) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_shots) return;
    const double *r = rnd + (int64_t)s * n_bits;
    int64_t prefix = 0;
    for (int k = 0; k < n_bits; ++k) {
        const int rest = n_bits - k; /* undecided lanes */
        const int64_t lo = prefix << rest, mid = lo + ((int64_t)1 << (rest - 1)), hi = lo + ((int64_t)1 << rest);
        const double c_lo = lo ? cum[lo - 1] : 0., c_mid = cum[mid - 1], c_hi = cum[hi - 1];
        const double all = c_hi - c_lo;
        const double p0 = all > 0. ? (c_mid - c_lo) / all : 1.;
        prefix = (prefix << 1) | (r[k] < p0 ? 0 : 1);
    }
    obs[s] = prefix;
}

inline unsigned grid_for(uint64_t n, unsigned nthr, unsigned cap) {
    uint64_t b = (n + nthr - 1) / nthr;
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return (unsigned)b;
}

/* grid-stride kernels: enough CTAs to fill 148 SMs several times over */
const unsigned kStreamCap = 148 * 16;

} // namespace

cudaError_t launch_simple_gate(int prec, void *amp, int n_lanes, const double *mat8, int target,
                               uint64_t ctrl_mask, uint64_t zero_mask, bool exact, cudaStream_t stream) {
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
        if (exact)
            simple_gate_kernel<double, true><<<nblocks, nthr, 0, stream>>>(reinterpret_cast<double2 *>(amp), n_pairs, skip,
                                                                          1ull << target, ctrl_mask, m);
        else
            simple_gate_kernel<double, false><<<nblocks, nthr, 0, stream>>>(reinterpret_cast<double2 *>(amp), n_pairs, skip,
                                                                           1ull << target, ctrl_mask, m);
    } else {
        Mat2<float> m;
        for (int i = 0; i < 8; ++i) m.m[i] = (float)mat8[i];
        if (exact)
            simple_gate_kernel<float, true><<<nblocks, nthr, 0, stream>>>(reinterpret_cast<float2 *>(amp), n_pairs, skip,
                                                                         1ull << target, ctrl_mask, m);
        else
            simple_gate_kernel<float, false><<<nblocks, nthr, 0, stream>>>(reinterpret_cast<float2 *>(amp), n_pairs, skip,
                                                                          1ull << target, ctrl_mask, m);
    }
    return cudaGetLastError();
}

cudaError_t launch_simple_parity(int prec, void *amp, int n_lanes, const double *d0, const double *d1, uint64_t parity,
                                 uint64_t ctrl_mask, cudaStream_t stream) {
    const uint64_t n = 1ull << n_lanes;
    const unsigned nblocks = (unsigned)((n + 255) / 256);
    if (prec == 1)
        simple_parity_kernel<double><<<nblocks, 256, 0, stream>>>(reinterpret_cast<double2 *>(amp), n, parity, ctrl_mask,
                                                                 d0[0], d0[1], d1[0], d1[1]);
    else
        simple_parity_kernel<float><<<nblocks, 256, 0, stream>>>(reinterpret_cast<float2 *>(amp), n, parity, ctrl_mask,
                                                                (float)d0[0], (float)d0[1], (float)d1[0], (float)d1[1]);
    return cudaGetLastError();
}

cudaError_t launch_set_basis_state(int prec, void *amp, uint64_t n_amps, uint64_t one_at,
                                   cudaStream_t stream) {
    const size_t bytes = n_amps * (prec == 1 ? 16 : 8);
    cudaError_t rc = cudaMemsetAsync(amp, 0, bytes, stream);
    if (rc != cudaSuccess) return rc;
    if (prec == 1)
        set_one_kernel<double><<<1, 1, 0, stream>>>(reinterpret_cast<double2 *>(amp), one_at);
    else
        set_one_kernel<float><<<1, 1, 0, stream>>>(reinterpret_cast<float2 *>(amp), one_at);
    return cudaGetLastError();
}

cudaError_t launch_prob0(int prec, const void *amp, int n_lanes, int lane, double *d_partials,
                         double *d_result, cudaStream_t stream) {
    const uint64_t n_half = 1ull << (n_lanes - 1);
    const unsigned nblocks = grid_for((n_half + 3) / 4, 256, 148 * 8);
    if (prec == 1)
        prob0_kernel<double><<<nblocks, 256, 0, stream>>>(reinterpret_cast<const double2 *>(amp), n_half,
                                                         lane, d_partials);
    else
        prob0_kernel<float><<<nblocks, 256, 0, stream>>>(reinterpret_cast<const float2 *>(amp), n_half,
                                                        lane, d_partials);
    cudaError_t rc = cudaGetLastError();
    if (rc != cudaSuccess) return rc;
    final_sum_kernel<<<1, 1024, 0, stream>>>(d_partials, (int)nblocks, d_result);
    return cudaGetLastError();
}

cudaError_t launch_norm(int prec, const void *amp, int n_lanes, double *d_partials, double *d_result,
                        cudaStream_t stream) {
    /* prob0_kernel with a "lane" above every index bit sums |a|^2 over the whole array */
    const uint64_t n = 1ull << n_lanes;
    const unsigned nblocks = grid_for((n + 3) / 4, 256, 148 * 8);
    if (prec == 1)
        prob0_kernel<double><<<nblocks, 256, 0, stream>>>(reinterpret_cast<const double2 *>(amp), n, 62,
                                                         d_partials);
    else
        prob0_kernel<float><<<nblocks, 256, 0, stream>>>(reinterpret_cast<const float2 *>(amp), n, 62,
                                                        d_partials);
    cudaError_t rc = cudaGetLastError();
    if (rc != cudaSuccess) return rc;
    final_sum_kernel<<<1, 1024, 0, stream>>>(d_partials, (int)nblocks, d_result);
    return cudaGetLastError();
}

cudaError_t launch_decohere(int prec, void *amp, int n_lanes, int lane, int value, double norm,
                            cudaStream_t stream) {
    const uint64_t n = 1ull << n_lanes;
    const unsigned nblocks = grid_for(n, 256, kStreamCap);
    if (prec == 1)
        decohere_kernel<double><<<nblocks, 256, 0, stream>>>(reinterpret_cast<double2 *>(amp), n, lane,
                                                            value, norm);
    else
        decohere_kernel<float><<<nblocks, 256, 0, stream>>>(reinterpret_cast<float2 *>(amp), n, lane,
                                                           value, (float)norm);
    return cudaGetLastError();
}

cudaError_t launch_decohere_separate(int prec, void *dst, const void *src, int n_src_lanes, int lane,
                                     int value, double norm, cudaStream_t stream) {
    const uint64_t n_dst = 1ull << (n_src_lanes - 1);
    const unsigned nblocks = grid_for(n_dst, 256, kStreamCap);
    if (prec == 1)
        decohere_separate_kernel<double><<<nblocks, 256, 0, stream>>>(
            reinterpret_cast<double2 *>(dst), reinterpret_cast<const double2 *>(src), n_dst, lane, value,
            norm);
    else
        decohere_separate_kernel<float><<<nblocks, 256, 0, stream>>>(
            reinterpret_cast<float2 *>(dst), reinterpret_cast<const float2 *>(src), n_dst, lane, value,
            (float)norm);
    return cudaGetLastError();
}

cudaError_t launch_apply_reset(int prec, void *amp, int n_lanes, int lane, cudaStream_t stream) {
    const uint64_t n_half = 1ull << (n_lanes - 1);
    const unsigned nblocks = grid_for(n_half, 256, kStreamCap);
    if (prec == 1)
        apply_reset_kernel<double><<<nblocks, 256, 0, stream>>>(reinterpret_cast<double2 *>(amp), n_half,
                                                               lane);
    else
        apply_reset_kernel<float><<<nblocks, 256, 0, stream>>>(reinterpret_cast<float2 *>(amp), n_half,
                                                              lane);
    return cudaGetLastError();
}

cudaError_t launch_join(int prec, void *dst, int n_dst_lanes, int n_product_lanes, uint64_t index_offset,
                        const JoinParams &jp, cudaStream_t stream) {
    const uint64_t n_dst = 1ull << n_dst_lanes, n_product = 1ull << n_product_lanes;
    const unsigned nblocks = grid_for(n_dst, 256, kStreamCap);
    if (prec == 1)
        join_kernel<double><<<nblocks, 256, 0, stream>>>(reinterpret_cast<double2 *>(dst), n_dst,
                                                        n_product, index_offset, jp);
    else
        join_kernel<float><<<nblocks, 256, 0, stream>>>(reinterpret_cast<float2 *>(dst), n_dst, n_product,
                                                       index_offset, jp);
    return cudaGetLastError();
}

cudaError_t launch_get_states(int prec, void *d_out, int mathop, const GatherParams &gp,
                              uint64_t empty_mask, int64_t first, int64_t count, int64_t start,
                              int64_t step, cudaStream_t stream) {
    const unsigned nblocks = grid_for((uint64_t)count, 256, kStreamCap);
    if (prec == 1) {
        if (mathop == 0)
            get_states_kernel<double, 0><<<nblocks, 256, 0, stream>>>(d_out, gp, empty_mask, first, count,
                                                                     start, step);
        else
            get_states_kernel<double, 1><<<nblocks, 256, 0, stream>>>(d_out, gp, empty_mask, first, count,
                                                                     start, step);
    } else {
        if (mathop == 0)
            get_states_kernel<float, 0><<<nblocks, 256, 0, stream>>>(d_out, gp, empty_mask, first, count,
                                                                    start, step);
        else
            get_states_kernel<float, 1><<<nblocks, 256, 0, stream>>>(d_out, gp, empty_mask, first, count,
                                                                    start, step);
    }
    return cudaGetLastError();
}

cudaError_t launch_prob_array(int prec, double *d_out, const GatherParams &gp, int n_hidden, int64_t first,
                              int64_t count, bool compat, cudaStream_t stream) {
    const unsigned nblocks = grid_for((uint64_t)count, 256, kStreamCap);
    if (prec == 1) /* complex128: `real` is double, the two accumulations coincide */
        prob_array_kernel<double, double><<<nblocks, 256, 0, stream>>>(d_out, gp, n_hidden, first, count);
    else if (compat)
        prob_array_kernel<float, float><<<nblocks, 256, 0, stream>>>(d_out, gp, n_hidden, first, count);
    else
        prob_array_kernel<float, double><<<nblocks, 256, 0, stream>>>(d_out, gp, n_hidden, first, count);
    return cudaGetLastError();
}

cudaError_t launch_reduce_groups(double *d_out, const double *d_in, int log2_group, int64_t count,
                                 bool fp32_sums, cudaStream_t stream) {
    const unsigned nblocks = grid_for((uint64_t)count, 256, kStreamCap);
    if (fp32_sums)
        reduce_groups_kernel<float><<<nblocks, 256, 0, stream>>>(d_out, d_in, log2_group, count);
    else
        reduce_groups_kernel<double><<<nblocks, 256, 0, stream>>>(d_out, d_in, log2_group, count);
    return cudaGetLastError();
}

cudaError_t launch_compat_scan(int prec, double *d_prob, int64_t n, int64_t span, int n_spans,
                               double *d_span_total, cudaStream_t stream) {
    if (prec == 1)
        compat_scan_kernel<double><<<(unsigned)n_spans, 32, 0, stream>>>(d_prob, n, span, d_span_total);
    else
        compat_scan_kernel<float><<<(unsigned)n_spans, 32, 0, stream>>>(d_prob, n, span, d_span_total);
    return cudaGetLastError();
}

cudaError_t launch_compat_apply(int prec, double *d_prob, int64_t n, int64_t span, const double *d_span_offset,
                                double norm, cudaStream_t stream) {
    const unsigned nblocks = grid_for((uint64_t)n, 256, kStreamCap);
    if (prec == 1)
        compat_apply_kernel<double><<<nblocks, 256, 0, stream>>>(d_prob, n, span, d_span_offset, norm);
    else
        compat_apply_kernel<float><<<nblocks, 256, 0, stream>>>(d_prob, n, span, d_span_offset, norm);
    return cudaGetLastError();
}

cudaError_t launch_cast_from_double(int prec, void *d_out, const double *d_in, int64_t count,
                                    cudaStream_t stream) {
    const unsigned nblocks = grid_for((uint64_t)count, 256, kStreamCap);
    if (prec == 1)
        cast_from_double_kernel<double><<<nblocks, 256, 0, stream>>>(reinterpret_cast<double *>(d_out), d_in,
                                                                    count);
    else
        cast_from_double_kernel<float><<<nblocks, 256, 0, stream>>>(reinterpret_cast<float *>(d_out), d_in,
                                                                   count);
    return cudaGetLastError();
}

/* amp == nullptr: the source is d_cum itself (the marginal probability vector, scanned in place) */
cudaError_t launch_scan_phase1(int prec, const void *amp, const double *d_prob, int64_t n, double *d_block_sums,
                               cudaStream_t stream) {
    const unsigned nblocks = (unsigned)((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
    if (!amp)
        scan_phase1_kernel<SrcProb><<<nblocks, SCAN_THREADS, 0, stream>>>(SrcProb{d_prob}, n, d_block_sums);
    else if (prec == 1)
        scan_phase1_kernel<SrcAmp<double>><<<nblocks, SCAN_THREADS, 0, stream>>>(
            SrcAmp<double>{reinterpret_cast<const double2 *>(amp)}, n, d_block_sums);
    else
        scan_phase1_kernel<SrcAmp<float>><<<nblocks, SCAN_THREADS, 0, stream>>>(
            SrcAmp<float>{reinterpret_cast<const float2 *>(amp)}, n, d_block_sums);
    return cudaGetLastError();
}

cudaError_t launch_scan_phase2(double *d_block_sums, int64_t n_blocks, double *d_total, cudaStream_t stream) {
    scan_phase2_kernel<<<1, SCAN_THREADS, 0, stream>>>(d_block_sums, n_blocks, d_total);
    return cudaGetLastError();
}

cudaError_t launch_scan_phase3(int prec, const void *amp, double *d_cum, int64_t n, const double *d_block_sums,
                               double global_offset, double norm, cudaStream_t stream) {
    const unsigned nblocks = (unsigned)((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
    if (!amp)
        scan_phase3_kernel<SrcProb><<<nblocks, SCAN_THREADS, 0, stream>>>(SrcProb{d_cum}, d_cum, n, d_block_sums,
                                                                         global_offset, norm);
    else if (prec == 1)
        scan_phase3_kernel<SrcAmp<double>><<<nblocks, SCAN_THREADS, 0, stream>>>(
            SrcAmp<double>{reinterpret_cast<const double2 *>(amp)}, d_cum, n, d_block_sums, global_offset, norm);
    else
        scan_phase3_kernel<SrcAmp<float>><<<nblocks, SCAN_THREADS, 0, stream>>>(
            SrcAmp<float>{reinterpret_cast<const float2 *>(amp)}, d_cum, n, d_block_sums, global_offset, norm);
    return cudaGetLastError();
}

cudaError_t launch_sample_sequential(const double *d_cum, int n_bits, const double *d_rand, int64_t *d_obs,
                                     int n_shots, cudaStream_t stream) {
    if (n_shots <= 0) return cudaSuccess;
    sample_sequential_kernel<<<(unsigned)((n_shots + 255) / 256), 256, 0, stream>>>(d_cum, n_bits, d_rand, d_obs, n_shots);
    return cudaGetLastError();
}

cudaError_t launch_sample(int prec, const double *d_cum, int n_lanes, const double *d_rand, int64_t *d_obs,
                          int n_samples, SortedBits empty_lanes, cudaStream_t stream) {
    if (n_samples <= 0) return cudaSuccess;
    const unsigned nblocks = (unsigned)((n_samples + 255) / 256);
    sample_kernel<<<nblocks, 256, 0, stream>>>(d_cum, 1ll << n_lanes, d_rand, d_obs, n_samples, empty_lanes,
                                               prec == 2 ? 1 : 0);
    return cudaGetLastError();
}

} // namespace qgb
