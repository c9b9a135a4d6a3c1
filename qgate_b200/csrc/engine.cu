/*
 * engine.cu — host side of libqgate_b200.so: objects behind the C ABI of
 * include/qgate_b200.h, the deferred gate queue, memory pool, readout staging.
 *
 * Replaces (all paths under /root/reference/qgate/simulator/src):
 *   CUDAQubitStates.cpp / CUDAQubitProcessor.cpp / CUDAQubitsStatesGetter.cpp   the objects
 *   MultiDeviceMemoryStore.cpp                                                   the allocator
 *   DeviceGetStates.cu / TransferringRunner.cpp                                  readout pipeline
 *   glue.cpp / cudaext.cpp                                                       the boundary
 * Design differences that matter:
 *   - apply_gate* only APPENDS to a per-qstates queue; any observer (probability, readout,
 *     collapse, join, synchronize, pool creation) flushes it through the planner into fused
 *     tile passes (planner.cpp, kernels_tma.cu).  The reference launches one kernel per gate.
 *   - everything runs on one CUDA stream; the host blocks only where a value has to reach
 *     the host (calc_probability, get_states, sample, synchronize).
 *   - no CPU fallback of any kind: without a CUDA device every call fails with QGB_ERR_CUDA.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/qgate_b200.h"
#include "gate_matrix.h"
#include "kernels.h"
#include "planner.h"

namespace qgb {

namespace {

/* ---- errors ------------------------------------------------------------------------------ */

thread_local std::string g_last_error;

struct Error {
    int code;
    std::string msg;
};

[[noreturn]] void fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw Error{code, buf};
}

#define CUDA_CHECK(expr)                                                                           \
    do {                                                                                           \
        cudaError_t rc__ = (expr);                                                                 \
        if (rc__ != cudaSuccess)                                                                   \
            fail(rc__ == cudaErrorMemoryAllocation ? QGB_ERR_OOM : QGB_ERR_CUDA, "%s: %s (%s:%d)", \
                 #expr, cudaGetErrorString(rc__), __FILE__, __LINE__);                             \
    } while (0)

#define QGB_TRY try {
#define QGB_CATCH                                              \
    }                                                          \
    catch (const Error &e) {                                   \
        g_last_error = e.msg;                                  \
        return e.code;                                         \
    }                                                          \
    catch (const std::bad_alloc &) {                           \
        g_last_error = "out of host memory";                   \
        return QGB_ERR_OOM;                                    \
    }                                                          \
    catch (const std::exception &e) {                          \
        g_last_error = e.what();                               \
        return QGB_ERR_RUNTIME;                                \
    }                                                          \
    catch (...) {                                              \
        g_last_error = "unknown C++ exception";                \
        return QGB_ERR_RUNTIME;                                \
    }                                                          \
    return QGB_OK;

/* ---- memory pool --------------------------------------------------------------------------
 * Power-of-two size classes, freed blocks are cached and reused in stream order (one stream,
 * so no cross-stream hazards).  Replaces MultiDeviceMemoryStore.cpp:46-201, which
 * synchronises every device on every alloc / free. */
struct MemPool {
    std::map<size_t, std::vector<void *>> cached;
    std::unordered_map<void *, size_t> live;
    /* blocks other processes may have mapped (CUDA IPC, sharded state vectors): a peer that still maps
     * a block keeps its memory alive, so freeing it here would give nothing back — and dist.py keeps
     * its mappings across state vectors because unmapping a used 16 GiB buffer costs ~0.1 s.  Such
     * blocks are only ever freed by trim_exported(), which dist.py calls after every rank has dropped
     * its mappings (a sharded state vector of another size is about to be created), and at shutdown. */
    std::unordered_set<void *> exported;
    bool pin_exported = false; /* option pin_exported: set by dist.py when it keeps its mappings */
    size_t live_bytes = 0, cached_bytes = 0;
    int64_t budget = -1; /* memory_store_size preference; -1 = whatever the device has */

    static size_t size_class(size_t bytes) {
        size_t c = 512;
        while (c < bytes) c <<= 1;
        return c;
    }

    void trim(bool exported_too = false) {
        for (auto &kv : cached) {
            std::vector<void *> kept;
            for (void *p : kv.second) {
                if (!exported_too && pin_exported && exported.count(p)) {
                    kept.push_back(p);
                    continue;
                }
                cudaFree(p);
                exported.erase(p);
                cached_bytes -= kv.first;
            }
            kv.second.swap(kept);
        }
    }
    void trim_exported() {
        for (auto &kv : cached) {
            std::vector<void *> kept;
            for (void *p : kv.second) {
                if (!exported.count(p)) {
                    kept.push_back(p);
                    continue;
                }
                cudaFree(p);
                exported.erase(p);
                cached_bytes -= kv.first;
            }
            kv.second.swap(kept);
        }
    }

    void *alloc(size_t bytes) {
        const size_t c = size_class(bytes);
        auto it = cached.find(c);
        if (it != cached.end() && !it->second.empty()) {
            void *p = it->second.back();
            it->second.pop_back();
            cached_bytes -= c;
            live[p] = c;
            live_bytes += c;
            return p;
        }
        if (budget >= 0 && (int64_t)(live_bytes + c) > budget)
            fail(QGB_ERR_OOM, "Out of device memory (memory_store_size budget of %lld bytes).",
                 (long long)budget);
        void *p = nullptr;
        cudaError_t rc = cudaMalloc(&p, c);
        if (rc != cudaSuccess) {
            cudaGetLastError();
            trim();
            rc = cudaMalloc(&p, c);
        }
        if (rc != cudaSuccess) {
            cudaGetLastError();
            fail(QGB_ERR_OOM, "Out of device memory (%zu bytes requested, %zu live).", c, live_bytes);
        }
        live[p] = c;
        live_bytes += c;
        return p;
    }

    void release(void *p) {
        if (!p) return;
        auto it = live.find(p);
        if (it == live.end()) return;
        const size_t c = it->second;
        live.erase(it);
        live_bytes -= c;
        /* Keep the block for the next qstates of this size class: cudaMalloc + cudaFree of a 16 GiB
         * state vector cost ~0.3 s each (bench.py e2e split), re-use is free and stream-ordered.
         * Big blocks are capped: at most two per size class (a sharded state vector and the spare
         * buffer of its push exchange) and 136 GiB in total (one 128 GiB shard).  The block released
         * last is the likeliest to be asked for again: older blocks of other size classes make room
         * for it.  An allocation that fails trims the whole cache and retries (alloc). */
        const bool big = c > (size_t(1) << 28);
        const size_t cap = size_t(136) << 30;
        if (pin_exported && exported.count(p)) { /* stays until trim_exported(), see above */
            cached[c].push_back(p);
            cached_bytes += c;
            return;
        }
        if (big && (cached[c].size() >= 2 || c > cap)) {
            cudaFree(p);
            exported.erase(p);
            return;
        }
        if (big && cached_bytes + c > cap) {
            for (auto it2 = cached.rbegin(); it2 != cached.rend() && cached_bytes + c > cap; ++it2) { /* largest first */
                if (it2->first == c) continue;
                std::vector<void *> kept;
                for (void *q : it2->second) {
                    if ((pin_exported && exported.count(q)) || cached_bytes + c <= cap) {
                        kept.push_back(q);
                        continue;
                    }
                    cudaFree(q);
                    exported.erase(q);
                    cached_bytes -= it2->first;
                }
                it2->second.swap(kept);
            }
            if (cached_bytes + c > cap) {
                cudaFree(p);
                exported.erase(p);
                return;
            }
        }
        cached[c].push_back(p);
        cached_bytes += c;
    }

    void clear() {
        trim(true);
        cached.clear();
        cached_bytes = 0;
        for (auto &kv : live) cudaFree(kv.first);
        live.clear();
        exported.clear();
        live_bytes = 0;
    }
};

/* ---- objects ------------------------------------------------------------------------------- */

struct QStates {
    int prec;
    int n_lanes = -1;
    void *d_amp = nullptr;
    void *d_alt = nullptr; /* spare array for out-of-place exchanges (sharded states only) */
    /* reset_qstates does not touch the array: the first fused pass starts from |0...0> without reading
     * it (launch_tma_pass zero_input), any other consumer writes the basis state first (materialize).
     * Saves one write + one read of the state per run: 10% of a 30-qubit QFT. */
    bool zero_pending = false;
    /* lane map: lane l of the caller (the "local lane" of the reference's protocol) is bit perm[l]
     * of the array index.  Identity until a native Swap: exchanging two lanes is exchanging two
     * entries here, no amplitude moves (the reference expands a Swap into three
     * CX gates = three state sweeps, model/expand.py:14-20).  Queued gates, kernels and the tables of
     * a readout are in index-bit coordinates. */
    std::vector<int> perm;
    int phys(int lane) const { return perm[(size_t)lane]; }
    void reset_perm() {
        perm.resize((size_t)std::max(n_lanes, 0));
        for (int l = 0; l < n_lanes; ++l) perm[(size_t)l] = l;
    }
    std::vector<Gate> queue;
    size_t elem() const { return prec == QGB_PREC_FP64 ? 16 : 8; }
    size_t bytes() const { return elem() << n_lanes; }
};

struct QProc {
    int prec;
};

struct Getter {
    int prec;
};

struct Pool {
    int prec;
    int n_lanes;
    double *d_cum = nullptr;
    double *d_sums = nullptr; /* block sums + total, kept between `partial` and `finalize` */
    const void *src_amp = nullptr; /* pool over all lanes of one state vector in index order: the scan */
                                   /* reads |a|^2 straight from the amplitudes (not owned; the state    */
                                   /* must not change between `partial` and `finalize`)                 */
    bool finalized = false;
    SortedBits empty;
};

struct Options {
    int64_t fuse = 1;
    int64_t merge = 1;
    int64_t fans = 1; /* controlled phases that share a lane merge into one phase fan (program.h OP_FAN) */
    int64_t fan_cost = 2;
    int64_t lazy_reset = 1; /* |0...0> is produced by the first fused pass instead of a sweep of its own */
    /* streaming: once a queue holds stream_hi (merged) gates, passes are planned and launched until
     * stream_keep are left — the device works while the caller is still submitting (the reference front
     * end spends ~5 us per gate in Python).  The pass sequence is the one a single flush at the end
     * plans (379 passes either way on the bench circuit: the planner looks at the front of the queue). */
    int64_t queue_stream = 1, stream_hi = 512, stream_keep = 256;
    /* 1: every gate as submitted, one kernel each, in the reference CPU runtime's arithmetic
     * operation by operation (CPUQubitProcessor.cpp:316-324): amplitudes bit-identical to
     * qgate.simulator.cpu's.  A verification mode (one state sweep per gate), off by default. */
    int64_t exact = 0;
    /* complex128: 10-lane tiles (16 KiB, 64-thread CTAs, 6 per SM) overlap arithmetic and memory
     * traffic better than 11-lane ones: 0.72 vs 0.68 of the HBM roofline at the same pass cost
     * (profiles/r1z_pass_floor_and_sweeps.md) */
    int64_t tile_lanes_fp64 = 10, tile_lanes_fp32 = 11; /* 16 KiB tiles in both precisions */
    /* low contiguous lanes forced into every tile; 0 = 5 complex128 / 6 complex64 (512-byte runs);
     * the TMA kernel needs at least the 128-byte row of its tensor map (3 / 4) */
    int64_t low_lanes_fp64 = 0, low_lanes_fp32 = 0;
    int64_t max_gates_per_pass = QGB_MAX_OPS;
    /* cap on the summed op cost of a pass (dense 2x2 = 4, diagonal / swap = 1); 0 = 24 on the TMA
     * kernel — 6 dense ops keep a pass above 70% of the HBM roofline in both precisions; more ops
     * per pass raise throughput by up to 18% and lower that fraction (profiles/r1z sweeps) —
     * unlimited on the cp.async kernel */
    int64_t max_cost = 0;
    int64_t lookahead = 4096;
    int64_t queue_limit = 1 << 16; /* flush when a queue grows past this many gates */
    int64_t ctas_per_sm = 0;       /* smem budget: resident CTAs to aim for (0 = by shape)   */
    /* dense gates as three in-place shears (program.h OP_SHEAR): 12 instead of 16 FMAs per pair */
    int64_t shear_fp64 = 1, shear_fp32 = 1;
    int64_t warp_local = 0;        /* TMA kernel: warp instead of CTA barriers between stages where the */
                                   /* planner can arrange it (correct, measured neutral: off)           */
    int64_t reg_bits_fp64 = 4;     /* amplitudes per thread = 2^this (TMA kernel: 3 or 4)     */
    int64_t tma_buffers = 0;       /* tile buffers per CTA of the TMA kernel: 2, 3, 0 = 2 for  */
                                   /* complex128, 3 for complex64 (measured best, profiles/)  */
    /* W > 0: sampling pools and marginal probabilities are built exactly as a reference CPU runtime
     * with W worker threads builds them — running sums in the state precision, W sequential spans
     * (CPUSamplingPool.cpp:13-47, CPUQubitsStatesGetter.cpp:179-247) — so that sampled indices are
     * bit-identical to its for identical probabilities.  0: the parallel float64 scan. */
    int64_t pool_compat_workers = 0;
    int64_t pool_from_amplitudes = 1; /* 0: always materialise the probability vector first (A/B switch) */
};

struct Engine {
    bool initialized = false;
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    MemPool pool;
    double *d_partials = nullptr; /* 2048 doubles + scalars */
    double *h_scalar = nullptr;   /* pinned */
    void *d_stage = nullptr;      /* readout staging */
    void *h_stage[2] = {nullptr, nullptr};
    cudaEvent_t stage_done[2] = {nullptr, nullptr};
    size_t stage_bytes = size_t(32) << 20;
    void *h_gather = nullptr;     /* pinned staging of readout lane tables */
    size_t gather_bytes = 0;
    Options opt;
    qgb_stats stats;
    std::unordered_set<QStates *> qstates;
    std::unordered_set<QProc *> qprocs;
    std::unordered_set<Getter *> getters;
    std::unordered_set<Pool *> pools;
};

Engine g;

} // namespace
void close_all_ipc_mappings(); /* defined next to the IPC entry points */
namespace {

void require_init() {
    if (!g.initialized) fail(QGB_ERR_RUNTIME, "devices are not initialized (call qgb_devices_initialize).");
}

QStates *QS(qgb_handle h) {
    QStates *p = reinterpret_cast<QStates *>(h);
    if (!g.qstates.count(p)) fail(QGB_ERR_INVALID, "invalid qstates handle.");
    return p;
}
QProc *QP(qgb_handle h) {
    QProc *p = reinterpret_cast<QProc *>(h);
    if (!g.qprocs.count(p)) fail(QGB_ERR_INVALID, "invalid qproc handle.");
    return p;
}
Getter *QG(qgb_handle h) {
    Getter *p = reinterpret_cast<Getter *>(h);
    if (!g.getters.count(p)) fail(QGB_ERR_INVALID, "invalid getter handle.");
    return p;
}
Pool *SP(qgb_handle h) {
    Pool *p = reinterpret_cast<Pool *>(h);
    if (!g.pools.count(p)) fail(QGB_ERR_INVALID, "invalid sampling pool handle.");
    return p;
}

void check_prec(int prec) {
    if (prec != QGB_PREC_FP64 && prec != QGB_PREC_FP32) fail(QGB_ERR_INVALID, "unknown precision.");
}

void check_allocated(const QStates *qs) {
    if (!qs->d_amp || qs->n_lanes < 0) fail(QGB_ERR_RUNTIME, "qstates is not allocated.");
}

void check_lane(const QStates *qs, int lane) {
    if (lane < 0 || lane >= qs->n_lanes) fail(QGB_ERR_INVALID, "lane %d out of range [0, %d).", lane, qs->n_lanes);
}

/* ---- flush: deferred queue -> kernels ------------------------------------------------------ */

int min_tile_lanes(int prec) { return prec == QGB_PREC_FP64 ? 3 + 5 : 4 + 5; }

/* cap on the summed op cost of a pass (sheared 2x2 = 3, direct 2x2 = 4, diagonal / exchange = 1) */
int default_max_cost(bool fp32, bool shear) {
    /* measured on B200 at 30 qubits, sustained (power-capped) runs, profiles/r2c: complex128 8 sheared
     * gates per pass = 2.74e12 updates/s at 0.78 of the HBM roofline (7 gates: 2.72e12 at 0.79, 9:
     * 2.69e12 at 0.74); complex64 7 gates per pass = 4.9e12 at 0.71 (8: 5.05e12 at 0.69) */
    if (!shear) return 24;
    return fp32 ? 21 : 24;
}

template <typename real>
void flush_tiled(QStates *qs, size_t down_to = 0) {
    const bool fp32 = sizeof(real) == 4;
    PlanConfig cfg;
    cfg.fp32 = fp32;
    cfg.K = fp32 ? 4 : (g.opt.reg_bits_fp64 == 3 ? 3 : 4);
    cfg.T = (int)(fp32 ? g.opt.tile_lanes_fp32 : g.opt.tile_lanes_fp64);
    cfg.T = std::max(cfg.K + 5, std::min(cfg.T, cfg.K + 10));
    const int low_opt = (int)(fp32 ? g.opt.low_lanes_fp32 : g.opt.low_lanes_fp64);
    /* runs of 256 bytes.  128-byte runs (the TMA minimum) cost 27% of the DRAM bandwidth (7.83 vs
     * 5.95 ms per complex128 pass at 30 qubits, profiles/r1za); 256-byte runs cost 2.5% against
     * 512-byte ones but leave one more tile lane free for gates: 116 instead of 131 passes on the
     * bench circuit, 2.93e12 instead of 2.74e12 updates/s at 0.74 of the roofline (profiles/r2d) */
    const int low_auto = fp32 ? 5 : 4;
    cfg.L = low_opt > 0 ? low_opt : low_auto;
    const int n_buf = g.opt.tma_buffers == 0 ? (fp32 ? 3 : 2) : (g.opt.tma_buffers >= 3 ? 3 : 2);
    /* ops per pass: their matrices are staged in shared memory, their predicates are one bit each */
    const int max_ops = (int)std::max<int64_t>(1, std::min<int64_t>(g.opt.max_gates_per_pass, 32));
    /* smem limit: shrink the tile until it fits the device's opt-in shared memory */
    while (cfg.T > cfg.K + 5 && tma_pass_smem_bytes(qs->prec, cfg.T, cfg.K, 8, n_buf, max_ops, QGB_MAX_FANS) > (size_t)g.max_smem_optin)
        --cfg.T;
    cfg.T = std::min(cfg.T, qs->n_lanes);
    cfg.L = std::max(fp32 ? 1 : 0, std::min(cfg.L, cfg.T - 1));
    if (cfg.T >= qs->n_lanes) cfg.L = std::min(low_opt > 0 ? low_opt : low_auto, cfg.T);
    /* the 128-byte row of the tensor map lies inside every tile */
    cfg.row_lanes = fp32 ? 4 : 3;
    cfg.max_groups = QGB_MAX_GROUPS;
    cfg.L = std::max(cfg.L, cfg.row_lanes);
    cfg.max_ops = max_ops;
    cfg.warp_local = g.opt.warp_local != 0;
    cfg.shear = (fp32 ? g.opt.shear_fp32 : g.opt.shear_fp64) != 0;
    {
        /* stages: each costs a per-thread table entry in shared memory; keep the CTA small enough
         * for the occupancy its launch bounds ask for */
        const int nthr = 1 << (cfg.T - cfg.K);
        int want_ctas = nthr <= 256 ? (n_buf == 3 ? 2 : 3) : 1;
        if (g.opt.ctas_per_sm > 0) want_ctas = (int)g.opt.ctas_per_sm;
        const size_t budget = (size_t)(g.max_smem_optin + 1024) / want_ctas - 1024;
        int ms = QGB_MAX_STAGES;
        while (ms > 2 && tma_pass_smem_bytes(qs->prec, cfg.T, cfg.K, ms, n_buf, max_ops, QGB_MAX_FANS) > budget) --ms;
        cfg.max_stages = ms;
    }
    cfg.max_cost = g.opt.max_cost > 0 ? (int)std::min<int64_t>(g.opt.max_cost, 1 << 30) : default_max_cost(fp32, cfg.shear);
    cfg.lookahead = (int)g.opt.lookahead;
    cfg.fan_cost = (int)std::max<int64_t>(1, g.opt.fan_cost);
    static PassProgram<real> prog; /* ~11 KB, passed by value to the kernel */
    PlanStats st;
    while (qs->queue.size() > down_to) {
        plan_pass<real>(qs->queue, qs->n_lanes, cfg, prog, st);
        if (st.gates_in_pass <= 0) fail(QGB_ERR_RUNTIME, "planner made no progress.");
        if (prog.n_groups < 1) fail(QGB_ERR_RUNTIME, "planner produced a tile no tensor map describes.");
        CUDA_CHECK(launch_tma_pass<real>(prog, qs->d_amp, n_buf, 3, g.stream, qs->zero_pending ? 1 : 0));
        qs->zero_pending = false;
        g.stats.tma_passes += 1;
        g.stats.kernel_launches += 1;
        g.stats.tile_passes += 1;
        g.stats.h2d_bytes += (int64_t)sizeof(prog) + 256; /* the pass program travels as kernel parameters */
        g.stats.gates_executed += st.gates_in_pass;
        g.stats.pass_bytes += (int64_t)(2 * qs->bytes());
        g.stats.shear_ops += st.shear_ops;
        g.stats.direct_ops += st.direct_ops;
        g.stats.fan_ops += st.fan_ops;
        for (int f = 0; f < prog.n_fans; ++f)
            if (prog.fan[f].n_out > 0) { /* fan_tile_table_kernel ran before the pass */
                g.stats.kernel_launches += 1;
                break;
            }
    }
}

void materialize(QStates *qs) {
    if (!qs->zero_pending) return;
    qs->zero_pending = false;
    CUDA_CHECK(launch_set_basis_state(qs->prec, qs->d_amp, 1ull << qs->n_lanes, 0, g.stream));
    g.stats.kernel_launches += 1;
}

void flush(QStates *qs) {
    if (qs->queue.empty()) {
        if (qs->d_amp) materialize(qs);
        return;
    }
    check_allocated(qs);
    const bool exact = g.opt.exact != 0;
    if (!exact && g.opt.fuse && qs->n_lanes >= min_tile_lanes(qs->prec)) {
        if (qs->prec == QGB_PREC_FP64)
            flush_tiled<double>(qs);
        else
            flush_tiled<float>(qs);
    } else {
        materialize(qs);
        for (const Gate &gt : qs->queue) {
            if (!gt.fan.empty()) {
                /* a phase fan (formed while the tiled path was on): its controlled phases one by one */
                for (const Gate::FanTerm &t : gt.fan) {
                    const double cp[8] = {1., 0., 0., 0., 0., 0., t.re, t.im};
                    CUDA_CHECK(launch_simple_gate(qs->prec, qs->d_amp, qs->n_lanes, cp, t.lane, 1ull << gt.target, 0,
                                                  false, g.stream));
                    g.stats.kernel_launches += 1;
                }
                g.stats.gates_executed += 1;
                continue;
            }
            if (gt.parity) {
                CUDA_CHECK(launch_simple_parity(qs->prec, qs->d_amp, qs->n_lanes, gt.m, gt.m + 6, gt.parity,
                                                gt.ctrl_mask, g.stream));
            } else if (gt.mux >= 0) {
                /* multiplexed gate (host-side merging): m where lane mux is 0, m1 where it is 1 */
                CUDA_CHECK(launch_simple_gate(qs->prec, qs->d_amp, qs->n_lanes, gt.m, gt.target, 0,
                                              1ull << gt.mux, false, g.stream));
                CUDA_CHECK(launch_simple_gate(qs->prec, qs->d_amp, qs->n_lanes, gt.m1, gt.target,
                                              1ull << gt.mux, 0, false, g.stream));
                g.stats.kernel_launches += 1;
            } else {
                CUDA_CHECK(launch_simple_gate(qs->prec, qs->d_amp, qs->n_lanes, gt.m, gt.target, gt.ctrl_mask,
                                              0, exact, g.stream));
            }
            g.stats.kernel_launches += 1;
            g.stats.gates_executed += 1;
        }
        qs->queue.clear();
    }
}

uint64_t control_mask(const QStates *qs, const int *ctrl, int n_ctrl, uint64_t forbidden) {
    uint64_t mask = 0;
    for (int i = 0; i < n_ctrl; ++i) {
        check_lane(qs, ctrl[i]);
        const uint64_t bit = 1ull << qs->phys(ctrl[i]);
        if (bit & forbidden) fail(QGB_ERR_INVALID, "control lane equals target lane.");
        mask |= bit;
    }
    return mask;
}

void submit_queued(QStates *qs, const Gate &gt, int n_ctrl);

void submit_gate(QStates *qs, const double *mat8, const int *ctrl, int n_ctrl, int target) {
    check_allocated(qs);
    check_lane(qs, target);
    Gate gt;
    std::memcpy(gt.m, mat8, sizeof(gt.m));
    gt.target = qs->phys(target);
    gt.ctrl_mask = control_mask(qs, ctrl, n_ctrl, 1ull << gt.target);
    submit_queued(qs, gt, n_ctrl);
}

void submit_queued(QStates *qs, const Gate &gt, int n_ctrl) {
    /* merging pays on the tiled path only; the one-kernel-per-gate path keeps the submitted gates */
    const bool tiled = !g.opt.exact && g.opt.fuse && qs->n_lanes >= min_tile_lanes(qs->prec);
    enqueue_gate(qs->queue, gt, g.opt.merge != 0 && tiled, g.opt.fans != 0);
    g.stats.gates_submitted += 1;
    g.stats.gate_amp_updates += (int64_t)1 << (qs->n_lanes - n_ctrl);
    if ((int64_t)qs->queue.size() >= g.opt.queue_limit) flush(qs);
    if (tiled && g.opt.queue_stream && g.opt.stream_hi > 0 && (int64_t)qs->queue.size() >= g.opt.stream_hi) {
        const size_t keep = (size_t)std::max<int64_t>(0, std::min(g.opt.stream_keep, g.opt.stream_hi - 1));
        if (qs->prec == QGB_PREC_FP64)
            flush_tiled<double>(qs, keep);
        else
            flush_tiled<float>(qs, keep);
    }
}

void stream_sync() { CUDA_CHECK(cudaStreamSynchronize(g.stream)); }

void free_qstates_buffer(QStates *qs) {
    if (qs->d_amp) {
        g.pool.release(qs->d_amp);
        qs->d_amp = nullptr;
    }
    if (qs->d_alt) {
        g.pool.release(qs->d_alt);
        qs->d_alt = nullptr;
    }
    qs->zero_pending = false;
    qs->queue.clear();
}

/* ---- readout helpers ------------------------------------------------------------------------ */

/* lane tables of a readout -> device memory (pinned staging, stream-ordered); the caller hands the
 * device block back with release_gather once its kernels are queued */
void build_gather(GatherParams &gp, int prec, const int *lane_tables, const int *n_per,
                  const qgb_handle *list, int n_qstates, int n_ext_lanes) {
    if (n_qstates < 0) fail(QGB_ERR_INVALID, "negative number of qstates.");
    gp.n_qstates = n_qstates;
    gp.qs = nullptr;
    if (n_qstates == 0) return;
    const size_t bytes = sizeof(LaneTable) * (size_t)n_qstates;
    if (bytes > g.gather_bytes) {
        stream_sync(); /* the old staging block may still be the source of a copy */
        if (g.h_gather) CUDA_CHECK(cudaFreeHost(g.h_gather));
        g.h_gather = nullptr;
        g.gather_bytes = std::max<size_t>(2 * bytes, 64 * sizeof(LaneTable));
        CUDA_CHECK(cudaMallocHost(&g.h_gather, g.gather_bytes));
    }
    LaneTable *host = static_cast<LaneTable *>(g.h_gather);
    const int *p = lane_tables;
    for (int q = 0; q < n_qstates; ++q) {
        QStates *qs = QS(list[q]);
        check_allocated(qs);
        if (qs->prec != prec) fail(QGB_ERR_RUNTIME, "Wrong type");
        if (n_per[q] != qs->n_lanes)
            fail(QGB_ERR_INVALID, "lane table of qstates %d has %d entries, expected %d.", q, n_per[q],
                 qs->n_lanes);
        flush(qs);
        host[q].amp = qs->d_amp;
        host[q].n_lanes = qs->n_lanes;
        for (int l = 0; l < qs->n_lanes; ++l) {
            if (p[l] < 0 || p[l] >= n_ext_lanes)
                fail(QGB_ERR_INVALID, "external lane %d out of range [0, %d).", p[l], n_ext_lanes);
            host[q].ext[qs->phys(l)] = (int8_t)p[l]; /* the table is indexed by index bit */
        }
        p += n_per[q];
    }
    void *dev = g.pool.alloc(bytes);
    gp.qs = static_cast<const LaneTable *>(dev);
    CUDA_CHECK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, g.stream));
    g.stats.h2d_bytes += (int64_t)bytes;
}

void release_gather(GatherParams &gp) {
    if (gp.qs) g.pool.release(const_cast<LaneTable *>(gp.qs));
    gp.qs = nullptr;
}

/* marginal probability vector on the device, 2^n_lanes doubles (caller releases to the pool).
 * Default: hidden lanes summed in double, 8 lanes per level.  compat (option pool_compat_workers):
 * the reference CPU runtime's own levels and precision — 4 lanes per level, running sums in the
 * state precision (CPUQubitsStatesGetter.cpp:179-247). */
double *device_prob_array(int prec, const int *lane_tables, const int *n_per, const qgb_handle *list,
                          int n_qstates, int n_lanes, int n_hidden) {
    if (n_lanes < 0 || n_hidden < 0 || n_lanes > QGB_MAX_LANES || n_lanes + n_hidden > 62)
        fail(QGB_ERR_INVALID, "bad lane counts (%d, %d).", n_lanes, n_hidden);
    const bool compat = g.opt.pool_compat_workers > 0;
    const int per_level = compat ? 4 : 8;
    GatherParams gp;
    build_gather(gp, prec, lane_tables, n_per, list, n_qstates, n_lanes + n_hidden);
    double *cur = nullptr;
    try {
        int h0 = std::min(n_hidden, per_level);
        int rest = n_hidden - h0;
        int64_t count = (int64_t)1 << (n_lanes + rest);
        cur = static_cast<double *>(g.pool.alloc(sizeof(double) * (size_t)count));
        CUDA_CHECK(launch_prob_array(prec, cur, gp, h0, 0, count, compat, g.stream));
        g.stats.kernel_launches += 1;
        while (rest > 0) {
            const int h = std::min(rest, per_level);
            rest -= h;
            count = (int64_t)1 << (n_lanes + rest);
            double *next = static_cast<double *>(g.pool.alloc(sizeof(double) * (size_t)count));
            cudaError_t rc = launch_reduce_groups(next, cur, h, count, compat && prec == QGB_PREC_FP32, g.stream);
            g.pool.release(cur);
            cur = next;
            CUDA_CHECK(rc);
            g.stats.kernel_launches += 1;
        }
    } catch (...) {
        release_gather(gp);
        if (cur) g.pool.release(cur);
        throw;
    }
    release_gather(gp);
    return cur;
}

/* device -> caller's (pageable) host buffer.  Small results go directly; large ones (a 2^30-entry
 * probability array is 8 GiB) travel through the two pinned staging buffers, the host memcpy of chunk
 * i overlapping the DMA of chunk i + 1 (a pageable cudaMemcpy stages through a small driver buffer
 * and runs at a fraction of the link rate). */
void copy_to_host(void *dst, const void *d_src, size_t bytes) {
    g.stats.d2h_bytes += (int64_t)bytes;
    if (bytes <= (size_t(1) << 20) || !g.h_stage[0] || !g.h_stage[1]) {
        CUDA_CHECK(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, g.stream));
        stream_sync();
        return;
    }
    const size_t chunk = g.stage_bytes;
    const size_t n_chunks = (bytes + chunk - 1) / chunk;
    auto issue = [&](size_t i) {
        const size_t off = i * chunk, n = std::min(chunk, bytes - off);
        CUDA_CHECK(cudaMemcpyAsync(g.h_stage[i & 1], static_cast<const char *>(d_src) + off, n, cudaMemcpyDeviceToHost,
                                   g.stream));
        CUDA_CHECK(cudaEventRecord(g.stage_done[i & 1], g.stream));
    };
    issue(0);
    for (size_t i = 0; i < n_chunks; ++i) {
        if (i + 1 < n_chunks) issue(i + 1); /* (buffer (i + 1) & 1 was drained by the memcpy of chunk i - 1) */
        CUDA_CHECK(cudaEventSynchronize(g.stage_done[i & 1]));
        const size_t off = i * chunk, n = std::min(chunk, bytes - off);
        std::memcpy(static_cast<char *>(dst) + off, g.h_stage[i & 1], n);
    }
}

} // namespace

} // namespace qgb

using namespace qgb;

extern "C" {

const char *qgb_last_error(void) { return g_last_error.c_str(); }
const char *qgb_backend_name(void) { return "cuda-sm_100a"; }
int qgb_abi_version(void) { return 1; }

int qgb_device_count(int *count) {
    QGB_TRY
    int n = 0;
    cudaError_t rc = cudaGetDeviceCount(&n);
    if (rc != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *count = n;
    QGB_CATCH
}

int qgb_devices_initialize(const int *device_ids, int n_device_ids, int max_po2idx_per_chunk,
                           int64_t memory_store_size) {
    QGB_TRY
    (void)max_po2idx_per_chunk; /* one allocation per state vector: chunking is not needed */
    if (g.initialized) fail(QGB_ERR_RUNTIME, "already initialized.");
    int n = 0;
    CUDA_CHECK(cudaGetDeviceCount(&n));
    if (n == 0) fail(QGB_ERR_CUDA, "no CUDA device.");
    /* device_ids (cudaruntime.set_preference, qgate/simulator/cudaruntime.py:33-41).  This engine
     * drives ONE GPU per process (N GPUs = N ranks under torchrun, qgate_b200/dist.py), so:
     *   []               the current device;
     *   [d]              device d;
     *   [d, d, ..., d]   the reference's "logical devices on one GPU" (tests/test_multidevice.py:
     *                    13-21, examples/mgpu.py:10-16): k logical devices with a budget of
     *                    memory_store_size each are one device with k times that budget — one
     *                    allocation per state vector, chunking is never needed;
     *   distinct ids     refused: silently using the first one would give the caller one GPU's
     *                    memory where it asked for several. */
    int dev = 0;
    if (n_device_ids > 0) {
        dev = device_ids[0];
        for (int i = 1; i < n_device_ids; ++i)
            if (device_ids[i] != dev)
                fail(QGB_ERR_INVALID,
                     "device_ids names %d different GPUs: this runtime drives one GPU per process; launch one "
                     "process per GPU (torchrun --nproc-per-node N) and the state vector is sharded over the "
                     "ranks (qgate_b200.dist).", n_device_ids);
        if (memory_store_size >= 0) memory_store_size *= n_device_ids;
    } else {
        CUDA_CHECK(cudaGetDevice(&dev));
    }
    if (dev < 0 || dev >= n) fail(QGB_ERR_INVALID, "device id %d out of range [0, %d).", dev, n);
    CUDA_CHECK(cudaSetDevice(dev));
    g.device = dev;
    CUDA_CHECK(cudaDeviceGetAttribute(&g.sm_count, cudaDevAttrMultiProcessorCount, dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&g.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    CUDA_CHECK(cudaStreamCreateWithFlags(&g.own_stream, cudaStreamNonBlocking));
    g.stream = g.own_stream;
    CUDA_CHECK(tma_pass_configure(g.max_smem_optin, g.sm_count));
    g.pool.budget = memory_store_size;
    CUDA_CHECK(cudaMalloc(&g.d_partials, sizeof(double) * 4096));
    CUDA_CHECK(cudaMallocHost(&g.h_scalar, sizeof(double) * 16));
    CUDA_CHECK(cudaMalloc(&g.d_stage, g.stage_bytes));
    for (int i = 0; i < 2; ++i) {
        CUDA_CHECK(cudaMallocHost(&g.h_stage[i], g.stage_bytes));
        CUDA_CHECK(cudaEventCreateWithFlags(&g.stage_done[i], cudaEventDisableTiming));
    }
    g.initialized = true;
    QGB_CATCH
}

int qgb_devices_clear(void) {
    QGB_TRY
    if (!g.initialized) return QGB_OK;
    cudaStreamSynchronize(g.stream);
    for (QStates *qs : g.qstates) {
        qs->d_amp = nullptr;
        qs->d_alt = nullptr;
        qs->queue.clear();
        qs->n_lanes = -1;
    }
    for (Pool *p : g.pools) p->d_cum = p->d_sums = nullptr;
    close_all_ipc_mappings();
    g.pool.clear();
    tma_pass_release();
    if (g.h_gather) cudaFreeHost(g.h_gather);
    g.h_gather = nullptr;
    g.gather_bytes = 0;
    cudaFree(g.d_partials);
    cudaFreeHost(g.h_scalar);
    cudaFree(g.d_stage);
    for (int i = 0; i < 2; ++i) {
        cudaFreeHost(g.h_stage[i]);
        cudaEventDestroy(g.stage_done[i]);
        g.h_stage[i] = nullptr;
        g.stage_done[i] = nullptr;
    }
    g.d_partials = nullptr;
    g.h_scalar = nullptr;
    g.d_stage = nullptr;
    cudaStreamDestroy(g.own_stream);
    g.own_stream = g.stream = nullptr;
    g.initialized = false;
    QGB_CATCH
}

int qgb_set_stream(uint64_t cuda_stream) {
    QGB_TRY
    require_init();
    stream_sync();
    g.stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : g.own_stream;
    QGB_CATCH
}

int qgb_gate_matrix(int gate_id, const double *args, int n_args, int adjoint, double *mat8) {
    QGB_TRY
    const int want = gate_matrix_n_args(gate_id);
    if (want < 0) fail(QGB_ERR_RUNTIME, "Unknown gate type.");
    if (n_args != want) fail(QGB_ERR_INVALID, "wrong number of gate arguments.");
    gate_matrix(gate_id, args, adjoint, mat8);
    QGB_CATCH
}

/* ---- QubitStates ------------------------------------------------------------------------------ */

int qgb_qstates_new(int prec, qgb_handle *out) {
    QGB_TRY
    check_prec(prec);
    QStates *qs = new QStates();
    qs->prec = prec;
    g.qstates.insert(qs);
    *out = reinterpret_cast<qgb_handle>(qs);
    QGB_CATCH
}

int qgb_qstates_delete(qgb_handle h) {
    QGB_TRY
    QStates *qs = QS(h);
    free_qstates_buffer(qs);
    g.qstates.erase(qs);
    delete qs;
    QGB_CATCH
}

int qgb_qstates_deallocate(qgb_handle h) {
    QGB_TRY
    QStates *qs = QS(h);
    free_qstates_buffer(qs);
    qs->n_lanes = -1;
    QGB_CATCH
}

int qgb_qstates_get_n_lanes(qgb_handle h, int *n_lanes) {
    QGB_TRY
    *n_lanes = QS(h)->n_lanes;
    QGB_CATCH
}

/* ---- QubitProcessor ------------------------------------------------------------------------------ */

int qgb_qproc_new(int prec, qgb_handle *out) {
    QGB_TRY
    check_prec(prec);
    QProc *qp = new QProc();
    qp->prec = prec;
    g.qprocs.insert(qp);
    *out = reinterpret_cast<qgb_handle>(qp);
    QGB_CATCH
}

int qgb_qproc_delete(qgb_handle h) {
    QGB_TRY
    QProc *qp = QP(h);
    g.qprocs.erase(qp);
    delete qp;
    QGB_CATCH
}

int qgb_qproc_synchronize(qgb_handle h) {
    QGB_TRY
    QP(h);
    require_init();
    for (QStates *qs : g.qstates) flush(qs);
    stream_sync();
    QGB_CATCH
}

int qgb_qproc_reset(qgb_handle h) {
    QGB_TRY
    QP(h);
    QGB_CATCH
}

int qgb_qproc_initialize_qstates(qgb_handle qp, qgb_handle h, int n_lanes) {
    QGB_TRY
    QP(qp);
    require_init();
    QStates *qs = QS(h);
    if (n_lanes < 0 || n_lanes > QGB_MAX_LANES) fail(QGB_ERR_INVALID, "n_lanes %d out of range.", n_lanes);
    free_qstates_buffer(qs);
    qs->n_lanes = n_lanes;
    qs->reset_perm();
    qs->d_amp = g.pool.alloc(qs->bytes());
    QGB_CATCH
}

int qgb_qproc_reset_qstates(qgb_handle qp, qgb_handle h) {
    QGB_TRY
    QP(qp);
    require_init();
    QStates *qs = QS(h);
    check_allocated(qs);
    qs->queue.clear();
    qs->reset_perm();
    if (g.opt.lazy_reset) {
        qs->zero_pending = true; /* written by the first pass, or by materialize() */
    } else {
        qs->zero_pending = false;
        CUDA_CHECK(launch_set_basis_state(qs->prec, qs->d_amp, 1ull << qs->n_lanes, 0, g.stream));
        g.stats.kernel_launches += 1;
    }
    QGB_CATCH
}

int qgb_qproc_calc_probability(qgb_handle qp, qgb_handle h, int lane, double *prob) {
    QGB_TRY
    QP(qp);
    require_init();
    QStates *qs = QS(h);
    check_allocated(qs);
    check_lane(qs, lane);
    flush(qs);
    CUDA_CHECK(launch_prob0(qs->prec, qs->d_amp, qs->n_lanes, qs->phys(lane), g.d_partials, g.d_partials + 2048,
                            g.stream));
    g.stats.kernel_launches += 2;
    CUDA_CHECK(cudaMemcpyAsync(g.h_scalar, g.d_partials + 2048, sizeof(double), cudaMemcpyDeviceToHost,
                               g.stream));
    stream_sync();
    g.stats.d2h_bytes += sizeof(double);
    *prob = g.h_scalar[0];
    QGB_CATCH
}

static void join_impl(qgb_handle qp, qgb_handle hdst, const qgb_handle *src_list, int n_src, int n_new_lanes,
                      int n_total_lanes, int64_t index_offset) {
    QP(qp);
    require_init();
    QStates *dst = QS(hdst);
    check_allocated(dst);
    if (n_src < 1 || n_src > QGB_MAX_QSTATES) fail(QGB_ERR_INVALID, "bad number of source qstates (%d).", n_src);
    JoinParams jp;
    jp.n_src = n_src;
    int shift = 0;
    for (int k = n_src - 1; k >= 0; --k) {
        QStates *src = QS(src_list[k]);
        check_allocated(src);
        if (src == dst) fail(QGB_ERR_INVALID, "join destination is also a source.");
        if (src->prec != dst->prec) fail(QGB_ERR_RUNTIME, "Wrong type");
        flush(src);
        jp.src[k] = src->d_amp;
        jp.n_lanes[k] = src->n_lanes;
        jp.shift[k] = shift;
        shift += src->n_lanes;
    }
    if (n_total_lanes < 0) n_total_lanes = dst->n_lanes;
    if (n_new_lanes < 0 || shift + n_new_lanes != n_total_lanes || dst->n_lanes > n_total_lanes)
        fail(QGB_ERR_INVALID, "join: %d source lanes + %d new lanes != %d destination lanes.", shift,
             n_new_lanes, n_total_lanes);
    if (index_offset < 0 || (index_offset & (((int64_t)1 << dst->n_lanes) - 1)) != 0 ||
        index_offset >= ((int64_t)1 << n_total_lanes))
        fail(QGB_ERR_INVALID, "join: bad shard offset.");
    dst->queue.clear();
    dst->zero_pending = false; /* every amplitude of the product is written */
    CUDA_CHECK(launch_join(dst->prec, dst->d_amp, dst->n_lanes, shift, (uint64_t)index_offset, jp, g.stream));
    g.stats.kernel_launches += 1;
    /* the product keeps every source's array as it is: source k's lane l (its index bit perm[l])
     * becomes the product's lane and index bit shift_k + ...; new lanes follow in order */
    dst->reset_perm();
    if (dst->n_lanes == n_total_lanes)
        for (int k = 0; k < n_src; ++k) {
            const QStates *src = QS(src_list[k]);
            for (int l = 0; l < src->n_lanes; ++l) dst->perm[(size_t)(jp.shift[k] + l)] = jp.shift[k] + src->phys(l);
        }
}

int qgb_qproc_join(qgb_handle qp, qgb_handle hdst, const qgb_handle *src_list, int n_src, int n_new_lanes) {
    QGB_TRY
    join_impl(qp, hdst, src_list, n_src, n_new_lanes, -1, 0);
    QGB_CATCH
}

int qgb_qproc_join_shard(qgb_handle qp, qgb_handle hdst, const qgb_handle *src_list, int n_src, int n_new_lanes,
                         int n_total_lanes, int64_t index_offset) {
    QGB_TRY
    join_impl(qp, hdst, src_list, n_src, n_new_lanes, n_total_lanes, index_offset);
    QGB_CATCH
}

int qgb_qproc_calc_norm(qgb_handle qp, qgb_handle h, double *norm) {
    QGB_TRY
    QP(qp);
    require_init();
    QStates *qs = QS(h);
    check_allocated(qs);
    flush(qs);
    CUDA_CHECK(launch_norm(qs->prec, qs->d_amp, qs->n_lanes, g.d_partials, g.d_partials + 2048, g.stream));
    g.stats.kernel_launches += 2;
    CUDA_CHECK(cudaMemcpyAsync(g.h_scalar, g.d_partials + 2048, sizeof(double), cudaMemcpyDeviceToHost,
                               g.stream));
    stream_sync();
    g.stats.d2h_bytes += sizeof(double);
    *norm = g.h_scalar[0];
    QGB_CATCH
}

int qgb_qstates_data_ptr(qgb_handle h, uint64_t *ptr, int64_t *bytes) {
    QGB_TRY
    require_init();
    QStates *qs = QS(h);
    check_allocated(qs);
    flush(qs);
    *ptr = reinterpret_cast<uint64_t>(qs->d_amp);
    if (bytes) *bytes = (int64_t)qs->bytes();
    QGB_CATCH
}

int qgb_qstates_alt_buffer(qgb_handle h, uint64_t *ptr) {
    QGB_TRY
    require_init();
    QStates *qs = QS(h);
    check_allocated(qs);
    if (!qs->d_alt) qs->d_alt = g.pool.alloc(qs->bytes());
    *ptr = reinterpret_cast<uint64_t>(qs->d_alt);
    QGB_CATCH
}

int qgb_qstates_flip(qgb_handle h) {
    QGB_TRY
    require_init();
    QStates *qs = QS(h);
    check_allocated(qs);
    if (!qs->d_alt) fail(QGB_ERR_RUNTIME, "qstates has no spare buffer.");
    flush(qs);
    std::swap(qs->d_amp, qs->d_alt);
    QGB_CATCH
}

/* CUDA IPC.  The handle names the whole driver allocation a pointer lives in (small cudaMalloc
 * blocks may share one), so the export also reports the pointer's offset inside it, and a
 * handle is mapped once per process however many qstates refer to it (reference counted). */
namespace {
typedef int (*GetAddressRangeFn)(unsigned long long *, size_t *, unsigned long long);
struct IpcMapping {
    void *base;
    int refs;
};
std::map<std::string, IpcMapping> g_ipc_open;
} // namespace

} /* extern "C" */
namespace qgb {
void close_all_ipc_mappings() {
    for (auto &kv : g_ipc_open) cudaIpcCloseMemHandle(kv.second.base);
    g_ipc_open.clear();
    cudaGetLastError();
}
} // namespace qgb
extern "C" {

int qgb_qstates_ipc_export(qgb_handle h, void *handle64, int64_t *offset) {
    QGB_TRY
    require_init();
    QStates *qs = QS(h);
    check_allocated(qs);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    void *base = qs->d_amp;
    static GetAddressRangeFn get_range = nullptr;
    if (!get_range) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            get_range = reinterpret_cast<GetAddressRangeFn>(fn);
        cudaGetLastError();
    }
    if (get_range) {
        unsigned long long b = 0;
        size_t size = 0;
        if (get_range(&b, &size, reinterpret_cast<unsigned long long>(qs->d_amp)) == 0 && b) base = reinterpret_cast<void *>(b);
    }
    cudaIpcMemHandle_t hd;
    CUDA_CHECK(cudaIpcGetMemHandle(&hd, base));
    g.pool.exported.insert(qs->d_amp); /* (pool blocks are whole allocations: d_amp is what release() sees) */
    std::memcpy(handle64, &hd, sizeof(hd));
    if (offset) *offset = (int64_t)(reinterpret_cast<char *>(qs->d_amp) - reinterpret_cast<char *>(base));
    QGB_CATCH
}

int qgb_qstates_ipc_export_alt(qgb_handle h, void *handle64, int64_t *offset) {
    QGB_TRY
    require_init();
    QStates *qs = QS(h);
    check_allocated(qs);
    if (!qs->d_alt) qs->d_alt = g.pool.alloc(qs->bytes());
    std::swap(qs->d_amp, qs->d_alt); /* export the spare buffer through the same code path */
    const int rc = qgb_qstates_ipc_export(h, handle64, offset);
    std::swap(qs->d_amp, qs->d_alt);
    if (rc != QGB_OK) return rc;
    QGB_CATCH
}

int qgb_pool_trim_exported(void) {
    QGB_TRY
    require_init();
    g.pool.trim_exported();
    QGB_CATCH
}

int qgb_ipc_open(const void *handle64, uint64_t *base_ptr) {
    QGB_TRY
    require_init();
    const std::string key(reinterpret_cast<const char *>(handle64), 64);
    auto it = g_ipc_open.find(key);
    if (it != g_ipc_open.end()) {
        it->second.refs += 1;
        *base_ptr = reinterpret_cast<uint64_t>(it->second.base);
    } else {
        cudaIpcMemHandle_t hd;
        std::memcpy(&hd, handle64, sizeof(hd));
        void *p = nullptr;
        CUDA_CHECK(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
        g_ipc_open[key] = IpcMapping{p, 1};
        *base_ptr = reinterpret_cast<uint64_t>(p);
    }
    QGB_CATCH
}

int qgb_ipc_close(uint64_t base_ptr) {
    QGB_TRY
    require_init();
    for (auto it = g_ipc_open.begin(); it != g_ipc_open.end(); ++it) {
        if (reinterpret_cast<uint64_t>(it->second.base) != base_ptr) continue;
        if (--it->second.refs == 0) {
            CUDA_CHECK(cudaStreamSynchronize(g.stream)); /* no kernel may still use the mapping */
            void *p = it->second.base;
            g_ipc_open.erase(it);
            CUDA_CHECK(cudaIpcCloseMemHandle(p));
        }
        break;
    }
    QGB_CATCH
}

int qgb_qstates_exchange_p2p(qgb_handle h, const uint64_t *peer_ptrs, int k, const int *victim_lanes,
                             int my_sel) {
    QGB_TRY
    require_init();
    QStates *qs = QS(h);
    check_allocated(qs);
    if (k < 1 || k > QGB_MAX_EXCHANGE_LANES) fail(QGB_ERR_INVALID, "exchange of %d lanes is not supported.", k);
    if (my_sel < 0 || my_sel >= (1 << k)) fail(QGB_ERR_INVALID, "bad exchange selector.");
    const int unit_shift = qs->prec == QGB_PREC_FP64 ? 0 : 1; /* 16-byte units */
    ExchangeParams ep;
    ep.k = k;
    ep.my_sel = my_sel;
    ep.n_unit_bits = qs->n_lanes - unit_shift;
    uint64_t vmask = 0;
    for (int i = 0; i < k; ++i) {
        check_lane(qs, victim_lanes[i]);
        if (victim_lanes[i] < unit_shift || (i > 0 && victim_lanes[i] <= victim_lanes[i - 1]))
            fail(QGB_ERR_INVALID, "victim lanes must ascend and lie above the 16-byte unit.");
        ep.victim[i] = victim_lanes[i] - unit_shift;
        vmask |= 1ull << ep.victim[i];
    }
    if (ep.n_unit_bits < k + 1) fail(QGB_ERR_INVALID, "state too small for a %d-lane exchange.", k);
    ep.split = -1; /* highest unit bit that is not a victim: halves the work of every rank pair */
    for (int b = ep.n_unit_bits - 1; b >= 0; --b)
        if (!(vmask & (1ull << b))) {
            ep.split = b;
            break;
        }
    for (int j = 0; j < (1 << k); ++j) ep.peer[j] = reinterpret_cast<void *>(peer_ptrs[j]);
    flush(qs);
    CUDA_CHECK(launch_exchange_p2p(qs->d_amp, ep, g.sm_count, g.stream));
    g.stats.kernel_launches += 1;
    QGB_CATCH
}

int qgb_qstates_exchange_push(qgb_handle h, const uint64_t *peer_alt_ptrs, int k, const int *victim_lanes,
                              int my_sel) {
    QGB_TRY
    require_init();
    QStates *qs = QS(h);
    check_allocated(qs);
    if (!qs->d_alt) fail(QGB_ERR_RUNTIME, "qstates has no spare buffer.");
    if (k < 1 || k > QGB_MAX_EXCHANGE_LANES) fail(QGB_ERR_INVALID, "exchange of %d lanes is not supported.", k);
    if (my_sel < 0 || my_sel >= (1 << k)) fail(QGB_ERR_INVALID, "bad exchange selector.");
    const int unit_shift = qs->prec == QGB_PREC_FP64 ? 0 : 1; /* 16-byte units */
    ExchangeParams ep;
    ep.k = k;
    ep.my_sel = my_sel;
    ep.n_unit_bits = qs->n_lanes - unit_shift;
    ep.split = -1;
    for (int i = 0; i < k; ++i) {
        check_lane(qs, victim_lanes[i]);
        if (victim_lanes[i] < unit_shift || (i > 0 && victim_lanes[i] <= victim_lanes[i - 1]))
            fail(QGB_ERR_INVALID, "victim lanes must ascend and lie above the 16-byte unit.");
        ep.victim[i] = victim_lanes[i] - unit_shift;
    }
    for (int j = 0; j < (1 << k); ++j) ep.peer[j] = reinterpret_cast<void *>(peer_alt_ptrs[j]);
    if (ep.peer[my_sel] != qs->d_alt) fail(QGB_ERR_INVALID, "own spare buffer expected at the rank's selector.");
    flush(qs);
    CUDA_CHECK(launch_exchange_push(qs->d_amp, ep, g.sm_count, g.stream));
    g.stats.kernel_launches += 1;
    std::swap(qs->d_amp, qs->d_alt); /* every rank of the exchange does the same */
    QGB_CATCH
}

int qgb_qproc_decohere(qgb_handle qp, int value, double prob, qgb_handle h, int lane) {
    QGB_TRY
    QP(qp);
    require_init();
    QStates *qs = QS(h);
    check_allocated(qs);
    check_lane(qs, lane);
    flush(qs);
    /* CPUQubitProcessor.cpp:227,235: the factor is computed in double, then cast */
    const double norm = (value == 0) ? 1. / std::sqrt(prob) : 1. / std::sqrt(1. - prob);
    CUDA_CHECK(launch_decohere(qs->prec, qs->d_amp, qs->n_lanes, qs->phys(lane), value ? 1 : 0, norm, g.stream));
    g.stats.kernel_launches += 1;
    QGB_CATCH
}

int qgb_qproc_decohere_and_separate(qgb_handle qp, int value, double prob, qgb_handle h0, qgb_handle h1,
                                    qgb_handle h, int lane) {
    QGB_TRY
    QP(qp);
    require_init();
    QStates *qs = QS(h), *qs0 = QS(h0), *qs1 = QS(h1);
    check_allocated(qs);
    check_allocated(qs0);
    check_allocated(qs1);
    check_lane(qs, lane);
    if (qs0->n_lanes != qs->n_lanes - 1 || qs1->n_lanes != 1)
        fail(QGB_ERR_INVALID, "decohere_and_separate: lane counts %d -> (%d, %d).", qs->n_lanes,
             qs0->n_lanes, qs1->n_lanes);
    if (qs0->prec != qs->prec || qs1->prec != qs->prec) fail(QGB_ERR_RUNTIME, "Wrong type");
    flush(qs);
    qs0->queue.clear();
    qs1->queue.clear();
    qs0->zero_pending = qs1->zero_pending = false; /* both are written in full below */
    const double norm = (value == 0) ? 1. / std::sqrt(prob) : 1. / std::sqrt(1. - prob);
    const int p = qs->phys(lane);
    CUDA_CHECK(launch_decohere_separate(qs->prec, qs0->d_amp, qs->d_amp, qs->n_lanes, p, value ? 1 : 0,
                                        norm, g.stream));
    CUDA_CHECK(launch_set_basis_state(qs1->prec, qs1->d_amp, 2, value ? 1 : 0, g.stream));
    g.stats.kernel_launches += 2;
    /* the remainder: `lane` leaves, the caller's lanes above it move down by one; index bit p
     * leaves, the index bits above it move down by one */
    qs0->perm.clear();
    for (int l = 0; l < qs->n_lanes; ++l) {
        if (l == lane) continue;
        const int q = qs->phys(l);
        qs0->perm.push_back(q > p ? q - 1 : q);
    }
    qs1->reset_perm();
    QGB_CATCH
}

int qgb_qproc_apply_reset(qgb_handle qp, qgb_handle h, int lane) {
    QGB_TRY
    QP(qp);
    require_init();
    QStates *qs = QS(h);
    check_allocated(qs);
    check_lane(qs, lane);
    flush(qs);
    CUDA_CHECK(launch_apply_reset(qs->prec, qs->d_amp, qs->n_lanes, qs->phys(lane), g.stream));
    g.stats.kernel_launches += 1;
    QGB_CATCH
}

int qgb_qproc_apply_gate(qgb_handle qp, const double *mat8, qgb_handle h, int lane) {
    QGB_TRY
    QP(qp);
    submit_gate(QS(h), mat8, nullptr, 0, lane);
    QGB_CATCH
}

int qgb_qproc_apply_controlled_gate(qgb_handle qp, const double *mat8, qgb_handle h, const int *ctrl,
                                    int n_ctrl, int target) {
    QGB_TRY
    QP(qp);
    submit_gate(QS(h), mat8, ctrl, n_ctrl, target);
    QGB_CATCH
}

int qgb_qproc_apply_gate_typed(qgb_handle qp, int gate_id, const double *args, int n_args, int adjoint,
                               qgb_handle h, const int *ctrl, int n_ctrl, int target) {
    QGB_TRY
    QP(qp);
    const int want = gate_matrix_n_args(gate_id);
    if (want < 0) fail(QGB_ERR_RUNTIME, "Unknown gate type.");
    if (n_args != want) fail(QGB_ERR_INVALID, "wrong number of gate arguments.");
    double m[8];
    gate_matrix(gate_id, args, adjoint, m);
    submit_gate(QS(h), m, ctrl, n_ctrl, target);
    QGB_CATCH
}

int qgb_qproc_apply_swap(qgb_handle qp, qgb_handle h, int lane_a, int lane_b) {
    QGB_TRY
    QP(qp);
    QStates *qs = QS(h);
    check_allocated(qs);
    check_lane(qs, lane_a);
    check_lane(qs, lane_b);
    std::swap(qs->perm[(size_t)lane_a], qs->perm[(size_t)lane_b]); /* gates queued so far keep their index bits */
    g.stats.gates_submitted += 1;
    g.stats.native_swaps += 1;
    QGB_CATCH
}

int qgb_qproc_apply_pauli_expi(qgb_handle qp, qgb_handle h, double theta, const int *lanes, const int *paulis,
                               int n, const int *ctrl, int n_ctrl) {
    QGB_TRY
    QP(qp);
    QStates *qs = QS(h);
    check_allocated(qs);
    if (n < 0 || n > QGB_MAX_LANES) fail(QGB_ERR_INVALID, "bad number of Pauli factors.");
    uint64_t zmask = 0;
    for (int i = 0; i < n; ++i) {
        check_lane(qs, lanes[i]);
        if (paulis[i] < 0 || paulis[i] > 3) fail(QGB_ERR_INVALID, "Pauli code %d (0 = I, 1 = X, 2 = Y, 3 = Z).", paulis[i]);
        const uint64_t bit = 1ull << qs->phys(lanes[i]);
        if (zmask & bit) fail(QGB_ERR_INVALID, "lane %d appears twice in the Pauli string.", lanes[i]);
        if (paulis[i] != 0) zmask |= bit;
    }
    const uint64_t cmask = control_mask(qs, ctrl, n_ctrl, zmask);
    /* exp(i theta P) = V exp(i theta Z..Z) V^+, V = product over the X factors of H and over the Y
     * factors of S H (H Z H = X, S X S^+ = Y); the basis changes need no controls (they cancel
     * where the middle does not act) and merge into the neighbouring gates of their lanes */
    double mh[8], ms[8], msd[8];
    gate_matrix(7 /* H */, nullptr, 0, mh);
    gate_matrix(8 /* S */, nullptr, 0, ms);
    gate_matrix(8 /* S */, nullptr, 1, msd);
    for (int i = 0; i < n; ++i) {
        if (paulis[i] == 2) submit_gate(qs, msd, nullptr, 0, lanes[i]);
        if (paulis[i] == 1 || paulis[i] == 2) submit_gate(qs, mh, nullptr, 0, lanes[i]);
    }
    Gate gt;
    for (int e = 0; e < 8; ++e) gt.m[e] = 0.;
    gt.m[0] = std::cos(theta), gt.m[1] = std::sin(theta);
    if (zmask) {
        gt.m[6] = std::cos(theta), gt.m[7] = -std::sin(theta);
        gt.target = __builtin_ctzll(zmask);
        gt.parity = zmask;
        gt.ctrl_mask = cmask;
    } else {
        /* the identity string: exp(i theta) on every amplitude the controls select */
        gt.m[6] = gt.m[0], gt.m[7] = gt.m[1];
        int free_lane = 0;
        while (free_lane < qs->n_lanes && ((cmask >> free_lane) & 1ull)) ++free_lane;
        if (free_lane < qs->n_lanes) {
            gt.target = free_lane;
            gt.ctrl_mask = cmask;
        } else { /* every lane is a control: one of them becomes the target of diag(1, e^{i theta}) */
            gt.target = 0;
            gt.ctrl_mask = cmask & ~1ull;
            gt.m[0] = 1., gt.m[1] = 0.;
        }
    }
    submit_queued(qs, gt, n_ctrl);
    for (int i = n - 1; i >= 0; --i) {
        if (paulis[i] == 1 || paulis[i] == 2) submit_gate(qs, mh, nullptr, 0, lanes[i]);
        if (paulis[i] == 2) submit_gate(qs, ms, nullptr, 0, lanes[i]);
    }
    g.stats.native_pauli_exps += 1;
    QGB_CATCH
}

int qgb_qproc_apply_gates_batch(qgb_handle qp, qgb_handle h, const qgb_gate_op *ops, int64_t n_ops) {
    QGB_TRY
    QP(qp);
    QStates *qs = QS(h);
    check_allocated(qs);
    if (n_ops < 0) fail(QGB_ERR_INVALID, "negative number of gates.");
    for (int64_t i = 0; i < n_ops; ++i) {
        const qgb_gate_op &op = ops[i];
        double m[8];
        if (op.gate_id == QGB_GATE_MATRIX) {
            std::memcpy(m, op.mat8, sizeof(m));
        } else {
            const int want = gate_matrix_n_args(op.gate_id);
            if (want < 0) fail(QGB_ERR_RUNTIME, "Unknown gate type.");
            gate_matrix(op.gate_id, op.args, op.adjoint, m);
        }
        check_lane(qs, op.target);
        int ctrl[64];
        int n_ctrl = 0;
        for (int lane = 0; lane < 64; ++lane)
            if ((op.ctrl_mask >> lane) & 1ull) {
                if (lane >= qs->n_lanes) fail(QGB_ERR_INVALID, "control lane %d out of range [0, %d).", lane, qs->n_lanes);
                ctrl[n_ctrl++] = lane;
            }
        submit_gate(qs, m, ctrl, n_ctrl, op.target);
    }
    QGB_CATCH
}

int qgb_qproc_flush(qgb_handle qp, qgb_handle h) {
    QGB_TRY
    QP(qp);
    require_init();
    flush(QS(h));
    QGB_CATCH
}

/* ---- QubitsStatesGetter ------------------------------------------------------------------------ */

int qgb_getter_new(int prec, qgb_handle *out) {
    QGB_TRY
    check_prec(prec);
    Getter *gt = new Getter();
    gt->prec = prec;
    g.getters.insert(gt);
    *out = reinterpret_cast<qgb_handle>(gt);
    QGB_CATCH
}

int qgb_getter_delete(qgb_handle h) {
    QGB_TRY
    Getter *gt = QG(h);
    g.getters.erase(gt);
    delete gt;
    QGB_CATCH
}

int qgb_getter_get_states(qgb_handle getter, void *array, int64_t array_offset, int mathop,
                          const int *lane_tables, const int *n_per, int64_t empty_lane_mask,
                          const qgb_handle *qstates_list, int n_qstates, int n_ext_lanes, int64_t n_states,
                          int64_t start, int64_t step) {
    QGB_TRY
    Getter *gt = QG(getter);
    /* range checks of glue.cpp:465-478 */
    if (n_ext_lanes < 0 || n_ext_lanes > 62) fail(QGB_ERR_INVALID, "value out of range");
    if (mathop != QGB_MATHOP_NULL && mathop != QGB_MATHOP_PROB) fail(QGB_ERR_RUNTIME, "unknown math op.");
    if (n_states <= 0) return QGB_OK; /* nothing to read: no range to check */
    const int64_t space = (int64_t)1 << n_ext_lanes;
    if (start < 0 || space <= start) fail(QGB_ERR_INVALID, "value out of range");
    const int64_t end = start + step * (n_states - 1);
    if (end < 0 || space <= end) fail(QGB_ERR_INVALID, "value out of range");
    const size_t real_size = gt->prec == QGB_PREC_FP64 ? 8 : 4;
    const size_t item = (mathop == QGB_MATHOP_NULL) ? 2 * real_size : real_size;
    char *dst = static_cast<char *>(array) + array_offset * item;
    if (n_qstates == 0) {
        /* no qstates at all: |0...0> (glue.cpp:481-496) — a host-side constant, no device work */
        std::memset(dst, 0, (size_t)n_states * item);
        if (start == 0) {
            if (real_size == 8)
                *reinterpret_cast<double *>(dst) = 1.;
            else
                *reinterpret_cast<float *>(dst) = 1.f;
        }
        return QGB_OK;
    }
    require_init();
    GatherParams gp;
    build_gather(gp, gt->prec, lane_tables, n_per, qstates_list, n_qstates, n_ext_lanes);
    /* kernel -> device staging -> pinned buffer (async) -> caller's array; the copy of chunk i to the
     * caller overlaps the kernel + D2H of chunk i+1 */
    const int64_t per_chunk = (int64_t)(g.stage_bytes / 2 / item);
    int64_t done = 0;
    int buf = 0;
    int64_t pending_count[2] = {0, 0};
    int64_t pending_first[2] = {0, 0};
    auto drain = [&](int b) {
        if (pending_count[b] == 0) return;
        CUDA_CHECK(cudaEventSynchronize(g.stage_done[b]));
        std::memcpy(dst + pending_first[b] * item, g.h_stage[b], (size_t)pending_count[b] * item);
        g.stats.d2h_bytes += pending_count[b] * (int64_t)item;
        pending_count[b] = 0;
    };
    try {
    while (done < n_states) {
        const int64_t count = std::min(per_chunk, n_states - done);
        drain(buf);
        char *d_half = static_cast<char *>(g.d_stage) + (size_t)buf * (g.stage_bytes / 2);
        CUDA_CHECK(launch_get_states(gt->prec, d_half, mathop, gp, (uint64_t)empty_lane_mask, done, count,
                                     start, step, g.stream));
        g.stats.kernel_launches += 1;
        CUDA_CHECK(cudaMemcpyAsync(g.h_stage[buf], d_half, (size_t)count * item, cudaMemcpyDeviceToHost,
                                   g.stream));
        CUDA_CHECK(cudaEventRecord(g.stage_done[buf], g.stream));
        pending_first[buf] = done;
        pending_count[buf] = count;
        done += count;
        buf ^= 1;
    }
    drain(buf);
    drain(buf ^ 1);
    } catch (...) {
        release_gather(gp);
        throw;
    }
    release_gather(gp);
    QGB_CATCH
}

int qgb_getter_prepare_prob_array(qgb_handle getter, void *prob, const int *lane_tables, const int *n_per,
                                  const qgb_handle *qstates_list, int n_qstates, int n_lanes, int n_hidden) {
    QGB_TRY
    Getter *gt = QG(getter);
    require_init();
    double *d_prob = device_prob_array(gt->prec, lane_tables, n_per, qstates_list, n_qstates, n_lanes,
                                       n_hidden);
    const int64_t count = (int64_t)1 << n_lanes;
    if (gt->prec == QGB_PREC_FP64) {
        copy_to_host(prob, d_prob, sizeof(double) * (size_t)count);
    } else {
        float *d_f = static_cast<float *>(g.pool.alloc(sizeof(float) * (size_t)count));
        CUDA_CHECK(launch_cast_from_double(gt->prec, d_f, d_prob, count, g.stream));
        g.stats.kernel_launches += 1;
        copy_to_host(prob, d_f, sizeof(float) * (size_t)count);
        g.pool.release(d_f);
    }
    g.pool.release(d_prob);
    QGB_CATCH
}

/* Fills the pool's source: the marginal probability vector in d_cum, or — one state vector, every
 * lane in the pool, index order, default (float64) scan — just a pointer to the amplitudes. */
static void pool_set_source(Pool *p, int prec, const int *lane_tables, const int *n_per, const qgb_handle *list,
                            int n_qstates, int n_lanes, int n_hidden) {
    if (n_qstates == 1 && n_hidden == 0 && g.opt.pool_compat_workers <= 0 && g.opt.pool_from_amplitudes) {
        QStates *qs = QS(list[0]);
        check_allocated(qs);
        bool identity = qs->prec == prec && qs->n_lanes == n_lanes && n_per[0] == n_lanes;
        for (int l = 0; identity && l < n_lanes; ++l) identity = lane_tables[l] == qs->phys(l);
        if (identity) {
            flush(qs);
            p->d_cum = static_cast<double *>(g.pool.alloc(sizeof(double) << n_lanes));
            p->src_amp = qs->d_amp;
            return;
        }
    }
    p->d_cum = device_prob_array(prec, lane_tables, n_per, list, n_qstates, n_lanes, n_hidden);
}

/* scan phases 1+2 over d_prob (2^n_lanes doubles): leaves the block sums in pool->d_sums and
 * returns the total (host value) */
static double pool_scan_partial(Pool *p) {
    const int64_t n = (int64_t)1 << p->n_lanes;
    const int64_t n_blocks = (n + 4095) / 4096;
    p->d_sums = static_cast<double *>(g.pool.alloc(sizeof(double) * (size_t)(n_blocks + 1)));
    double *d_total = p->d_sums + n_blocks;
    CUDA_CHECK(launch_scan_phase1(p->prec, p->src_amp, p->d_cum, n, p->d_sums, g.stream));
    CUDA_CHECK(launch_scan_phase2(p->d_sums, n_blocks, d_total, g.stream));
    CUDA_CHECK(cudaMemcpyAsync(g.h_scalar, d_total, sizeof(double), cudaMemcpyDeviceToHost, g.stream));
    stream_sync();
    g.stats.kernel_launches += 2;
    g.stats.d2h_bytes += sizeof(double);
    return g.h_scalar[0];
}

/* phase 3: cum[i] = (offset + scan[i]) * (1 / total) */
static void pool_scan_finalize(Pool *p, double offset, double total) {
    if (p->finalized || !p->d_sums) fail(QGB_ERR_RUNTIME, "sampling pool is already finalized.");
    const int64_t n = (int64_t)1 << p->n_lanes;
    CUDA_CHECK(launch_scan_phase3(p->prec, p->src_amp, p->d_cum, n, p->d_sums, offset, 1. / total, g.stream));
    p->src_amp = nullptr;
    g.stats.kernel_launches += 1;
    g.pool.release(p->d_sums); /* stream-ordered: reused only by later work on the same stream */
    p->d_sums = nullptr;
    p->finalized = true;
}

} /* extern "C" */
/* The reference-compatible build (option pool_compat_workers = W): W sequential spans in the state
 * precision V, span totals prefix-summed in V on the host (W values), then (p + offset) * (V(1) / sum)
 * — CPUSamplingPool.cpp:13-47, Parallel.h:12-23, Parallel.cpp:60-67 (one worker up to 2^16 entries).
 * Returns the total the reference checks against 1. */
template <typename V>
static double pool_scan_compat(Pool *p, int workers) {
    const int64_t n = (int64_t)1 << p->n_lanes;
    const int w = n > ((int64_t)1 << 16) ? workers : 1;
    int64_t span = (n + w - 1) / w;
    span = ((span + 15) / 16) * 16;
    if (w > 2048) fail(QGB_ERR_INVALID, "pool_compat_workers is limited to 2048.");
    double *d_tot = g.d_partials; /* 4096 doubles: totals in [0, w), offsets in [2048, 2048 + w) */
    CUDA_CHECK(launch_compat_scan(p->prec, p->d_cum, n, span, w, d_tot, g.stream));
    std::vector<double> tot((size_t)w);
    CUDA_CHECK(cudaMemcpyAsync(tot.data(), d_tot, sizeof(double) * (size_t)w, cudaMemcpyDeviceToHost, g.stream));
    stream_sync();
    V v = V(0);
    for (int i = 0; i < w; ++i) { /* CPUSamplingPool.cpp:26-30 */
        v += (V)tot[(size_t)i];
        tot[(size_t)i] = (double)v;
    }
    const V sum = v;
    if (0.05 < std::abs(sum - V(1.))) return (double)sum; /* the caller raises (CPUSamplingPool.cpp:31-34) */
    const V norm = V(1.) / sum;
    CUDA_CHECK(cudaMemcpyAsync(d_tot + 2048, tot.data(), sizeof(double) * (size_t)w, cudaMemcpyHostToDevice, g.stream));
    CUDA_CHECK(launch_compat_apply(p->prec, p->d_cum, n, span, d_tot + 2048, (double)norm, g.stream));
    stream_sync(); /* `tot` is the source of an asynchronous copy */
    g.stats.kernel_launches += 2;
    g.stats.d2h_bytes += (int64_t)sizeof(double) * w;
    p->finalized = true;
    return (double)sum;
}
extern "C" {

static void pool_destroy(Pool *p);
static void pool_check_total(Pool *p, double total);

/* scan + normalise the marginal vector in p->d_cum; destroys the pool and raises on a bad total */
static void pool_build(Pool *p) {
    try {
        if (g.opt.pool_compat_workers > 0) {
            const int w = (int)std::min<int64_t>(g.opt.pool_compat_workers, 1 << 20);
            const double total = p->prec == QGB_PREC_FP64 ? pool_scan_compat<double>(p, w) : pool_scan_compat<float>(p, w);
            pool_check_total(p, total);
            return;
        }
        const double total = pool_scan_partial(p);
        pool_check_total(p, total);
        pool_scan_finalize(p, 0., total);
    } catch (...) {
        if (g.pools.count(p)) pool_destroy(p);
        throw;
    }
}

static void pool_set_empty_lanes(Pool *p, const int *empty_lanes, int n_empty) {
    if (n_empty < 0 || n_empty > QGB_MAX_LANES) fail(QGB_ERR_INVALID, "bad number of empty lanes.");
    std::vector<int> sorted(empty_lanes, empty_lanes + n_empty);
    std::sort(sorted.begin(), sorted.end());
    p->empty.n = 0;
    for (int v : sorted) {
        if (v < 0 || v >= 63) fail(QGB_ERR_INVALID, "bad empty lane %d.", v);
        p->empty.pos[p->empty.n++] = (int8_t)v;
    }
}

static void pool_destroy(Pool *p) {
    if (!g.pools.count(p)) return;
    if (p->d_cum) g.pool.release(p->d_cum);
    if (p->d_sums) g.pool.release(p->d_sums);
    g.pools.erase(p);
    delete p;
}

/* CPUSamplingPool.cpp:31-34 */
static void pool_check_total(Pool *p, double total) {
    if (!(std::fabs(total - 1.) <= 0.05)) {
        pool_destroy(p);
        fail(QGB_ERR_RUNTIME, "error in probability sum is beyond 0.05., %g.", total);
    }
}

int qgb_getter_create_sampling_pool(qgb_handle getter, const int *lane_tables, const int *n_per,
                                    const qgb_handle *qstates_list, int n_qstates, int n_lanes, int n_hidden,
                                    const int *empty_lanes, int n_empty, qgb_handle *out) {
    QGB_TRY
    Getter *gt = QG(getter);
    require_init();
    if (n_empty < 0 || n_empty > QGB_MAX_LANES) fail(QGB_ERR_INVALID, "bad number of empty lanes.");
    Pool *p = new Pool();
    p->prec = gt->prec;
    p->n_lanes = n_lanes;
    p->empty.n = 0;
    g.pools.insert(p);
    try {
        pool_set_source(p, gt->prec, lane_tables, n_per, qstates_list, n_qstates, n_lanes, n_hidden);
        pool_set_empty_lanes(p, empty_lanes, n_empty);
    } catch (...) {
        pool_destroy(p);
        throw;
    }
    pool_build(p);
    *out = reinterpret_cast<qgb_handle>(p);
    QGB_CATCH
}

int qgb_getter_create_sampling_pool_partial(qgb_handle getter, const int *lane_tables, const int *n_per,
                                            const qgb_handle *qstates_list, int n_qstates, int n_lanes,
                                            int n_hidden, qgb_handle *out, double *local_total) {
    QGB_TRY
    Getter *gt = QG(getter);
    require_init();
    Pool *p = new Pool();
    p->prec = gt->prec;
    p->n_lanes = n_lanes;
    p->empty.n = 0;
    g.pools.insert(p);
    try {
        pool_set_source(p, gt->prec, lane_tables, n_per, qstates_list, n_qstates, n_lanes, n_hidden);
    } catch (...) {
        pool_destroy(p);
        throw;
    }
    try {
        *local_total = pool_scan_partial(p);
    } catch (...) {
        pool_destroy(p);
        throw;
    }
    *out = reinterpret_cast<qgb_handle>(p);
    QGB_CATCH
}

int qgb_pool_finalize(qgb_handle pool, double offset, double total) {
    QGB_TRY
    Pool *p = SP(pool);
    require_init();
    if (!(total > 0.)) fail(QGB_ERR_INVALID, "total probability must be positive.");
    pool_scan_finalize(p, offset, total);
    QGB_CATCH
}

int qgb_pool_from_prob_array(int prec, const double *prob, int n_lanes, const int *empty_lanes, int n_empty,
                             qgb_handle *out) {
    QGB_TRY
    check_prec(prec);
    require_init();
    if (n_lanes < 0 || n_lanes > QGB_MAX_LANES) fail(QGB_ERR_INVALID, "bad lane count %d.", n_lanes);
    Pool *p = new Pool();
    p->prec = prec;
    p->n_lanes = n_lanes;
    p->empty.n = 0;
    g.pools.insert(p);
    try {
        const size_t bytes = sizeof(double) << n_lanes;
        p->d_cum = static_cast<double *>(g.pool.alloc(bytes));
        CUDA_CHECK(cudaMemcpyAsync(p->d_cum, prob, bytes, cudaMemcpyHostToDevice, g.stream));
        g.stats.h2d_bytes += (int64_t)bytes;
        pool_set_empty_lanes(p, empty_lanes, n_empty);
    } catch (...) {
        pool_destroy(p);
        throw;
    }
    pool_build(p);
    *out = reinterpret_cast<qgb_handle>(p);
    QGB_CATCH
}

/* ---- SamplingPool ------------------------------------------------------------------------------ */

int qgb_pool_sample(qgb_handle pool, int64_t *obs, int n_samples, const double *randnum) {
    QGB_TRY
    Pool *p = SP(pool);
    require_init();
    if (!p->d_cum) fail(QGB_ERR_RUNTIME, "sampling pool was released.");
    if (!p->finalized) fail(QGB_ERR_RUNTIME, "sampling pool is not finalized.");
    if (n_samples < 0) fail(QGB_ERR_INVALID, "negative number of samples.");
    if (n_samples == 0) return QGB_OK;
    const size_t n = (size_t)n_samples;
    char *d_buf = static_cast<char *>(g.pool.alloc(n * 16));
    double *d_rand = reinterpret_cast<double *>(d_buf);
    int64_t *d_obs = reinterpret_cast<int64_t *>(d_buf + n * 8);
    CUDA_CHECK(cudaMemcpyAsync(d_rand, randnum, n * 8, cudaMemcpyHostToDevice, g.stream));
    CUDA_CHECK(launch_sample(p->prec, p->d_cum, p->n_lanes, d_rand, d_obs, n_samples, p->empty, g.stream));
    CUDA_CHECK(cudaMemcpyAsync(obs, d_obs, n * 8, cudaMemcpyDeviceToHost, g.stream));
    stream_sync();
    g.pool.release(d_buf);
    g.stats.kernel_launches += 1;
    g.stats.h2d_bytes += (int64_t)n * 8;
    g.stats.d2h_bytes += (int64_t)n * 8;
    QGB_CATCH
}

int qgb_pool_sample_sequential(qgb_handle pool, int64_t *obs, int n_shots, const double *randnum) {
    QGB_TRY
    Pool *p = SP(pool);
    require_init();
    if (!p->d_cum || !p->finalized) fail(QGB_ERR_RUNTIME, "sampling pool is not ready.");
    if (n_shots < 0) fail(QGB_ERR_INVALID, "negative number of shots.");
    if (n_shots == 0 || p->n_lanes == 0) {
        for (int i = 0; i < n_shots; ++i) obs[i] = 0;
        return QGB_OK;
    }
    const size_t n_rand = (size_t)n_shots * (size_t)p->n_lanes;
    char *d_buf = static_cast<char *>(g.pool.alloc(n_rand * 8 + (size_t)n_shots * 8));
    double *d_rand = reinterpret_cast<double *>(d_buf);
    int64_t *d_obs = reinterpret_cast<int64_t *>(d_buf + n_rand * 8);
    CUDA_CHECK(cudaMemcpyAsync(d_rand, randnum, n_rand * 8, cudaMemcpyHostToDevice, g.stream));
    CUDA_CHECK(launch_sample_sequential(p->d_cum, p->n_lanes, d_rand, d_obs, n_shots, g.stream));
    CUDA_CHECK(cudaMemcpyAsync(obs, d_obs, (size_t)n_shots * 8, cudaMemcpyDeviceToHost, g.stream));
    stream_sync();
    g.pool.release(d_buf);
    g.stats.kernel_launches += 1;
    g.stats.h2d_bytes += (int64_t)n_rand * 8;
    g.stats.d2h_bytes += (int64_t)n_shots * 8;
    QGB_CATCH
}

int qgb_pool_delete(qgb_handle pool) {
    QGB_TRY
    pool_destroy(SP(pool));
    QGB_CATCH
}

/* ---- instrumentation ------------------------------------------------------------------------------ */

int qgb_stats_get(qgb_stats *out) {
    tma_pass_phase_report(); /* diagnostic builds only */
    *out = g.stats;
    return QGB_OK;
}

int qgb_stats_reset(void) {
    std::memset(&g.stats, 0, sizeof(g.stats));
    return QGB_OK;
}

int qgb_set_option(const char *name, int64_t value) {
    QGB_TRY
    const std::string k(name ? name : "");
    if (k == "fuse") g.opt.fuse = value;
    else if (k == "merge") g.opt.merge = value;
    else if (k == "fans") g.opt.fans = value;
    else if (k == "fan_cost") g.opt.fan_cost = value;
    else if (k == "lazy_reset") g.opt.lazy_reset = value;
    else if (k == "pin_exported") g.pool.pin_exported = value != 0;
    else if (k == "queue_stream") g.opt.queue_stream = value;
    else if (k == "stream_hi") g.opt.stream_hi = value;
    else if (k == "stream_keep") g.opt.stream_keep = value;
    else if (k == "exact") g.opt.exact = value;
    else if (k == "tile_lanes_fp64") g.opt.tile_lanes_fp64 = value;
    else if (k == "tile_lanes_fp32") g.opt.tile_lanes_fp32 = value;
    else if (k == "low_lanes_fp64") g.opt.low_lanes_fp64 = value;
    else if (k == "low_lanes_fp32") g.opt.low_lanes_fp32 = value;
    else if (k == "low_lanes") g.opt.low_lanes_fp64 = g.opt.low_lanes_fp32 = value;
    else if (k == "max_gates_per_pass") g.opt.max_gates_per_pass = value;
    else if (k == "max_cost") g.opt.max_cost = value;
    else if (k == "lookahead") g.opt.lookahead = value;
    else if (k == "queue_limit") g.opt.queue_limit = value;
    else if (k == "tile_buffers" || k == "tma") { /* accepted for old scripts: TMA staging is the only path */ }
    else if (k == "ctas_per_sm") g.opt.ctas_per_sm = value;
    else if (k == "shear") g.opt.shear_fp64 = g.opt.shear_fp32 = value;
    else if (k == "shear_fp64") g.opt.shear_fp64 = value;
    else if (k == "shear_fp32") g.opt.shear_fp32 = value;
    else if (k == "l2_hint") tma_pass_set_l2_hint((int)value);
    else if (k == "debug_pass_mode") tma_pass_set_debug_mode((int)value);
    else if (k == "tma_buffers") g.opt.tma_buffers = value;
    else if (k == "reg_bits_fp64") g.opt.reg_bits_fp64 = value;
    else if (k == "warp_local") g.opt.warp_local = value;
    else if (k == "pool_compat_workers") g.opt.pool_compat_workers = value;
    else if (k == "pool_from_amplitudes") g.opt.pool_from_amplitudes = value;
    else if (k == "tma_ws") tma_pass_set_warp_specialised((int)value);
    else fail(QGB_ERR_INVALID, "unknown option '%s'.", k.c_str());
    QGB_CATCH
}

} // extern "C"
