/*
 * qgate_b200.h — C ABI of the B200-native state-vector runtime.
 *
 * This is the drop-in boundary for qgate's native runtime layer.  Every entry
 * point replaces one function of the reference's CPython glue / device
 * extension (paths relative to /root/reference):
 *
 *   qgate/simulator/src/glue.cpp:605-631     22 METH_VARARGS functions over the
 *                                            4 abstract classes of
 *   qgate/simulator/src/Interfaces.h:7-81    QubitStates / QubitProcessor /
 *                                            QubitsStatesGetter / SamplingPool
 *   qgate/simulator/src/cudaext.cpp:127-136  devices_initialize, devices_clear,
 *                                            qubit_states_new, qubit_processor_new,
 *                                            qubits_states_getter_new
 *
 * Rules of the ABI
 *   - extern "C", plain pointers and sizes, no CPython / NumPy / torch types.
 *   - every call returns an int status (QGB_OK == 0).  On failure the message
 *     is available from qgb_last_error() (thread-local).  The reference lets
 *     C++ exceptions escape glue.cpp (no try/catch) and abort()s on broken
 *     invariants; this ABI never throws and never aborts.
 *   - handles are opaque 64-bit values (the reference boxes raw pointers in
 *     numpy.uint64 scalars, glue.cpp:84-88,154-177); the caller owns them and
 *     releases them with the matching *_delete.
 *   - `prec` follows the reference's enum Precision (Types.h:56-60).
 *   - all host buffers are caller-allocated and contiguous; amplitudes are
 *     interleaved (re, im) in the state precision.
 *
 * Two libraries implement this header:
 *   qgate_b200/lib/libqgate_b200.so   the product: hand-written sm_100a CUDA.
 *   oracle/_ref/libqgate_ref_cpu.so   test infrastructure: the UNMODIFIED
 *                                     reference CPU runtime compiled from
 *                                     /root/reference behind the same ABI
 *                                     (oracle/ref_shim.cpp).
 */
#ifndef QGATE_B200_H
#define QGATE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t qgb_handle;

/* status codes.  INVALID maps to Python ValueError, everything else to RuntimeError
 * (glue.cpp:465-478 raises ValueError for range errors, RuntimeError otherwise). */
#define QGB_OK            0
#define QGB_ERR_INVALID   1
#define QGB_ERR_RUNTIME   2
#define QGB_ERR_OOM       3
#define QGB_ERR_CUDA      4

/* enum Precision, Types.h:56-60 */
#define QGB_PREC_FP64     1
#define QGB_PREC_FP32     2

/* enum MathOp, Types.h:62-65 */
#define QGB_MATHOP_NULL   0
#define QGB_MATHOP_PROB   1

/* gate ids for qgb_gate_matrix / qgb_qproc_apply_gate_typed, one per matrix
 * factory registered in glue.cpp:62-81 (names in glue.cpp:100-131). */
#define QGB_GATE_U        0   /* U(theta, phi, lambda)  GateMatrix.cpp:13-26 */
#define QGB_GATE_U2       1   /* U2(phi, lambda)        GateMatrix.cpp:28-38 */
#define QGB_GATE_U1       2   /* U1(lambda)             GateMatrix.cpp:40-47 */
#define QGB_GATE_ID       3
#define QGB_GATE_X        4
#define QGB_GATE_Y        5
#define QGB_GATE_Z        6
#define QGB_GATE_H        7
#define QGB_GATE_S        8
#define QGB_GATE_T        9
#define QGB_GATE_RX       10
#define QGB_GATE_RY       11
#define QGB_GATE_RZ       12
#define QGB_GATE_EXPII    13
#define QGB_GATE_EXPIZ    14
#define QGB_GATE_SH       15
#define QGB_N_GATES       16

/* ---- library ------------------------------------------------------------ */

/* message of the last failed call on this thread ("" if none). */
const char *qgb_last_error(void);
/* "cuda-sm_100a" for the product, "reference-cpu" for the oracle shim. */
const char *qgb_backend_name(void);
/* ABI revision; bumped when a signature changes. */
int qgb_abi_version(void);

/* ---- devices  (cudaext.cpp:24-72 devices_initialize / devices_clear) ----- */

/* device_ids: CUDA ordinals to use (n == 0: current device).  One id = one
 * shard owner.  max_po2idx_per_chunk / memory_store_size keep the reference's
 * meaning (log2 bytes of the largest single allocation, byte budget per
 * device; -1 = automatic). */
int qgb_devices_initialize(const int *device_ids, int n_device_ids,
                           int max_po2idx_per_chunk, int64_t memory_store_size);
int qgb_devices_clear(void);
/* number of visible devices; 0 for the CPU shim. */
int qgb_device_count(int *count);
/* run all work of this library on an externally owned CUDA stream
 * (cudaStream_t cast to uint64; 0 restores the library's own stream).
 * Lets the caller bracket launches with its own CUDA events. */
int qgb_set_stream(uint64_t cuda_stream);

/* ---- gate matrices  (GateMatrix.cpp:13-158, glue.cpp:382-431) ------------ */

/* mat8 receives (re,im) of m00, m01, m10, m11 in double; adjoint != 0 applies
 * the conjugate transpose (GateMatrix.cpp:160-168). */
int qgb_gate_matrix(int gate_id, const double *args, int n_args, int adjoint, double *mat8);

/* ---- QubitStates  (Interfaces.h:7-19) ------------------------------------ */

int qgb_qstates_new(int prec, qgb_handle *out);                 /* cudaext.cpp:74-91  */
int qgb_qstates_delete(qgb_handle qstates);                     /* glue.cpp:191-203   */
int qgb_qstates_deallocate(qgb_handle qstates);                 /* glue.cpp:179-189   */
int qgb_qstates_get_n_lanes(qgb_handle qstates, int *n_lanes);  /* glue.cpp:231-240   */

/* ---- QubitProcessor  (Interfaces.h:22-50) -------------------------------- */

int qgb_qproc_new(int prec, qgb_handle *out);                   /* cudaext.cpp:93-109 */
int qgb_qproc_delete(qgb_handle qproc);                         /* glue.cpp:205-216   */
int qgb_qproc_synchronize(qgb_handle qproc);                    /* glue.cpp:243-253   */
int qgb_qproc_reset(qgb_handle qproc);                          /* glue.cpp:255-265   */
int qgb_qproc_initialize_qstates(qgb_handle qproc, qgb_handle qstates, int n_lanes); /* glue.cpp:267-285 */
int qgb_qproc_reset_qstates(qgb_handle qproc, qgb_handle qstates);                   /* glue.cpp:287-298 */
/* P(lane == 0), glue.cpp:300-311 */
int qgb_qproc_calc_probability(qgb_handle qproc, qgb_handle qstates, int local_lane, double *prob);
/* Kronecker join; LAST list element gets the lowest lanes, new |0> lanes on top. glue.cpp:313-335 */
int qgb_qproc_join(qgb_handle qproc, qgb_handle dst, const qgb_handle *src_list, int n_src,
                   int n_new_lanes);
int qgb_qproc_decohere(qgb_handle qproc, int value, double prob, qgb_handle qstates,
                       int local_lane);                          /* glue.cpp:337-349 */
int qgb_qproc_decohere_and_separate(qgb_handle qproc, int value, double prob,
                                    qgb_handle qstates0, qgb_handle qstates1,
                                    qgb_handle qstates, int local_lane); /* glue.cpp:351-366 */
int qgb_qproc_apply_reset(qgb_handle qproc, qgb_handle qstates, int local_lane); /* glue.cpp:368-379 */
/* mat8 as produced by qgb_gate_matrix.  glue.cpp:382-405 */
int qgb_qproc_apply_gate(qgb_handle qproc, const double *mat8, qgb_handle qstates, int local_lane);
/* glue.cpp:408-431 */
int qgb_qproc_apply_controlled_gate(qgb_handle qproc, const double *mat8, qgb_handle qstates,
                                    const int *control_lanes, int n_controls, int target_lane);
/* one call per gate from the host: matrix factory + adjoint + (controlled) apply,
 * i.e. exactly what glue.cpp:382-431 does from (gate_type.cmatf, gate_type.args, adjoint). */
int qgb_qproc_apply_gate_typed(qgb_handle qproc, int gate_id, const double *args, int n_args,
                               int adjoint, qgb_handle qstates,
                               const int *control_lanes, int n_controls, int target_lane);

/* ---- QubitsStatesGetter  (Interfaces.h:53-72) ----------------------------- */

int qgb_getter_new(int prec, qgb_handle *out);                  /* cudaext.cpp:111-125 */
int qgb_getter_delete(qgb_handle getter);                       /* glue.cpp:218-229    */
/* lane_tables: the per-qstates local->external lane lists concatenated,
 * n_lanes_per_qstates[i] entries for qstates_list[i] (glue.cpp:447-506).
 * array: complex<real> (NULL op) or real (PROB op), written at [array_offset, +n_states).
 * Value j = prod_qstates op(amp[perm(start + step*j)]), 0 if an empty-lane bit is set. */
int qgb_getter_get_states(qgb_handle getter, void *array, int64_t array_offset, int mathop,
                          const int *lane_tables, const int *n_lanes_per_qstates,
                          int64_t empty_lane_mask,
                          const qgb_handle *qstates_list, int n_qstates,
                          int n_ext_lanes, int64_t n_states, int64_t start, int64_t step);
/* marginal probability vector, 2^n_lanes reals; lanes with external index >= n_lanes
 * (hidden in MSB) or < n_hidden (hidden in LSB) are summed out as the tables say.
 * The tables carry external positions in [0, n_lanes + n_hidden_lanes) with the
 * hidden lanes in the LOW positions (cpuruntime.py:20-24 convention), glue.cpp:508-536. */
int qgb_getter_prepare_prob_array(qgb_handle getter, void *prob,
                                  const int *lane_tables, const int *n_lanes_per_qstates,
                                  const qgb_handle *qstates_list, int n_qstates,
                                  int n_lanes, int n_hidden_lanes);
/* glue.cpp:538-559 */
int qgb_getter_create_sampling_pool(qgb_handle getter,
                                    const int *lane_tables, const int *n_lanes_per_qstates,
                                    const qgb_handle *qstates_list, int n_qstates,
                                    int n_lanes, int n_hidden_lanes,
                                    const int *empty_lanes, int n_empty_lanes,
                                    qgb_handle *pool);

/* ---- SamplingPool  (Interfaces.h:75-81) ----------------------------------- */

/* obs[i] = deposit(upper_bound(cumprob, (real)randnum[i])) around the empty lanes.
 * glue.cpp:561-589 (n_samples is a C int there too). */
int qgb_pool_sample(qgb_handle pool, int64_t *obs, int n_samples, const double *randnum);
int qgb_pool_delete(qgb_handle pool);                           /* glue.cpp:591-602 */

/* ---- sharding support ------------------------------------------------------------
 * The reference shares the chunks of one state vector between the devices of ONE
 * process through peer pointers (MultiChunkPtr.h:22-41, CUDADevice.cpp:135-142).  Here one
 * process drives one GPU and owns one shard of 2^(n - g) amplitudes (the g high-order
 * "global" lanes select the rank); the host layer (qgate_b200/dist.py) moves lanes between
 * the shards.  These entry points are what it needs from the native layer. */

/* address of the amplitude array (device memory for the CUDA library, host memory for the
 * CPU shim) and its size; queued gates are flushed first. */
int qgb_qstates_data_ptr(qgb_handle qstates, uint64_t *ptr, int64_t *bytes);
/* a second buffer of the same size (allocated on first use) for out-of-place exchanges,
 * and the switch that makes it the amplitude array (the old array becomes the spare). */
int qgb_qstates_alt_buffer(qgb_handle qstates, uint64_t *ptr);
int qgb_qstates_flip(qgb_handle qstates);
/* sum of |a|^2 over this array (a shard's share of the norm). */
int qgb_qproc_calc_norm(qgb_handle qproc, qgb_handle qstates, double *norm);
/* join for a sharded destination: dst holds elements [index_offset, index_offset + 2^n_dst)
 * of the n_total_lanes-lane Kronecker product that qgb_qproc_join would build. */
int qgb_qproc_join_shard(qgb_handle qproc, qgb_handle dst, const qgb_handle *src_list, int n_src,
                         int n_new_lanes, int n_total_lanes, int64_t index_offset);
/* two-step sampling pool for a sharded probability vector: `partial` builds the local
 * (un-normalised) inclusive scan and reports its total; `finalize` turns it into
 * cum[i] = (offset + scan[i]) / total, where offset is the sum of the totals of the
 * lower ranks and total the global sum.  A pool made this way deposits no empty lanes. */
int qgb_getter_create_sampling_pool_partial(qgb_handle getter,
                                            const int *lane_tables, const int *n_lanes_per_qstates,
                                            const qgb_handle *qstates_list, int n_qstates,
                                            int n_lanes, int n_hidden_lanes,
                                            qgb_handle *pool, double *local_total);
int qgb_pool_finalize(qgb_handle pool, double offset, double total);
/* pool over an explicit probability vector of 2^n_lanes doubles in host memory. */
int qgb_pool_from_prob_array(int prec, const double *prob, int n_lanes,
                             const int *empty_lanes, int n_empty_lanes, qgb_handle *pool);
/* CUDA IPC: export the allocation behind a qstates (64-byte cudaIpcMemHandle_t + the
 * offset of the array inside it), map / unmap a peer's export in this process. */
int qgb_qstates_ipc_export(qgb_handle qstates, void *handle64, int64_t *offset);
/* Frees the cached blocks that were exported through CUDA IPC.  The pool never frees such a block on
 * its own (a peer that still maps it would keep the memory alive); the sharding layer calls this once
 * every rank has closed its mappings — before a sharded state vector of another size is created. */
int qgb_pool_trim_exported(void);
int qgb_ipc_open(const void *handle64, uint64_t *base_ptr);
int qgb_ipc_close(uint64_t base_ptr);
/* in-place exchange of k global lanes with the k local lanes victim_lanes[] (ascending):
 * peer_ptrs[j] (j in [0, 2^k)) is the mapped amplitude array of the rank whose selected
 * global bits equal j; my_sel is this rank's own value.  Amplitude i of this shard whose
 * victim bits are j != my_sel trades places with amplitude i' of rank j whose victim bits
 * are my_sel.  One kernel over peer memory; the caller provides the cross-rank ordering
 * (all ranks idle on this state before and after). */
int qgb_qstates_exchange_p2p(qgb_handle qstates, const uint64_t *peer_ptrs, int k,
                             const int *victim_lanes, int my_sel);
/* The push variant of the exchange: ranks write their outgoing amplitudes into each other's SPARE
 * buffers (peer_alt_ptrs[j] = spare buffer of the rank whose selector bits equal j; the caller's
 * own spare buffer at j == my_sel), then every rank makes its spare buffer the state vector.
 * Write-only NVLink traffic; needs twice the shard in memory (qgb_qstates_exchange_p2p is the
 * in-place fallback).  qgb_qstates_ipc_export_alt exports the spare buffer (allocating it). */
int qgb_qstates_ipc_export_alt(qgb_handle qstates, void *handle64, int64_t *offset);
int qgb_qstates_exchange_push(qgb_handle qstates, const uint64_t *peer_alt_ptrs, int k,
                              const int *victim_lanes, int my_sel);

/* ---- instrumentation (no reference counterpart) --------------------------- */

typedef struct qgb_stats {
    int64_t kernel_launches;      /* kernels of this library launched since reset     */
    int64_t gates_submitted;      /* apply_gate / apply_controlled_gate calls         */
    int64_t gates_executed;       /* gates after host-side merging                    */
    int64_t tile_passes;          /* fused tile passes launched                       */
    int64_t gate_amp_updates;     /* sum over submitted gates of 2^(n - n_controls)   */
    int64_t pass_bytes;           /* algorithmic bytes of all tile passes (2*2^n*B)   */
    int64_t h2d_bytes;            /* host->device bytes moved by this library         */
    int64_t d2h_bytes;            /* device->host bytes moved by this library         */
    int64_t tma_passes;           /* tile passes staged by TMA tensor-map copies      */
    int64_t shear_ops;            /* dense 2x2 gates applied as three in-place shears  */
    int64_t direct_ops;           /* dense 2x2 gates applied as a direct 2x2 product   */
    int64_t native_swaps;         /* qgb_qproc_apply_swap calls (lane relabellings)     */
    int64_t native_pauli_exps;    /* qgb_qproc_apply_pauli_expi calls                  */
    int64_t fan_ops;              /* phase fans applied (controlled phases sharing a lane, */
                                  /* merged into one op: a QFT's n(n-1)/2 become n)         */
} qgb_stats;
int qgb_stats_get(qgb_stats *out);
int qgb_stats_reset(void);
/* ---- "next" rows of the hot path ---------------------------------------
 * What the reference's front end does in Python around the native calls above, as native calls.
 *
 * f-3  multi-qubit ops without their expansion into CX chains (qgate/model/expand.py:14-90):
 *   qgb_qproc_apply_swap        Swap(a, b).  The reference applies three CX gates (three sweeps of
 *                               the state); here the two lanes exchange their index bits in the
 *                               qstates' lane map — no amplitude moves.
 *   qgb_qproc_apply_pauli_expi  exp(i theta P), P a Pauli string (paulis[i] in {0: I, 1: X, 2: Y,
 *                               3: Z} on lanes[i]), optionally controlled.  The reference applies
 *                               basis changes + 2(k-1) CX gates + ExpiZ/ExpiI; here: the basis
 *                               changes (1-qubit gates, merged into their neighbours) around ONE
 *                               parity-phase diagonal op of the fused pass.
 * f-1  batch submit (model_executor.py:80-175 + rop_executor.py:28-53 make one native call per
 *      gate): qgb_qproc_apply_gates_batch queues n_ops gates in one call.
 * f-2  Simulator.sample (simulator.py:85-119 re-runs the circuit once per shot): for circuits whose
 *      measurements are all at the end, qgb_pool_sample_sequential answers n_shots shots from ONE
 *      pool over the measured qubits — pool lane n_lanes-1 = the first measured qubit ... lane 0 = the
 *      last; randnum[s * n_lanes + k] is the draw of shot s for its k-th Measure, consumed exactly
 *      as the reference's Measure ops consume np.random (model_executor.py:117-122); obs[s] = the pool
 *      index of the outcome (no empty-lane deposit). */
int qgb_pool_sample_sequential(qgb_handle pool, int64_t *obs, int n_shots, const double *randnum);
int qgb_qproc_apply_swap(qgb_handle qproc, qgb_handle qstates, int lane_a, int lane_b);
int qgb_qproc_apply_pauli_expi(qgb_handle qproc, qgb_handle qstates, double theta, const int *lanes,
                               const int *paulis, int n, const int *ctrl_lanes, int n_ctrl);
#define QGB_GATE_MATRIX (-1)
typedef struct qgb_gate_op {
    int32_t gate_id;      /* index of the gate-matrix factory (qgb_gate_matrix), or QGB_GATE_MATRIX */
    int32_t adjoint;
    int32_t target;       /* local lane */
    int32_t reserved;
    uint64_t ctrl_mask;   /* bit l set: local lane l is a control */
    double args[3];       /* factory arguments (gate_id >= 0) */
    double mat8[8];       /* (re, im) of m00, m01, m10, m11 (gate_id == QGB_GATE_MATRIX) */
} qgb_gate_op;
int qgb_qproc_apply_gates_batch(qgb_handle qproc, qgb_handle qstates, const qgb_gate_op *ops, int64_t n_ops);

/* force the deferred gate queue of one qstates onto the device (no host sync). */
int qgb_qproc_flush(qgb_handle qproc, qgb_handle qstates);
/* tuning knobs: "max_gates_per_pass", "tile_lanes_fp32", "tile_lanes_fp64",
 * "low_lanes", "fuse" (0 = one pass per gate). */
int qgb_set_option(const char *name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* QGATE_B200_H */
