/*
 * kernels.h — launchers of the sm_100a kernels (definitions in kernels_tma.cu, kernels_ops.cu, dist.cu).
 * Every launcher enqueues on `stream` and returns the cudaError_t of the launch; none of
 * them synchronises.  `prec` is QGB_PREC_FP64 (1) or QGB_PREC_FP32 (2); amplitudes are
 * interleaved (re, im) in that precision.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "program.h"

namespace qgb {

/* maximum number of source qstates of one join (the product has at least that many lanes) */
#define QGB_MAX_QSTATES 40

struct SortedBits {
    int32_t n;
    int8_t pos[QGB_MAX_LANES]; /* ascending bit positions to be skipped by the pair index */
};

/* external index -> local index of one qstates: local bit l comes from external bit ext[l] */
struct LaneTable {
    const void *amp;
    int32_t n_lanes;
    int8_t ext[QGB_MAX_LANES];
};

/* any number of qstates (dynamic qubit grouping keeps every unentangled qreg as its own 1-lane
 * qstates): the tables live in device memory */
struct GatherParams {
    int32_t n_qstates;
    const LaneTable *qs; /* device array of n_qstates tables */
};

/* ---- gates ----------------------------------------------------------------------- */
/* the fused gate pass, TMA tensor-map staging (kernels_tma.cu); needs prog.n_groups >= 1;
 * n_buf = 2 or 3 tile buffers per CTA; min_ctas = resident CTAs per SM the register budget is
 * set for (3: 80 registers per thread, 2: up to 128); zero_input = 1: the array has not been
 * written since the state was reset to |0...0> — the pass reads nothing and starts every tile
 * from that state (half the traffic of the first pass, and no initialisation sweep before it) */
template <typename real>
cudaError_t launch_tma_pass(const PassProgram<real> &prog, void *amp, int n_buf, int min_ctas, cudaStream_t stream,
                            int zero_input = 0);
/* shared memory of one CTA: n_buf tiles, the matrices of n_ops ops, n_stages thread tables */
size_t tma_pass_smem_bytes(int prec, int T, int K, int n_stages, int n_buf, int n_ops, int n_fans);
cudaError_t tma_pass_configure(int max_smem_optin, int sm_count);
void tma_pass_release(); /* frees the scratch of the phase-fan tile tables */
void tma_pass_set_warp_specialised(int on); /* 1: a producer warp owns the TMA traffic (default 0) */
void tma_pass_set_l2_hint(int mode);    /* 1: loads, 2: stores, 3: both with an L2 evict_first policy */
void tma_pass_set_debug_mode(int mode); /* measurement only: 1 = stages without ops, 2 = no stages     */
void tma_pass_phase_report(); /* no-op unless built with -DQGB_PHASE_TIMING */

/* one gate per launch: pairs whose ctrl_mask bits are 1 and zero_mask bits are 0.  exact: the
 * reference CPU runtime's arithmetic, operation by operation (bit-identical amplitudes) */
cudaError_t launch_simple_gate(int prec, void *amp, int n_lanes, const double *mat8, int target,
                               uint64_t ctrl_mask, uint64_t zero_mask, bool exact, cudaStream_t stream);

/* parity diagonal, one launch: a *= d0 / d1 by the parity of the index bits in `parity` (where the
 * ctrl_mask bits are set) */
cudaError_t launch_simple_parity(int prec, void *amp, int n_lanes, const double *d0, const double *d1, uint64_t parity,
                                 uint64_t ctrl_mask, cudaStream_t stream);

/* ---- state-vector maintenance ------------------------------------------------------ */
cudaError_t launch_set_basis_state(int prec, void *amp, uint64_t n_amps, uint64_t one_at,
                                   cudaStream_t stream);
/* partial sums -> *d_result (double) = sum_{bit lane == 0} |a|^2; d_partials holds >= 2048 doubles */
cudaError_t launch_prob0(int prec, const void *amp, int n_lanes, int lane, double *d_partials,
                         double *d_result, cudaStream_t stream);
/* *d_result = sum of |a|^2 over the whole array */
cudaError_t launch_norm(int prec, const void *amp, int n_lanes, double *d_partials, double *d_result,
                        cudaStream_t stream);
cudaError_t launch_decohere(int prec, void *amp, int n_lanes, int lane, int value, double norm,
                            cudaStream_t stream);
cudaError_t launch_decohere_separate(int prec, void *dst, const void *src, int n_src_lanes, int lane,
                                     int value, double norm, cudaStream_t stream);
cudaError_t launch_apply_reset(int prec, void *amp, int n_lanes, int lane, cudaStream_t stream);

struct JoinParams {
    int32_t n_src;
    const void *src[QGB_MAX_QSTATES]; /* in list order; the LAST one holds the lowest lanes */
    int32_t shift[QGB_MAX_QSTATES];
    int32_t n_lanes[QGB_MAX_QSTATES];
};
/* dst[d] = prod_k src_k[(i >> shift_k) & mask_k] for i = index_offset + d < 2^n_product, 0 above
 * (index_offset != 0: dst is one shard of the product) */
cudaError_t launch_join(int prec, void *dst, int n_dst_lanes, int n_product_lanes,
                        uint64_t index_offset, const JoinParams &jp, cudaStream_t stream);

/* ---- readout ------------------------------------------------------------------------ */
/* out[j] (complex<real> for mathop 0, real for mathop 1), j in [0, count):
 * ext = start + step * (first + j); 0 if ext & empty_mask else prod_qs op(amp_qs[perm(ext)]) */
cudaError_t launch_get_states(int prec, void *d_out, int mathop, const GatherParams &gp,
                              uint64_t empty_mask, int64_t first, int64_t count, int64_t start,
                              int64_t step, cudaStream_t stream);
/* marginal probabilities: d_out[d] (double), d in [first, first+count):
 * sum_{h < 2^n_hidden} prod_qs |amp_qs[perm((d << n_hidden) | h)]|^2 (product in `real`, sum in double) */
cudaError_t launch_prob_array(int prec, double *d_out, const GatherParams &gp, int n_hidden,
                              int64_t first, int64_t count, bool compat, cudaStream_t stream);
/* d_out[j] = sum of 2^log2_group consecutive d_in values (sequential; in double, or in float when
 * fp32_sums), j < count */
cudaError_t launch_reduce_groups(double *d_out, const double *d_in, int log2_group, int64_t count,
                                 bool fp32_sums, cudaStream_t stream);
/* reference-compatible pool scan (CPUSamplingPool.cpp:13-47): sequential running sums in the state
 * precision over n_spans spans of `span` entries; then (p + offset of the span) * norm */
cudaError_t launch_compat_scan(int prec, double *d_prob, int64_t n, int64_t span, int n_spans,
                               double *d_span_total, cudaStream_t stream);
cudaError_t launch_compat_apply(int prec, double *d_prob, int64_t n, int64_t span, const double *d_span_offset,
                                double norm, cudaStream_t stream);
/* double -> real conversion of a device array (for handing the prob array to the host) */
cudaError_t launch_cast_from_double(int prec, void *d_out, const double *d_in, int64_t count,
                                    cudaStream_t stream);

/* ---- sampling pool ------------------------------------------------------------------- */
/* inclusive scan in double with a fixed tiling (deterministic) of either d_prob[0..n) in place
 * (amp == nullptr) or of |amp[i]|^2 computed on the fly (amp = 2^n_lanes amplitudes of precision
 * `prec`; the probability vector is then never materialised).  d_block_sums holds ceil(n / 4096)
 * doubles; *d_total receives the grand total. */
cudaError_t launch_scan_phase1(int prec, const void *amp, const double *d_prob, int64_t n, double *d_block_sums,
                               cudaStream_t stream);
cudaError_t launch_scan_phase2(double *d_block_sums, int64_t n_blocks, double *d_total, cudaStream_t stream);
/* cum[i] = (global_offset + offset_block + inclusive_scan_in_block) * norm */
cudaError_t launch_scan_phase3(int prec, const void *amp, double *d_cum, int64_t n, const double *d_block_sums,
                               double global_offset, double norm, cudaStream_t stream);
/* obs[i] = deposit(upper_bound(cum, r_i), perm); r is cast to float first when prec is FP32
 * (CPUSamplingPool.cpp:74) */
cudaError_t launch_sample(int prec, const double *d_cum, int n_lanes, const double *d_rand,
                          int64_t *d_obs, int n_samples, SortedBits empty_lanes, cudaStream_t stream);

/* obs[s] = outcome of n_bits sequential measurements (most significant pool lane first) decided by
 * the draws rnd[s * n_bits + k] against conditional probabilities taken from the cumulative array */
cudaError_t launch_sample_sequential(const double *d_cum, int n_bits, const double *d_rand, int64_t *d_obs,
                                     int n_shots, cudaStream_t stream);

/* ---- sharded states: lane exchange over peer memory (dist.cu) --------------------------- */
#define QGB_MAX_EXCHANGE_LANES 4

struct ExchangeParams {
    int32_t k;                 /* lanes exchanged at once                                        */
    int32_t my_sel;            /* this rank's value of the k selected global bits               */
    int32_t n_unit_bits;       /* log2 of the shard size in 16-byte units                        */
    int32_t split;             /* unit bit that decides which rank of a pair moves an element   */
    int32_t victim[QGB_MAX_EXCHANGE_LANES]; /* ascending unit-bit positions of the local lanes */
    void *peer[1 << QGB_MAX_EXCHANGE_LANES]; /* peer[j]: shard of the rank whose bits equal j   */
};
cudaError_t launch_exchange_p2p(void *local, const ExchangeParams &ep, int sm_count, cudaStream_t stream);
/* push variant: ep.peer[j] = SPARE buffer of the rank whose selector bits equal j (this rank's own
 * spare buffer for j == my_sel); `local` is only read */
cudaError_t launch_exchange_push(const void *local, const ExchangeParams &ep, int sm_count, cudaStream_t stream);

} // namespace qgb
