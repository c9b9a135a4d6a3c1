/*
 * planner.cpp — deferred gate queue -> pass programs (see program.h for the vocabulary).
 *
 * The reference executes one kernel per gate (CUDAQubitProcessor.cpp:270-335); it has no
 * counterpart of this file.  What must be preserved is the RESULT of applying the queued
 * gates in submission order (CPUQubitProcessor.cpp:307-362), so the planner only reorders
 * gates that commute: two gates commute when, on every lane they share, both act
 * diagonally (a control acts as the projector |1><1|, a diagonal matrix acts diagonally).
 */
#include "planner.h"

#include <algorithm>
#include <cstring>

namespace qgb {

namespace {

inline void cmul(const double *a, const double *b, double *out) {
    /* out = a * b, complex */
    double re = a[0] * b[0] - a[1] * b[1];
    double im = a[0] * b[1] + a[1] * b[0];
    out[0] = re;
    out[1] = im;
}

/* m = a * b for 2x2 complex matrices stored (re,im) row-major */
void matmul2(const double *a, const double *b, double *m) {
    double t[8];
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 2; ++c) {
            double p0[2], p1[2];
            cmul(&a[(r * 2 + 0) * 2], &b[(0 * 2 + c) * 2], p0);
            cmul(&a[(r * 2 + 1) * 2], &b[(1 * 2 + c) * 2], p1);
            t[(r * 2 + c) * 2] = p0[0] + p1[0];
            t[(r * 2 + c) * 2 + 1] = p0[1] + p1[1];
        }
    std::memcpy(m, t, sizeof(t));
}

inline int popcount64(uint64_t v) { return __builtin_popcountll(v); }

/* shared-memory bank class of a tile bit under the 128-byte XOR swizzle the kernels use
 * (byte address bits [6:4] ^= bits [9:7]); -1: the bit does not move the bank. */
inline int bank_class(int tile_bit, bool fp32) {
    if (fp32) { /* 8-byte elements, 16 lanes of a half warp must hit 16 distinct 8-byte slots */
        if (tile_bit == 0) return 0;
        if (tile_bit < 7) return 1 + (tile_bit - 1) % 3;
        return -1;
    }
    /* 16-byte elements, 8 lanes of a quarter warp must hit 8 distinct 16-byte slots */
    if (tile_bit < 6) return tile_bit % 3;
    return -1;
}

/* Thread bits of a stage: the lane bits (thread bits 0..4) first, chosen by shared-memory bank
 * class, then the rest.  `forced_high` (a mask of tile bits, none of them a register bit) is
 * placed LAST: those become the warp-index bits. */
void choose_thread_bits(const int8_t *R, int K, int T, bool fp32, int8_t *W, uint32_t forced_high = 0) {
    bool used[QGB_MAX_TILE_LANES] = {false};
    for (int j = 0; j < K; ++j) used[R[j]] = true;
    for (int b = 0; b < T; ++b)
        if (forced_high & (1u << b)) used[b] = true;
    int n_classes = fp32 ? 4 : 3;
    int n = 0;
    for (int c = 0; c < n_classes; ++c)
        for (int b = 0; b < T; ++b)
            if (!used[b] && bank_class(b, fp32) == c) {
                W[n++] = (int8_t)b;
                used[b] = true;
                break;
            }
    for (int b = 0; b < T; ++b)
        if (!used[b]) W[n++] = (int8_t)b;
    for (int b = 0; b < T; ++b)
        if (forced_high & (1u << b)) W[n++] = (int8_t)b;
    for (; n < QGB_MAX_TILE_LANES; ++n) W[n] = 0;
}

struct Picked {
    int queue_idx;
    int stage;
};

} // namespace

namespace {

inline bool mat_is_diag(const double *m) { return m[2] == 0. && m[3] == 0. && m[4] == 0. && m[5] == 0.; }

inline uint64_t touched_lanes(const Gate &p) {
    return p.ctrl_mask | (1ull << p.target) | (p.mux >= 0 ? (1ull << p.mux) : 0ull);
}

/* does p act non-diagonally on `lane`?  (a control / multiplexer lane is acted on diagonally) */
inline bool acts_nondiag_on(const Gate &p, int lane) { return p.target == lane && !gate_is_diag(p); }

const double kIdentity[8] = {1., 0., 0., 0., 0., 0., 1., 0.};

} // namespace

/* Host-side merging (matrices composed in double).  The new gate g is moved BACKWARDS over the
 * queued gates it commutes with until it meets a gate p on the same target it can fold into:
 *   plain g (no control):      commutes with everything that does not touch its target;
 *                              folds into a plain or multiplexed p (both branches get g.m on top);
 *   g with one control c:      commutes with everything that neither touches its target nor acts
 *                              non-diagonally on c; folds into a plain p (which becomes multiplexed
 *                              by c: m1 = g.m p.m) or into a p already multiplexed by c, provided
 *                              the result is dense (diagonal controlled gates stay cheap phases);
 *   anything else:             folds only into an identical-signature gate directly before it. */
bool enqueue_gate(std::vector<Gate> &queue, const Gate &g, bool merge) {
    if (merge && !queue.empty() && g.mux < 0) {
        const int n_ctrl = popcount64(g.ctrl_mask);
        if (n_ctrl <= 1) {
            const uint64_t tm = 1ull << g.target;
            const int c = n_ctrl ? __builtin_ctzll(g.ctrl_mask) : -1;
            const int lo = std::max(0, (int)queue.size() - 256);
            for (int i = (int)queue.size() - 1; i >= lo; --i) {
                Gate &p = queue[i];
                if (!(touched_lanes(p) & tm)) {
                    if (c >= 0 && acts_nondiag_on(p, c)) break; /* g does not commute with p */
                    continue;
                }
                if (p.target != g.target) break;
                if (c < 0) {
                    if (p.ctrl_mask == 0) { /* plain or multiplexed p */
                        matmul2(g.m, p.m, p.m);
                        if (p.mux >= 0) matmul2(g.m, p.m1, p.m1);
                        return true;
                    }
                } else if (p.ctrl_mask == g.ctrl_mask && p.mux < 0 && i == (int)queue.size() - 1) {
                    matmul2(g.m, p.m, p.m); /* same control, same target, adjacent */
                    return true;
                } else if (p.ctrl_mask == 0 && (p.mux < 0 || p.mux == c)) {
                    const double *low = p.m, *high = p.mux < 0 ? p.m : p.m1;
                    double m1[8];
                    matmul2(g.m, high, m1);
                    if (!(mat_is_diag(low) && mat_is_diag(m1))) {
                        std::memcpy(p.m1, m1, sizeof(m1));
                        p.mux = c;
                        return true;
                    }
                }
                break;
            }
        } else {
            Gate &p = queue.back();
            if (p.target == g.target && p.ctrl_mask == g.ctrl_mask && p.mux < 0) {
                matmul2(g.m, p.m, p.m);
                return true;
            }
        }
    }
    queue.push_back(g);
    return false;
}

template <typename real>
void plan_pass(std::vector<Gate> &queue, int n_lanes, const PlanConfig &cfg,
               PassProgram<real> &prog, PlanStats &stats) {
    const int n = n_lanes;
    const int T = std::min(std::min(cfg.T, n), QGB_MAX_TILE_LANES);
    /* at least one free tile lane, or a non-diagonal gate on a high lane could never run */
    const int L = T < n ? std::min(cfg.L, T - 1) : std::min(cfg.L, T);
    const int K = std::min(std::min(cfg.K, T), QGB_MAX_REG_BITS);
    const int max_ops = std::min(cfg.max_ops, QGB_MAX_OPS);
    const int max_stages = std::min(cfg.max_stages, QGB_MAX_STAGES);
    const uint64_t all_lanes = n >= 64 ? ~0ull : ((1ull << n) - 1);

    uint64_t S = (T >= n) ? all_lanes : ((1ull << L) - 1);
    int count = popcount64(S);
    /* TMA staging: the tile must stay describable by a 5-dimensional tensor map */
    auto fits_tensor_map = [&](uint64_t lanes) {
        if (cfg.max_groups <= 0) return true;
        int8_t gs[QGB_MAX_GROUPS], gt[QGB_MAX_GROUPS], gr[QGB_MAX_GROUPS];
        return tile_groups(lanes, n, cfg.row_lanes, gs, gt, gr) <= std::min(cfg.max_groups, QGB_MAX_GROUPS);
    };
    uint64_t blockedX = 0, blockedZ = 0;
    std::vector<Picked> picked;
    std::vector<uint64_t> stageR; /* lane masks of the register bits of each stage */
    std::vector<uint64_t> stageMux; /* multiplexer lanes of the gates of each stage */
    int first_block = -1;
    int cost = 0;

    for (int i = 0; i < (int)queue.size(); ++i) {
        if (first_block >= 0 && i - first_block > cfg.lookahead) break;
        const Gate &g = queue[i];
        const bool diag = gate_is_diag(g);
        const uint64_t tm = 1ull << g.target;
        const uint64_t xq = diag ? 0 : tm;
        const uint64_t zq = g.ctrl_mask | (diag ? tm : 0) | (g.mux >= 0 ? (1ull << g.mux) : 0ull);
        bool blocked = (xq & (blockedX | blockedZ)) || (zq & blockedX);
        const int gcost = diag ? 1 : (gate_is_antidiag(g) ? 1 : 4);
        if (!blocked && ((int)picked.size() >= max_ops || (cost + gcost > cfg.max_cost && !picked.empty())))
            blocked = true;
        bool add_lane = false;
        if (!blocked && xq && !(S & xq)) {
            if (count < T && fits_tensor_map(S | xq))
                add_lane = true;
            else
                blocked = true;
        }
        /* stage bookkeeping: 0 = stay in the open stage, 1 = open a new stage, 2 = add the
         * target to the open stage's register bits */
        int action = 0;
        if (!blocked) {
            if (stageR.empty())
                action = 1;
            else if (xq && !(stageR.back() & xq))
                action = popcount64(stageR.back()) < K ? 2 : 1;
            if (action == 1 && (int)stageR.size() >= max_stages) blocked = true;
        }
        if (blocked) {
            blockedX |= xq;
            blockedZ |= zq;
            if (first_block < 0) first_block = i;
            if ((blockedX & all_lanes) == all_lanes) break;
            continue;
        }
        if (add_lane) {
            S |= xq;
            ++count;
        }
        if (action == 1) {
            stageR.push_back(xq);
            stageMux.push_back(0);
        } else if (action == 2) {
            stageR.back() |= xq;
        }
        if (g.mux >= 0) stageMux.back() |= 1ull << g.mux;
        picked.push_back({i, (int)stageR.size() - 1});
        cost += gcost;
    }

    /* fill the tile with the highest unused lanes when fewer than T were needed (keeps the
     * kernel shape fixed; any lane works because unused tile lanes are just carried). */
    while (count < T) {
        int lane = n - 1;
        /* (a lane next to an existing run never adds a group, so there is always a candidate) */
        while (lane >= 0 && ((S & (1ull << lane)) || !fits_tensor_map(S | (1ull << lane)))) --lane;
        if (lane < 0) break;
        S |= 1ull << lane;
        ++count;
    }

    std::memset(&prog, 0, sizeof(prog));
    prog.n_lanes = n;
    prog.T = T;
    prog.L = L;
    prog.K = K;
    prog.row_lanes = cfg.row_lanes;
    prog.n_groups = 0;
    if (cfg.max_groups > 0 && count == T && ((S & ((1ull << cfg.row_lanes) - 1)) == (1ull << cfg.row_lanes) - 1)) {
        const int ng = tile_groups(S, n, cfg.row_lanes, prog.grp_start, prog.grp_t, prog.grp_r);
        prog.n_groups = ng <= QGB_MAX_GROUPS ? ng : 0;
    }
    int lane_to_tile[64];
    {
        int p = 0, r = 0;
        for (int lane = 0; lane < n; ++lane) {
            if (S & (1ull << lane)) {
                lane_to_tile[lane] = p;
                prog.tile_lane[p++] = (int8_t)lane;
            } else {
                lane_to_tile[lane] = -1;
                prog.rest_lane[r++] = (int8_t)lane;
            }
        }
    }
    /* stages */
    const int n_stages = std::max<int>(1, (int)stageR.size());
    prog.n_stages = n_stages;
    for (int s = 0; s < n_stages; ++s) {
        Stage &st = prog.stage[s];
        uint64_t Rl = s < (int)stageR.size() ? stageR[s] : 0;
        bool used[QGB_MAX_TILE_LANES] = {false};
        int k = 0;
        for (int lane = 0; lane < n; ++lane)
            if (Rl & (1ull << lane)) used[lane_to_tile[lane]] = true, ++k;
        /* pad first with multiplexer lanes of this stage's gates (the matrix choice is then uniform
         * per register pair instead of a per-thread select), then with the highest free tile
         * bits: keeps the low (bank-selecting) bits for threads */
        const uint64_t Ml = s < (int)stageMux.size() ? (stageMux[s] & S & ~Rl) : 0;
        for (int lane = n - 1; lane >= 0 && k < K; --lane)
            if ((Ml & (1ull << lane)) && !used[lane_to_tile[lane]]) used[lane_to_tile[lane]] = true, ++k;
        for (int b = T - 1; b >= 0 && k < K; --b)
            if (!used[b]) used[b] = true, ++k;
        int j = 0;
        for (int b = 0; b < T; ++b)
            if (used[b]) st.R[j++] = (int8_t)b;
        choose_thread_bits(st.R, K, T, cfg.fp32, st.W);
        st.warp_local = 0;
        for (int r = 0; r < (1 << K); ++r) {
            uint32_t roff = 0;
            for (int jj = 0; jj < K; ++jj)
                if (r & (1 << jj)) roff |= 1u << st.R[jj];
            st.sro[r] = (uint16_t)tile_swizzle(roff, cfg.fp32);
        }
        for (int jj = 0; jj < QGB_MAX_REG_BITS; ++jj)
            st.xb[jj] = jj < K ? tile_swizzle(1u << st.R[jj], cfg.fp32) * (uint32_t)(2 * sizeof(real)) : 0u;
        st.op_begin = st.op_end = 0;
    }
    /* Warp-local stage transitions: a run of consecutive stages whose register bits leave at
     * least (T - K - 5) tile bits untouched gets those bits as its warp-index bits.  Every warp
     * then owns the same 2^(K+5) tile elements through the whole run. */
    const int n_warp_bits = T - K - 5;
    if (cfg.warp_local && n_warp_bits > 0 && (int)stageR.size() >= 2) {
        const uint32_t tile_mask = (1u << T) - 1u;
        std::vector<uint32_t> rmask(n_stages, 0u);
        for (int s = 0; s < n_stages; ++s)
            for (int j = 0; j < K; ++j) rmask[s] |= 1u << prog.stage[s].R[j];
        int s = 0;
        while (s < n_stages) {
            uint32_t uni = rmask[s];
            int e = s + 1;
            while (e < n_stages && __builtin_popcount(~(uni | rmask[e]) & tile_mask) >= n_warp_bits) uni |= rmask[e++];
            if (e - s >= 2) {
                uint32_t forced = 0;
                int need = n_warp_bits;
                for (int b = T - 1; b >= 0 && need > 0; --b) /* the highest free bits: not bank-selecting */
                    if (!((uni >> b) & 1u)) forced |= 1u << b, --need;
                for (int i = s; i < e; ++i) {
                    choose_thread_bits(prog.stage[i].R, K, T, cfg.fp32, prog.stage[i].W, forced);
                    prog.stage[i].warp_local = i + 1 < e ? 1 : 0;
                }
            }
            s = e;
        }
    }

    /* ops, grouped by stage in pick order (pick order is stage-monotone) */
    int n_ops = 0;
    int cur_stage = -1;
    for (const Picked &pk : picked) {
        const Gate &g = queue[pk.queue_idx];
        while (cur_stage < pk.stage) {
            if (cur_stage >= 0) prog.stage[cur_stage].op_end = (int16_t)n_ops;
            ++cur_stage;
            prog.stage[cur_stage].op_begin = (int16_t)n_ops;
        }
        Op<real> &op = prog.op[n_ops++];
        const Stage &st = prog.stage[pk.stage];
        const uint64_t ctrl_in = g.ctrl_mask & S;
        op.ctrl_out = g.ctrl_mask & ~S;
        uint32_t ct = 0; /* controls inside the tile, tile-bit coordinates */
        for (int lane = 0; lane < n; ++lane)
            if (ctrl_in & (1ull << lane)) ct |= 1u << lane_to_tile[lane];
        const bool in_tile = (S >> g.target) & 1ull;
        auto regbit = [&](int lane) {
            int tb = lane_to_tile[lane];
            for (int j = 0; j < K; ++j)
                if (st.R[j] == tb) return j;
            return -1;
        };
        op.tsel = 0;
        op.regsel = 0;
        op.bit = 0;
        op.mux_out = 0;
        op.code = 0;
        int mux_regbit = -1;
        if (gate_is_diag(g)) {
            const bool d0_is_one = (g.m[0] == 1. && g.m[1] == 0.);
            if (d0_is_one) {
                /* phase gate: the target acts as one more control */
                op.kind = OP_DIAG;
                if (in_tile)
                    ct |= 1u << lane_to_tile[g.target];
                else
                    op.ctrl_out |= 1ull << g.target;
                op.m[0] = op.m[2] = (real)g.m[6];
                op.m[1] = op.m[3] = (real)g.m[7];
            } else {
                op.kind = in_tile ? OP_DIAG : OP_DIAG_OUT;
                op.m[0] = (real)g.m[0];
                op.m[1] = (real)g.m[1];
                op.m[2] = (real)g.m[6];
                op.m[3] = (real)g.m[7];
                if (in_tile) {
                    const int j = regbit(g.target);
                    if (j >= 0) {
                        for (int r = 0; r < (1 << K); ++r)
                            if (r & (1 << j)) op.regsel |= 1u << r;
                    } else {
                        op.tsel = 1u << lane_to_tile[g.target];
                    }
                } else {
                    op.bit = g.target;
                }
            }
        } else if (g.mux < 0 && g.m[0] == 0. && g.m[1] == 0. && g.m[6] == 0. && g.m[7] == 0. && g.m[2] == 1. &&
                   g.m[3] == 0. && g.m[4] == 1. && g.m[5] == 0.) {
            op.kind = OP_SWAP;
            op.bit = regbit(g.target);
        } else {
            op.kind = OP_GEN;
            op.bit = regbit(g.target);
            for (int e = 0; e < 8; ++e) op.m[e] = (real)g.m[e];
        }
        uint32_t mux_arm = 0;
        if (g.mux >= 0) {
            /* multiplexed 2x2: which matrix a pair gets depends on one more lane */
            for (int e = 0; e < 8; ++e) op.m1[e] = (real)g.m1[e];
            if ((S >> g.mux) & 1ull) {
                const int j = regbit(g.mux);
                if (j >= 0) {
                    mux_arm = ARM_MUX_REG;
                    mux_regbit = j;
                    for (int r = 0; r < (1 << K); ++r)
                        if (r & (1 << j)) op.regsel |= 1u << r;
                } else {
                    mux_arm = ARM_MUX_THR;
                    op.tsel = 1u << lane_to_tile[g.mux];
                }
            } else {
                mux_arm = ARM_MUX_OUT;
                op.mux_out = g.mux;
            }
        }
        if (op.kind == OP_GEN)
            op.arm = ARM_GEN(op.bit) | mux_arm;
        else if (op.kind == OP_SWAP)
            op.arm = ARM_SWAP(op.bit);
        else
            op.arm = op.regsel ? ARM_DIAG_REG : ARM_DIAG_THR;
        /* split the in-tile control predicate into its thread part and its register part */
        uint32_t rmask = 0;
        for (int j = 0; j < K; ++j) rmask |= 1u << st.R[j];
        op.cmt = ct & ~rmask;
        op.regmask = 0;
        for (int r = 0; r < (1 << K); ++r) {
            uint32_t roff = 0;
            for (int j = 0; j < K; ++j)
                if (r & (1 << j)) roff |= 1u << st.R[j];
            if ((roff & ct & rmask) == (ct & rmask)) op.regmask |= 1u << r;
        }
        const uint32_t all_regs = (1u << K) >= 32 ? 0xffffffffu : ((1u << (1 << K)) - 1u);
        if (op.kind == OP_GEN) {
            if (mux_regbit >= 0)
                op.code = OPC_GEN_REGMUX(op.bit, mux_regbit);
            else
                op.code = op.regmask == all_regs ? OPC_GEN(op.bit) : OPC_GEN_MASKED(op.bit);
        } else if (op.kind == OP_SWAP) {
            op.code = OPC_SWAP(op.bit);
        } else {
            op.code = op.regsel ? OPC_DIAG_REG : OPC_DIAG_THR;
            /* selected factor at the same place as a multiplexed matrix: m = d0.., m1 = d1.. */
            op.m1[0] = op.m[2];
            op.m1[1] = op.m[3];
        }
        {
            int sel_lane = -1;
            if (mux_arm == ARM_MUX_OUT) sel_lane = op.mux_out;
            if (op.kind == OP_DIAG_OUT) sel_lane = op.bit;
            if (op.ctrl_out != 0 || sel_lane >= 0) {
                auto &ref = prog.out[prog.n_out++];
                ref.ctrl_mask = op.ctrl_out;
                ref.op = (int16_t)(n_ops - 1);
                ref.sel_lane = (int16_t)sel_lane;
                ref.pad_ = 0;
            }
        }
    }
    if (cur_stage >= 0) prog.stage[cur_stage].op_end = (int16_t)n_ops;
    for (int s = cur_stage + 1; s < n_stages; ++s)
        prog.stage[s].op_begin = prog.stage[s].op_end = (int16_t)n_ops;
    prog.n_ops = n_ops;

    /* drop the executed gates, keep the rest in order */
    {
        std::vector<char> done(queue.size(), 0);
        for (const Picked &pk : picked) done[pk.queue_idx] = 1;
        size_t w = 0;
        for (size_t i = 0; i < queue.size(); ++i)
            if (!done[i]) {
                if (w != i) queue[w] = queue[i];
                ++w;
            }
        queue.resize(w);
    }
    stats.gates_in_pass = (int)picked.size();
    stats.ops_in_pass = n_ops;
    stats.stages_in_pass = n_stages;
}

template void plan_pass<float>(std::vector<Gate> &, int, const PlanConfig &, PassProgram<float> &,
                               PlanStats &);
template void plan_pass<double>(std::vector<Gate> &, int, const PlanConfig &, PassProgram<double> &,
                                PlanStats &);

} // namespace qgb
