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
#include <cmath>
#include <complex>
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

const double kIdentity[8] = {1., 0., 0., 0., 0., 0., 1., 0.};

inline uint64_t touched_lanes(const Gate &p) {
    return p.ctrl_mask | (1ull << p.target) | (p.mux >= 0 ? (1ull << p.mux) : 0ull) | p.parity | gate_fan_lanes(p);
}

/* a controlled phase CP(a, b; e^{i phi}): diag(1, e^{i phi}) on the target under ONE control, which is
 * symmetric in its two lanes */
inline bool is_controlled_phase(const Gate &p) {
    return p.fan.empty() && p.mux < 0 && p.parity == 0 && popcount64(p.ctrl_mask) == 1 && mat_is_diag(p.m) &&
           p.m[0] == 1. && p.m[1] == 0.;
}

const int kMaxFanTerms = 48;

void fan_add_term(Gate &fan, int lane, double re, double im) {
    for (Gate::FanTerm &t : fan.fan)
        if (t.lane == lane) {
            const double r = t.re * re - t.im * im, i = t.re * im + t.im * re;
            t.re = r, t.im = i;
            return;
        }
    fan.fan.push_back({lane, re, im});
}

/* Phase fans: the controlled phase g = CP(a, b) moves backwards over the queued gates it commutes
 * with (everything that acts diagonally on a and on b) until it meets a fan whose hub is a or b (it
 * becomes one more term), or another controlled phase that shares a lane with it (the two become a
 * fan with that lane as the hub).  The n - 1 - i controlled phases a QFT applies onto qubit i end up
 * as ONE gate, whatever order the circuit lists them in. */
bool merge_into_fan(std::vector<Gate> &queue, const Gate &g) {
    const int a = __builtin_ctzll(g.ctrl_mask), b = g.target;
    const uint64_t ab = (1ull << a) | (1ull << b);
    const int lo = std::max(0, (int)queue.size() - 256);
    for (int i = (int)queue.size() - 1; i >= lo; --i) {
        Gate &p = queue[i];
        if (!p.fan.empty()) {
            if ((p.target == a || p.target == b) && p.ctrl_mask == 0 && (int)p.fan.size() < kMaxFanTerms) {
                fan_add_term(p, p.target == a ? b : a, g.m[6], g.m[7]);
                return true;
            }
            continue; /* diagonal: commutes */
        }
        if (is_controlled_phase(p)) {
            const int pa = __builtin_ctzll(p.ctrl_mask), pb = p.target;
            const uint64_t pab = (1ull << pa) | (1ull << pb);
            if (pab == ab) { /* the same pair: the phases multiply */
                const double r = p.m[6] * g.m[6] - p.m[7] * g.m[7], im = p.m[6] * g.m[7] + p.m[7] * g.m[6];
                p.m[6] = r, p.m[7] = im;
                return true;
            }
            if (pab & ab) {
                const int hub = __builtin_ctzll(pab & ab);
                Gate f;
                std::memcpy(f.m, kIdentity, sizeof(f.m));
                f.target = hub;
                f.ctrl_mask = 0;
                f.fan.push_back({pa == hub ? pb : pa, p.m[6], p.m[7]});
                f.fan.push_back({a == hub ? b : a, g.m[6], g.m[7]});
                p = f;
                return true;
            }
            continue;
        }
        if ((p.target == a || p.target == b) && !gate_is_diag(p)) return false; /* does not commute */
    }
    return false;
}

/* does p act non-diagonally on `lane`?  (a control / multiplexer lane is acted on diagonally) */
inline bool acts_nondiag_on(const Gate &p, int lane) { return p.target == lane && !gate_is_diag(p); }

typedef std::complex<double> cd;

/* M (row-major (re, im) x 4) = phi * P^s * N with det N = 1 and
 * N = [[1,x],[0,1]] [[1,0],[g,1]] [[1,y],[0,1]]: g = N10, x = (N00 - 1) / g, y = (N11 - 1) / g.
 * P exchanges the two outputs (rows).  Of the two square roots of the determinant the one with
 * Re(N00 + N11) >= 0 is taken (smaller x, y).  `size` = the largest coefficient magnitude: the
 * error growth of the three shears.  False when M is singular or g vanishes. */
struct ShearCoef {
    double y[2], g[2], x[2], phi[2];
    double size;
};

bool shear_factor(const double *m, int s, ShearCoef &out) {
    cd a(m[0], m[1]), b(m[2], m[3]), c(m[4], m[5]), d(m[6], m[7]);
    if (s) std::swap(a, c), std::swap(b, d);
    const cd det = a * d - b * c;
    const double ad = std::abs(det);
    if (!(ad > 1e-280) || !std::isfinite(ad)) return false;
    cd phi = std::sqrt(det);
    if (((a + d) / phi).real() < 0.) phi = -phi;
    const cd n00 = a / phi, n11 = d / phi, g = c / phi;
    const double ag = std::abs(g);
    if (!(ag > 1e-280)) return false;
    const cd x = (n00 - 1.) / g, y = (n11 - 1.) / g;
    out.y[0] = y.real(), out.y[1] = y.imag();
    out.g[0] = g.real(), out.g[1] = g.imag();
    out.x[0] = x.real(), out.x[1] = x.imag();
    out.phi[0] = phi.real(), out.phi[1] = phi.imag();
    out.size = std::max(std::abs(x), std::max(std::abs(y), ag));
    return std::isfinite(out.size);
}

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
bool enqueue_gate(std::vector<Gate> &queue, const Gate &g, bool merge, bool fans) {
    if (merge && fans && !queue.empty() && is_controlled_phase(g) && merge_into_fan(queue, g)) return true;
    if (merge && !queue.empty() && g.mux < 0 && g.parity == 0 && g.fan.empty()) {
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
                if (!p.fan.empty()) {
                    /* a fan acts diagonally on all its lanes: a diagonal g commutes with it */
                    if (mat_is_diag(g.m)) continue;
                    break;
                }
                if (p.target != g.target || p.parity) break;
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
            if (p.target == g.target && p.ctrl_mask == g.ctrl_mask && p.mux < 0 && p.parity == 0 && p.fan.empty()) {
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
    int cost = 0, slots = 0, n_fans = 0, n_fan_terms = 0;

    for (int i = 0; i < (int)queue.size(); ++i) {
        if (first_block >= 0 && i - first_block > cfg.lookahead) break;
        const Gate &g = queue[i];
        const bool diag = gate_is_diag(g);
        const uint64_t tm = 1ull << g.target;
        const uint64_t xq = diag ? 0 : tm;
        const uint64_t zq = g.ctrl_mask | (diag ? gate_diag_lanes(g) : 0) | (g.mux >= 0 ? (1ull << g.mux) : 0ull);
        bool blocked = (xq & (blockedX | blockedZ)) || (zq & blockedX);
        const bool is_fan = !g.fan.empty();
        /* (a fan multiplies half of the amplitudes once and builds one factor per thread and tile) */
        const int gcost = is_fan ? cfg.fan_cost : (diag ? 1 : (gate_is_antidiag(g) ? 1 : (cfg.shear ? 3 : 4)));
        /* op slots: a sheared gate with controls or a multiplexer may need its phase applied as one
         * more (diagonal) op of this pass */
        const int gslots = (cfg.shear && !diag && (g.mux >= 0 || g.ctrl_mask != 0)) ? 2 : 1;
        if (!blocked && !picked.empty() && (slots + gslots > max_ops || cost + gcost > cfg.max_cost))
            blocked = true;
        if (!blocked && is_fan && (n_fans + 1 > QGB_MAX_FANS || n_fan_terms + (int)g.fan.size() > QGB_MAX_FAN_TERMS))
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
        slots += gslots;
        if (is_fan) ++n_fans, n_fan_terms += (int)g.fan.size();
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

    /* ops, grouped by stage in pick order (pick order is stage-monotone).
     *
     * Dense gates become OP_SHEAR where the factorisation M = phi * P^s * N is well conditioned
     * (shear_factor).  What it leaves behind:
     *   P^s   a relabelling of the stage's registers (`flip`): every later op of the stage is
     *         expressed in relabelled registers here, the kernel stores through Stage::store_xor;
     *   phi   a diagonal "residual" (a scalar, or diag(phi0, phi1) on the multiplexer lane, or a
     *         controlled phase) that commutes with the gate itself: it is folded into the next
     *         gate that acts non-diagonally on its lane (picked later in this pass, or still
     *         queued), or applied as a diagonal op when there is none. */
    std::vector<char> is_picked(queue.size(), 0);
    for (const Picked &pk : picked) is_picked[pk.queue_idx] = 1;
    std::vector<Gate> deferred; /* residuals for the front of the remaining queue */
    const int op_limit = max_ops;
    int n_ops = 0;
    int cur_stage = -1;
    uint32_t flip = 0; /* register relabelling of the open stage */
    stats.shear_ops = stats.direct_ops = stats.residual_ops = 0;
    prog.n_direct = 0;

    auto close_stage = [&](int sidx) {
        Stage &st = prog.stage[sidx];
        st.op_end = (int16_t)n_ops;
        st.store_flip = flip;
        st.store_xor = 0;
        for (int j = 0; j < K; ++j)
            if (flip & (1u << j)) st.store_xor ^= st.xb[j];
    };

    /* one op for gate g in stage sidx; dense gates may leave a residual in `res` (res_kind: 0 none,
     * 1 diagonal gate `res`) */
    auto lower = [&](const Gate &g, int sidx, Gate &res) -> int {
        Op<real> &op = prog.op[n_ops++];
        const Stage &st = prog.stage[sidx];
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
        /* tile offset of the element register r holds NOW (relabelled by the stage's shears) */
        auto roff_of = [&](int r) {
            const uint32_t logical = (uint32_t)r ^ flip;
            uint32_t roff = 0;
            for (int j = 0; j < K; ++j)
                if (logical & (1u << j)) roff |= 1u << st.R[j];
            return roff;
        };
        op.tsel = 0;
        op.regsel = 0;
        op.bit = 0;
        op.mux_out = 0;
        op.code = 0;
        int res_kind = 0;
        int mux_regbit = -1;
        bool shear = false;
        uint32_t mux_arm = 0;
        const bool is_x = g.mux < 0 && g.m[0] == 0. && g.m[1] == 0. && g.m[6] == 0. && g.m[7] == 0. && g.m[2] == 1. &&
                          g.m[3] == 0. && g.m[4] == 1. && g.m[5] == 0.;
        uint64_t sel_out = 0; /* lanes outside the tile whose parity picks the second matrix / factor */
        int fan_idx = -1;
        if (!g.fan.empty()) {
            /* phase fan: the hub is one more control; the terms are sorted by where their lane lives in
             * this stage (thread bit / outside the tile / register bit) */
            op.kind = OP_FAN;
            if (in_tile)
                ct |= 1u << lane_to_tile[g.target];
            else
                op.ctrl_out |= 1ull << g.target;
            op.m[0] = op.m[2] = (real)1;
            fan_idx = prog.n_fans++;
            op.bit = fan_idx;
            auto &fn = prog.fan[fan_idx];
            fn.first = (int16_t)prog.n_fan_terms;
            fn.n_thr = fn.n_out = fn.n_reg = 0;
            fn.base[0] = 1., fn.base[1] = 0.;
            for (int j = 0; j < QGB_MAX_REG_BITS; ++j) fn.reg[j][0] = 1., fn.reg[j][1] = 0.;
            for (int cls = 0; cls < 3; ++cls) /* 0: thread bits, 1: outside, 2: register bits */
                for (const Gate::FanTerm &t : g.fan) {
                    const bool inside = (S >> t.lane) & 1ull;
                    const int j = inside ? regbit(t.lane) : -1;
                    const int c = !inside ? 1 : (j >= 0 ? 2 : 0);
                    if (c != cls) continue;
                    if (cls == 2) {
                        /* register r holds the element of register index r ^ flip: on a relabelled bit
                         * the term belongs to the registers whose bit j is CLEAR: t everywhere, 1 / t
                         * where the bit is set */
                        fn.n_reg = (int16_t)(fn.n_reg | (1 << j));
                        if (flip & (1u << j)) {
                            const cd tt(t.re, t.im), b = cd(fn.base[0], fn.base[1]) * tt, inv = 1. / tt;
                            fn.base[0] = b.real(), fn.base[1] = b.imag();
                            fn.reg[j][0] = inv.real(), fn.reg[j][1] = inv.imag();
                        } else {
                            fn.reg[j][0] = t.re, fn.reg[j][1] = t.im;
                        }
                        continue;
                    }
                    auto &ft = prog.fan_term[prog.n_fan_terms++];
                    ft.re = t.re, ft.im = t.im, ft.pad_ = 0;
                    if (cls == 0) {
                        int pos = -1;
                        for (int i = 0; i < T - K; ++i)
                            if (st.W[i] == lane_to_tile[t.lane]) pos = i;
                        ft.bit = pos;
                        ++fn.n_thr;
                    } else {
                        ft.bit = t.lane;
                        ++fn.n_out;
                    }
                }
        } else if (gate_is_diag(g)) {
            const uint64_t lanes = gate_diag_lanes(g);
            const bool d0_is_one = g.parity == 0 && (g.m[0] == 1. && g.m[1] == 0.);
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
                /* d0 where an even number of the gate's lanes is 1, d1 where an odd number is: the
                 * parity splits into a register part (regsel), a thread part (tsel) and a part
                 * that is the same for the whole tile (sel_out) */
                op.kind = (lanes & S) ? OP_DIAG : OP_DIAG_OUT;
                op.m[0] = (real)g.m[0];
                op.m[1] = (real)g.m[1];
                op.m[2] = (real)g.m[6];
                op.m[3] = (real)g.m[7];
                uint32_t reg_bits = 0;
                for (int lane = 0; lane < n; ++lane) {
                    if (!((lanes >> lane) & 1ull)) continue;
                    if (!((S >> lane) & 1ull)) {
                        sel_out |= 1ull << lane;
                        continue;
                    }
                    const int j = regbit(lane);
                    if (j >= 0)
                        reg_bits |= 1u << j;
                    else
                        op.tsel |= 1u << lane_to_tile[lane];
                }
                for (int r = 0; r < (1 << K); ++r)
                    if (__builtin_popcount(((uint32_t)r ^ flip) & reg_bits) & 1) op.regsel |= 1u << r;
                if (op.kind == OP_DIAG_OUT) op.bit = g.target;
            }
        } else if (is_x) {
            op.kind = OP_SWAP;
            op.bit = regbit(g.target);
        } else {
            op.kind = OP_GEN;
            const int j = regbit(g.target);
            op.bit = j;
            /* the pair (a[r0], a[r0 | bit]) holds (q1, q0) when an earlier shear of this stage
             * exchanged the outputs of register bit j: the gate then acts as P M P on it */
            double m0[8], m1[8];
            auto oriented = [&](const double *src, double *dst) {
                if (flip & (1u << j)) {
                    const int perm[4] = {3, 2, 1, 0};
                    for (int e = 0; e < 4; ++e) dst[2 * e] = src[2 * perm[e]], dst[2 * e + 1] = src[2 * perm[e] + 1];
                } else {
                    std::memcpy(dst, src, 8 * sizeof(double));
                }
            };
            oriented(g.m, m0);
            if (g.mux >= 0) oriented(g.m1, m1);
            if (g.mux >= 0) {
                /* multiplexed 2x2: which matrix a pair gets depends on one more lane */
                if ((S >> g.mux) & 1ull) {
                    const int j2 = regbit(g.mux);
                    if (j2 >= 0) {
                        mux_arm = ARM_MUX_REG;
                        mux_regbit = j2;
                        for (int r = 0; r < (1 << K); ++r)
                            if (r & (1 << j2)) op.regsel |= 1u << r;
                        /* registers with bit j2 SET take the second matrix: under a relabelled j2
                         * those hold the elements whose multiplexer bit is 0 */
                        if (flip & (1u << j2)) {
                            double t[8];
                            std::memcpy(t, m0, sizeof(t));
                            std::memcpy(m0, m1, sizeof(t));
                            std::memcpy(m1, t, sizeof(t));
                        }
                    } else {
                        mux_arm = ARM_MUX_THR;
                        op.tsel = 1u << lane_to_tile[g.mux];
                    }
                } else {
                    mux_arm = ARM_MUX_OUT;
                    op.mux_out = g.mux;
                }
            }
            /* three shears?  Exchanged outputs (s = 1) only where EVERY pair of every thread gets
             * the gate: no controls, and both branches of a multiplexed gate agree on s */
            ShearCoef c0[2], c1[2];
            int s_pick = -1;
            if (cfg.shear) {
                double best = cfg.shear_max_coef;
                const int s_hi = g.ctrl_mask == 0 ? 1 : 0;
                for (int sx = 0; sx <= s_hi; ++sx) {
                    if (!shear_factor(m0, sx, c0[sx])) continue;
                    double worst = c0[sx].size;
                    if (g.mux >= 0) {
                        if (!shear_factor(m1, sx, c1[sx])) continue;
                        worst = std::max(worst, c1[sx].size);
                    }
                    if (worst <= best) best = worst, s_pick = sx;
                }
            }
            if (s_pick >= 0) {
                shear = true;
                op.kind = OP_SHEAR;
                const ShearCoef &a0 = c0[s_pick];
                const double k0[8] = {a0.y[0], a0.y[1], a0.g[0], a0.g[1], a0.x[0], a0.x[1], 0., 0.};
                for (int e = 0; e < 8; ++e) op.m[e] = (real)k0[e];
                double phi0[2] = {a0.phi[0], a0.phi[1]}, phi1[2] = {a0.phi[0], a0.phi[1]};
                if (g.mux >= 0) {
                    const ShearCoef &a1 = c1[s_pick];
                    const double k1[8] = {a1.y[0], a1.y[1], a1.g[0], a1.g[1], a1.x[0], a1.x[1], 0., 0.};
                    for (int e = 0; e < 8; ++e) op.m1[e] = (real)k1[e];
                    phi1[0] = a1.phi[0], phi1[1] = a1.phi[1];
                    if (mux_regbit >= 0 && (flip & (1u << mux_regbit))) std::swap(phi0[0], phi1[0]), std::swap(phi0[1], phi1[1]);
                }
                if (s_pick) flip ^= 1u << j;
                /* the residual: phi on every amplitude the gate touched */
                const bool trivial = phi0[0] == 1. && phi0[1] == 0. && phi1[0] == 1. && phi1[1] == 0.;
                if (!trivial) {
                    res = Gate();
                    for (int e = 0; e < 8; ++e) res.m[e] = 0.;
                    if (g.mux >= 0) { /* diag(phi0, phi1) on the multiplexer lane */
                        res.target = g.mux;
                        res.ctrl_mask = 0;
                        res.m[0] = phi0[0], res.m[1] = phi0[1], res.m[6] = phi1[0], res.m[7] = phi1[1];
                    } else if (g.ctrl_mask == 0) { /* a scalar: diag(phi, phi), any lane */
                        res.target = g.target;
                        res.ctrl_mask = 0;
                        res.m[0] = res.m[6] = phi0[0], res.m[1] = res.m[7] = phi0[1];
                    } else { /* controlled phase: one control lane becomes the target of diag(1, phi) */
                        const int lane = __builtin_ctzll(g.ctrl_mask);
                        res.target = lane;
                        res.ctrl_mask = g.ctrl_mask & ~(1ull << lane);
                        res.m[0] = 1., res.m[6] = phi0[0], res.m[7] = phi0[1];
                    }
                    res_kind = 1;
                }
            } else {
                for (int e = 0; e < 8; ++e) op.m[e] = (real)m0[e];
                if (g.mux >= 0)
                    for (int e = 0; e < 8; ++e) op.m1[e] = (real)m1[e];
            }
        }
        if (op.kind == OP_GEN || op.kind == OP_SHEAR)
            op.arm = ARM_GEN(op.bit) | mux_arm;
        else if (op.kind == OP_FAN)
            op.arm = ARM_DIAG_THR;
        else if (op.kind == OP_SWAP)
            op.arm = ARM_SWAP(op.bit);
        else
            op.arm = op.regsel ? ARM_DIAG_REG : ARM_DIAG_THR;
        /* split the in-tile control predicate into its thread part and its register part */
        uint32_t rmask = 0;
        for (int j = 0; j < K; ++j) rmask |= 1u << st.R[j];
        op.cmt = ct & ~rmask;
        op.regmask = 0;
        for (int r = 0; r < (1 << K); ++r)
            if ((roff_of(r) & ct & rmask) == (ct & rmask)) op.regmask |= 1u << r;
        const uint32_t all_regs = (1u << K) >= 32 ? 0xffffffffu : ((1u << (1 << K)) - 1u);
        if (op.kind == OP_GEN || op.kind == OP_SHEAR) {
            if (mux_regbit >= 0)
                op.code = shear ? OPC_SHEAR_REGMUX(op.bit, mux_regbit) : OPC_GEN_REGMUX(op.bit, mux_regbit);
            else if (op.regmask == all_regs)
                op.code = shear ? OPC_SHEAR(op.bit) : OPC_GEN(op.bit);
            else
                op.code = shear ? OPC_SHEAR_MASKED(op.bit) : OPC_GEN_MASKED(op.bit);
            if (shear)
                ++stats.shear_ops;
            else
                ++stats.direct_ops, ++prog.n_direct;
        } else if (op.kind == OP_SWAP) {
            op.code = OPC_SWAP(op.bit);
        } else if (op.kind == OP_FAN) {
            /* the hub predicate over the registers: all of them (hub on a thread bit or outside the
             * tile), or the half whose bit j is 1 (0 after a relabelling of that bit) */
            op.code = op.regmask == all_regs ? OPC_FAN_ALL : OPC_FAN;
            for (int j = 0; j < K; ++j)
                for (int pol = 0; pol < 2; ++pol) {
                    uint32_t pattern = 0;
                    for (int r = 0; r < (1 << K); ++r)
                        if (((r >> j) & 1) == pol) pattern |= 1u << r;
                    if (op.regmask == pattern) op.code = OPC_FAN_REG(j, pol);
                }
        } else {
            /* regsel != 0 or a relabelled register part: the register-diagonal body (it also serves
             * regsel == 0 after a relabelling turned every register to the same side) */
            op.code = op.regsel ? OPC_DIAG_REG : OPC_DIAG_THR;
            /* the block a thread / tile with odd parity takes: the factors exchanged, so that the
             * selected block always starts with the factor of the registers outside regsel */
            op.m1[0] = op.m[2];
            op.m1[1] = op.m[3];
            op.m1[2] = op.m[0];
            op.m1[3] = op.m[1];
        }
        {
            if (mux_arm == ARM_MUX_OUT) sel_out = 1ull << op.mux_out;
            if (op.ctrl_out != 0 || sel_out != 0) {
                auto &ref = prog.out[prog.n_out++];
                ref.ctrl_mask = op.ctrl_out;
                ref.sel_mask = sel_out;
                ref.op = (int16_t)(n_ops - 1);
                ref.pad_[0] = ref.pad_[1] = ref.pad_[2] = 0;
            }
        }
        return res_kind;
    };

    /* D (diagonal, fresh from gate number pi of `picked`) -> into the next gate that acts
     * non-diagonally on its lane.  Returns 0: folded away, 1: apply it now (the next such gate is
     * controlled and runs in this pass), 2: no gate in this pass needs it applied first */
    auto fold_forward = [&](const Gate &D, size_t pi) -> int {
        const bool scalar = D.ctrl_mask == 0 && D.m[0] == D.m[6] && D.m[1] == D.m[7];
        auto absorbs = [&](Gate &G) {
            /* G <- G D: D acts first */
            const double d[2][2] = {{D.m[0], D.m[1]}, {D.m[6], D.m[7]}};
            auto scale_cols = [&](double *m) {
                for (int row = 0; row < 2; ++row)
                    for (int col = 0; col < 2; ++col) {
                        double *e = &m[(row * 2 + col) * 2];
                        const double re = e[0] * d[col][0] - e[1] * d[col][1];
                        const double im = e[0] * d[col][1] + e[1] * d[col][0];
                        e[0] = re, e[1] = im;
                    }
            };
            scale_cols(G.m);
            if (G.mux >= 0) scale_cols(G.m1);
        };
        auto visit = [&](Gate &G, bool in_pass) -> int { /* -1: keep looking */
            if (gate_is_diag(G)) return -1;
            if (scalar) {
                if (G.ctrl_mask != 0) return -1; /* a scalar commutes with everything: look further */
                absorbs(G);
                return 0;
            }
            /* D is diagonal on its target and on its controls */
            if (!(((D.ctrl_mask | (1ull << D.target)) >> G.target) & 1ull)) return -1;
            if (G.ctrl_mask == 0 && D.ctrl_mask == 0) {
                absorbs(G);
                return 0;
            }
            return in_pass ? 1 : 2;
        };
        for (size_t k = pi + 1; k < picked.size(); ++k) {
            const int r = visit(queue[picked[k].queue_idx], true);
            if (r >= 0) return r;
        }
        for (size_t k = 0; k < queue.size(); ++k) {
            if (is_picked[k]) continue;
            const int r = visit(queue[k], false);
            if (r >= 0) return r;
        }
        return 2;
    };

    bool any_unpicked = picked.size() < queue.size();
    for (size_t pi = 0; pi < picked.size(); ++pi) {
        const Picked &pk = picked[pi];
        while (cur_stage < pk.stage) {
            if (cur_stage >= 0) close_stage(cur_stage);
            ++cur_stage;
            prog.stage[cur_stage].op_begin = (int16_t)n_ops;
            flip = 0;
        }
        Gate res;
        const Gate g = queue[pk.queue_idx]; /* (a copy: residuals may rewrite later queue entries) */
        if (lower(g, pk.stage, res)) {
            const int where = fold_forward(res, pi);
            if (where == 1 || (where == 2 && !any_unpicked)) {
                /* apply the phase here, as a diagonal op of this stage (where == 1: the slot was
                 * reserved when the gate was picked; otherwise only if the pass has room) */
                Gate none;
                const int still_to_come = (int)(picked.size() - pi - 1);
                if (where == 1 || (n_ops + still_to_come < op_limit && n_ops + still_to_come < QGB_MAX_OPS)) {
                    lower(res, pk.stage, none);
                    ++stats.residual_ops;
                } else {
                    deferred.push_back(res);
                }
            } else if (where == 2) {
                deferred.push_back(res);
            }
        }
    }
    if (cur_stage >= 0) close_stage(cur_stage);
    for (int s = cur_stage + 1; s < n_stages; ++s) {
        prog.stage[s].op_begin = prog.stage[s].op_end = (int16_t)n_ops;
        prog.stage[s].store_flip = 0;
        prog.stage[s].store_xor = 0;
    }
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
        /* phases nothing could absorb: their place is right after this pass */
        if (!deferred.empty()) queue.insert(queue.begin(), deferred.begin(), deferred.end());
    }
    stats.fan_ops = prog.n_fans;
    stats.gates_in_pass = (int)picked.size();
    stats.ops_in_pass = n_ops;
    stats.stages_in_pass = n_stages;
}

template void plan_pass<float>(std::vector<Gate> &, int, const PlanConfig &, PassProgram<float> &,
                               PlanStats &);
template void plan_pass<double>(std::vector<Gate> &, int, const PlanConfig &, PassProgram<double> &,
                                PlanStats &);

} // namespace qgb
