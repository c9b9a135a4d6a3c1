/*
 * planner_emul.cpp — CPU test of the flush planner (qgate_b200/csrc/planner.cpp).
 *
 * Interprets the pass programs exactly as kernels_tile.cu's tile_pass_kernel does (same tile
 * gather, same thread/register bit maps, same per-op semantics, ops applied per "thread") on
 * a host array, and compares with applying the submitted gates one by one in submission
 * order (the reference semantics, CPUQubitProcessor.cpp:307-362).  Also checks that every
 * stage's (thread, register) map covers the tile exactly once.
 *
 * Test infrastructure only: built and run by tests/test_planner.py, never shipped.
 */
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "planner.h"

using namespace qgb;
typedef std::complex<double> cd;

static int g_failures = 0;
static bool g_expect_tma = false;
static long g_warp_local = 0;
static long g_mux_thr = 0, g_mux_reg = 0, g_mux_out = 0;
static long g_fan_ops = 0, g_fan_thr = 0, g_fan_out = 0, g_fan_reg = 0, g_fan_codes[4] = {0, 0, 0, 0};
static long g_shear = 0, g_direct = 0, g_flipped_stages = 0, g_residual_ops = 0, g_parity_split = 0, g_parity_gates = 0;
#define CHECK(cond, ...)                      \
    do {                                      \
        if (!(cond)) {                        \
            ++g_failures;                     \
            std::printf("FAIL: " __VA_ARGS__); \
            std::printf("\n");                \
        }                                     \
    } while (0)

static void apply_direct(std::vector<cd> &amp, int n, const Gate &g) {
    if (!g.fan.empty()) { /* product of CP(target, lane) */
        for (uint64_t i = 0; i < (1ull << n); ++i) {
            if (!((i >> g.target) & 1ull)) continue;
            for (const Gate::FanTerm &t : g.fan)
                if ((i >> t.lane) & 1ull) amp[i] *= cd(t.re, t.im);
        }
        return;
    }
    const cd m00(g.m[0], g.m[1]), m01(g.m[2], g.m[3]), m10(g.m[4], g.m[5]), m11(g.m[6], g.m[7]);
    if (g.parity) { /* d0 on an even number of set lanes, d1 on an odd number */
        for (uint64_t i = 0; i < (1ull << n); ++i)
            if ((i & g.ctrl_mask) == g.ctrl_mask) amp[i] *= (__builtin_popcountll(i & g.parity) & 1) ? m11 : m00;
        return;
    }
    const uint64_t tb = 1ull << g.target;
    for (uint64_t i = 0; i < (1ull << n); ++i) {
        if (i & tb) continue;
        if ((i & g.ctrl_mask) != g.ctrl_mask) continue;
        const cd q0 = amp[i], q1 = amp[i | tb];
        amp[i] = m00 * q0 + m01 * q1;
        amp[i | tb] = m10 * q0 + m11 * q1;
    }
}

template <typename real>
static void emulate_pass(const PassProgram<real> &p, std::vector<cd> &amp) {
    const int n = p.n_lanes, T = p.T, L = p.L, K = p.K;
    CHECK(T >= K && L <= T && T <= n, "bad shape n=%d T=%d L=%d K=%d", n, T, L, K);
    for (int i = 0; i < L; ++i) CHECK(p.tile_lane[i] == i, "low lane %d not in the tile", i);
    for (int i = 1; i < T; ++i) CHECK(p.tile_lane[i] > p.tile_lane[i - 1], "tile lanes not ascending");
    const uint32_t tile_size = 1u << T, lowmask = (1u << L) - 1u;
    std::vector<uint64_t> choff(1u << (T - L));
    for (uint32_t c = 0; c < choff.size(); ++c) {
        uint64_t off = 0;
        for (int b = 0; b < T - L; ++b) off |= (uint64_t)((c >> b) & 1u) << p.tile_lane[L + b];
        choff[c] = off;
    }
    /* the out-of-tile references the TMA-staged kernel turns into per-tile op masks */
    {
        int expect = 0;
        for (int o = 0; o < p.n_ops; ++o) {
            const Op<real> &op = p.op[o];
            const bool has_ref = expect < p.n_out && p.out[expect].op == o;
            if ((op.kind == OP_GEN || op.kind == OP_SHEAR) && (op.arm & ARM_MUX_OUT))
                CHECK(has_ref && p.out[expect].sel_mask == (1ull << op.mux_out), "multiplexer of op %d not referenced", o);
            if (op.kind == OP_DIAG_OUT) CHECK(has_ref && p.out[expect].sel_mask != 0, "outside diagonal %d not referenced", o);
            if (op.kind == OP_DIAG || op.kind == OP_DIAG_OUT)
                CHECK(op.m1[0] == op.m[2] && op.m1[1] == op.m[3] && op.m1[2] == op.m[0] && op.m1[3] == op.m[1],
                      "diag op %d: factors not mirrored into m1", o);
            if (op.ctrl_out != 0) CHECK(has_ref, "outside controls of op %d not referenced", o);
            if (!has_ref) continue;
            CHECK(p.out[expect].ctrl_mask == op.ctrl_out, "out-of-tile reference of op %d wrong", o);
            for (int lane = 0; lane < n; ++lane)
                if ((p.out[expect].sel_mask | p.out[expect].ctrl_mask) & (1ull << lane)) {
                    bool in_tile = false;
                    for (int t = 0; t < T; ++t) in_tile = in_tile || p.tile_lane[t] == lane;
                    CHECK(!in_tile, "op %d: tile lane %d referenced as outside the tile", o, lane);
                }
            ++expect;
        }
        CHECK(expect == p.n_out, "%d out-of-tile references, expected %d", p.n_out, expect);
    }
    /* warp-local stage transitions: where a stage is flagged, the next stage's warps must own
     * exactly the tile elements they owned in it (the kernel only synchronises the warp there) */
    {
        auto owner_map = [&](const Stage &st) {
            std::vector<int> owner(tile_size, -1);
            for (uint32_t tid = 0; tid < (1u << (T - K)); ++tid) {
                uint32_t ebase = 0;
                for (int i = 0; i < T - K; ++i) ebase |= ((tid >> i) & 1u) << st.W[i];
                for (int r = 0; r < (1 << K); ++r) {
                    uint32_t o = 0;
                    for (int j = 0; j < K; ++j)
                        if (r & (1 << j)) o |= 1u << st.R[j];
                    owner[ebase | o] = (int)(tid >> 5);
                }
            }
            return owner;
        };
        for (int s = 0; s + 1 < p.n_stages; ++s) {
            if (!p.stage[s].warp_local) continue;
            ++g_warp_local;
            CHECK(T - K > 5, "warp-local flag on a single-warp tile");
            const std::vector<int> a = owner_map(p.stage[s]), b = owner_map(p.stage[s + 1]);
            bool same = true;
            for (uint32_t e = 0; e < tile_size; ++e) same = same && a[e] == b[e] && a[e] >= 0;
            CHECK(same, "stage %d is flagged warp-local but stage %d moves elements between warps", s, s + 1);
        }
        CHECK(p.stage[p.n_stages - 1].warp_local == 0, "the last stage must end with a CTA barrier");
    }
    std::vector<cd> tile(tile_size);
    if (g_expect_tma) CHECK(p.n_groups >= 1 && p.n_groups <= QGB_MAX_GROUPS, "tile does not fit a tensor map (%d groups)", p.n_groups);
    for (uint64_t bid = 0; bid < (1ull << (n - T)); ++bid) {
        uint64_t base = 0;
        for (int i = 0; i < n - T; ++i) base |= ((bid >> i) & 1ull) << p.rest_lane[i];
        if (p.n_groups >= 1) {
            /* the TMA view of the same tile (kernels_tma.cu): box element e of the tensor map whose
             * dimension 0 is the 128-byte row and dimension d + 1 is lane group d must be the
             * amplitude the cp.async gather puts at tile element e */
            int consumed = 0;
            uint64_t coord[QGB_MAX_GROUPS];
            uint64_t tbase = 0;
            int tbits_total = p.row_lanes;
            for (int d = 0; d < p.n_groups; ++d) {
                const uint64_t rest = (bid >> consumed) & ((1ull << p.grp_r[d]) - 1ull);
                coord[d] = rest << p.grp_t[d];
                tbase |= rest << (p.grp_start[d] + p.grp_t[d]);
                consumed += p.grp_r[d];
                tbits_total += p.grp_t[d];
                CHECK(p.grp_t[d] <= 8 && p.grp_t[d] + p.grp_r[d] <= 32, "group %d exceeds the tensor-map limits", d);
            }
            CHECK(consumed == n - T && tbits_total == T, "groups do not cover the lanes (%d rest, %d tile)", consumed, tbits_total);
            CHECK(tbase == base, "tile origin from groups differs");
            for (uint32_t e = 0; e < tile_size; e += 37) {
                uint64_t idx = e & ((1u << p.row_lanes) - 1u);
                uint32_t rem = e >> p.row_lanes;
                for (int d = 0; d < p.n_groups; ++d) {
                    const uint64_t x = rem & ((1u << p.grp_t[d]) - 1u);
                    rem >>= p.grp_t[d];
                    idx += (coord[d] + x) << p.grp_start[d];
                }
                CHECK(idx == (base | choff[e >> L] | (e & lowmask)), "tensor-map element %u maps to a different amplitude", e);
            }
        }
        for (uint32_t e = 0; e < tile_size; ++e) tile[e] = amp[base | choff[e >> L] | (e & lowmask)];
        for (int s = 0; s < p.n_stages; ++s) {
            const Stage &st = p.stage[s];
            if (st.op_begin == st.op_end) continue;
            std::vector<int> seen(tile_size, 0);
            uint32_t rb[QGB_MAX_REG_BITS];
            for (int j = 0; j < K; ++j) rb[j] = 1u << st.R[j];
            for (int r = 0; r < (1 << K); ++r) {
                uint32_t o = 0;
                for (int j = 0; j < K; ++j)
                    if (r & (1 << j)) o |= rb[j];
                CHECK(st.sro[r] == tile_swizzle(o, sizeof(real) == 4), "stage %d: bad swizzled offset of register %d", s, r);
                uint32_t x = 0;
                for (int j = 0; j < K; ++j)
                    if (r & (1 << j)) x ^= st.xb[j];
                CHECK(x == st.sro[r] * 2u * sizeof(real), "stage %d: byte XOR constants disagree with register %d", s, r);
            }
            for (uint32_t tid = 0; tid < (1u << (T - K)); ++tid) {
                uint32_t ebase = 0;
                for (int i = 0; i < T - K; ++i) ebase |= ((tid >> i) & 1u) << st.W[i];
                auto roff = [&](int r) {
                    uint32_t o = 0;
                    for (int j = 0; j < K; ++j)
                        if (r & (1 << j)) o |= rb[j];
                    return o;
                };
                std::vector<cd> a(1u << K);
                for (int r = 0; r < (1 << K); ++r) {
                    a[r] = tile[ebase | roff(r)];
                    seen[ebase | roff(r)]++;
                }
                uint32_t rmask = 0;
                for (int j = 0; j < K; ++j) rmask |= rb[j];
                for (int o = st.op_begin; o < st.op_end; ++o) {
                    const Op<real> &op = p.op[o];
                    if ((base & op.ctrl_out) != op.ctrl_out) continue;
                    CHECK((op.cmt & rmask) == 0, "thread-part controls overlap the register bits");
                    /* parity of the lanes outside the tile that select the op's second matrix / factor */
                    bool out_sel = false;
                    for (int i = 0; i < p.n_out; ++i)
                        if (p.out[i].op == o) out_sel = __builtin_popcountll(base & p.out[i].sel_mask) & 1;
                    {
                        uint32_t want = 0;
                        if (op.kind == OP_GEN || op.kind == OP_SHEAR) want = ARM_GEN(op.bit) | (op.arm & (ARM_MUX_THR | ARM_MUX_REG | ARM_MUX_OUT));
                        else if (op.kind == OP_SWAP) want = ARM_SWAP(op.bit);
                        else if (op.kind == OP_FAN) want = ARM_DIAG_THR;
                        else want = op.regsel ? ARM_DIAG_REG : ARM_DIAG_THR;
                        CHECK(op.arm == want, "arm selector %u of op kind %d bit %d", op.arm, op.kind, op.bit);
                        /* Op::code, the body selector of the TMA-staged kernel */
                        const uint32_t all_regs = (1u << (1 << K)) - 1u;
                        int code = -1;
                        const bool sh = op.kind == OP_SHEAR;
                        if ((op.kind == OP_GEN || sh) && (op.arm & ARM_MUX_REG)) {
                            for (int j2 = 0; j2 < K; ++j2) {
                                uint32_t sel = 0;
                                for (int r = 0; r < (1 << K); ++r)
                                    if (r & (1 << j2)) sel |= 1u << r;
                                if (sel == op.regsel) code = sh ? OPC_SHEAR_REGMUX(op.bit, j2) : OPC_GEN_REGMUX(op.bit, j2);
                            }
                            CHECK(op.regmask == all_regs && op.cmt == 0 && op.ctrl_out == 0, "multiplexed op with controls");
                        } else if (op.kind == OP_GEN || sh) {
                            if (sh)
                                code = op.regmask == all_regs ? OPC_SHEAR(op.bit) : OPC_SHEAR_MASKED(op.bit);
                            else
                                code = op.regmask == all_regs ? OPC_GEN(op.bit) : OPC_GEN_MASKED(op.bit);
                        } else if (op.kind == OP_SWAP) {
                            code = OPC_SWAP(op.bit);
                        } else if (op.kind == OP_FAN) {
                            code = op.regmask == all_regs ? OPC_FAN_ALL : OPC_FAN;
                            for (int j = 0; j < K; ++j)
                                for (int pol = 0; pol < 2; ++pol) {
                                    uint32_t pattern = 0;
                                    for (int r = 0; r < (1 << K); ++r)
                                        if (((r >> j) & 1) == pol) pattern |= 1u << r;
                                    if (op.regmask == pattern) code = OPC_FAN_REG(j, pol);
                                }
                            if (tid == 0 && bid == 0) g_fan_codes[code == OPC_FAN ? 0 : code == OPC_FAN_ALL ? 1 : 2 + (code & 1)]++;
                        } else {
                            code = op.regsel ? OPC_DIAG_REG : OPC_DIAG_THR;
                        }
                        CHECK(op.code == code, "op code %d, expected %d", op.code, code);
                    }
                    const bool active = (ebase & op.cmt) == op.cmt;
                    const cd m0(op.m[0], op.m[1]), m1(op.m[2], op.m[3]), m2(op.m[4], op.m[5]),
                        m3(op.m[6], op.m[7]);
                    if (op.kind == OP_SHEAR) {
                        /* q0 += y q1; q1 += g q0; q0 += x q1 on the pairs of register bit `bit`, exactly
                         * as kernels_tma.cu does it (in place, in `real`) */
                        CHECK(op.bit >= 0 && op.bit < K, "register bit %d out of range", op.bit);
                        ++g_shear;
                        for (int r0 = 0; r0 < (1 << K); ++r0) {
                            if (r0 & (1 << op.bit)) continue;
                            const int r1 = r0 | (1 << op.bit);
                            if (!((op.regmask >> r0) & 1u) || !active) continue;
                            bool one = false;
                            if (op.arm & ARM_MUX_THR) one = (ebase & op.tsel) != 0;
                            if (op.arm & ARM_MUX_REG) one = (op.regsel >> r0) & 1u;
                            if (op.arm & ARM_MUX_OUT) one = out_sel;
                            if (tid == 0 && r0 == 0 && bid == 0) {
                                g_mux_thr += (op.arm & ARM_MUX_THR) != 0;
                                g_mux_reg += (op.arm & ARM_MUX_REG) != 0;
                                g_mux_out += (op.arm & ARM_MUX_OUT) != 0;
                            }
                            const real *k = one ? op.m1 : op.m;
                            const cd y(k[0], k[1]), g(k[2], k[3]), x(k[4], k[5]);
                            a[r0] += y * a[r1];
                            a[r1] += g * a[r0];
                            a[r0] += x * a[r1];
                        }
                    } else if (op.kind == OP_GEN || op.kind == OP_SWAP) {
                        if (op.kind == OP_GEN) ++g_direct;
                        CHECK(op.bit >= 0 && op.bit < K, "register bit %d out of range", op.bit);
                        for (int r0 = 0; r0 < (1 << K); ++r0) {
                            if (r0 & (1 << op.bit)) continue;
                            const int r1 = r0 | (1 << op.bit);
                            if (!((op.regmask >> r0) & 1u) || !active) continue;
                            const cd q0 = a[r0], q1 = a[r1];
                            if (op.kind == OP_GEN) {
                                if (tid == 0 && r0 == 0 && bid == 0) {
                                    g_mux_thr += (op.arm & ARM_MUX_THR) != 0;
                                    g_mux_reg += (op.arm & ARM_MUX_REG) != 0;
                                    g_mux_out += (op.arm & ARM_MUX_OUT) != 0;
                                }
                                bool one = false;
                                if (op.arm & ARM_MUX_THR) one = (ebase & op.tsel) != 0;
                                if (op.arm & ARM_MUX_REG) one = (op.regsel >> r0) & 1u;
                                if (op.arm & ARM_MUX_OUT) one = out_sel;
                                if (one) {
                                    const cd n0(op.m1[0], op.m1[1]), n1(op.m1[2], op.m1[3]),
                                        n2(op.m1[4], op.m1[5]), n3(op.m1[6], op.m1[7]);
                                    a[r0] = n0 * q0 + n1 * q1;
                                    a[r1] = n2 * q0 + n3 * q1;
                                } else {
                                    a[r0] = m0 * q0 + m1 * q1;
                                    a[r1] = m2 * q0 + m3 * q1;
                                }
                            } else {
                                a[r0] = q1;
                                a[r1] = q0;
                            }
                        }
                    } else if (op.kind == OP_DIAG || op.kind == OP_DIAG_OUT) {
                        CHECK(op.kind == OP_DIAG || (op.tsel == 0 && op.regsel == 0), "DIAG_OUT with tile target");
                        CHECK((op.tsel & rmask) == 0, "diag thread target overlaps the register bits");
                        if (op.tsel && op.regsel) ++g_parity_split;
                        for (int r = 0; r < (1 << K); ++r) {
                            if (!((op.regmask >> r) & 1u) || !active) continue;
                            /* d1 where an odd number of the gate's lanes is set: register part ^
                             * thread part ^ the part common to the tile */
                            const bool one = (((op.regsel >> r) & 1u) != 0) ^ ((__builtin_popcount(ebase & op.tsel) & 1) != 0) ^ out_sel;
                            a[r] *= one ? m1 : m0;
                        }
                    } else if (op.kind == OP_FAN) {
                        /* as kernels_tma.cu: one factor per thread (thread-bit terms) and tile (outside
                         * terms) on the registers of regmask, then one phase per register-bit term */
                        CHECK(op.bit >= 0 && op.bit < p.n_fans, "fan index %d out of range", op.bit);
                        const auto &fn = p.fan[op.bit];
                        CHECK(fn.first >= 0 && fn.first + fn.n_thr + fn.n_out <= p.n_fan_terms, "fan terms out of range");
                        if (tid == 0 && bid == 0) ++g_fan_ops, g_fan_thr += fn.n_thr, g_fan_out += fn.n_out, g_fan_reg += __builtin_popcount(fn.n_reg);
                        cd f(fn.base[0], fn.base[1]);
                        for (int k = 0; k < fn.n_thr; ++k) {
                            const auto &t = p.fan_term[fn.first + k];
                            CHECK(t.bit >= 0 && t.bit < T - K, "fan thread bit %d out of range", t.bit);
                            if ((tid >> t.bit) & 1u) f *= cd(t.re, t.im);
                        }
                        for (int k = 0; k < fn.n_out; ++k) {
                            const auto &t = p.fan_term[fn.first + fn.n_thr + k];
                            bool in_tile = false;
                            for (int tt = 0; tt < T; ++tt) in_tile = in_tile || p.tile_lane[tt] == t.bit;
                            CHECK(!in_tile && t.bit >= 0 && t.bit < n, "fan outside lane %d is a tile lane", t.bit);
                            if ((base >> t.bit) & 1ull) f *= cd(t.re, t.im);
                        }
                        /* per register: the factors of its set bits (kernels_tma.cu walks the registers in
                         * Gray-code order with a running product) */
                        for (int j = 0; j < QGB_MAX_REG_BITS; ++j)
                            CHECK(((fn.n_reg >> j) & 1) || (fn.reg[j][0] == 1. && fn.reg[j][1] == 0.), "fan register bit %d: factor without a term", j);
                        CHECK((fn.n_reg >> K) == 0, "fan register bits out of range");
                        for (int r = 0; r < (1 << K); ++r) {
                            if (!((op.regmask >> r) & 1u) || !active) continue;
                            cd fr = f;
                            for (int j = 0; j < K; ++j)
                                if (r & (1 << j)) fr *= cd(fn.reg[j][0], fn.reg[j][1]);
                            a[r] *= cd((real)fr.real(), (real)fr.imag());
                        }
                    } else {
                        CHECK(false, "unknown op kind %d", op.kind);
                    }
                }
                /* registers relabelled by the stage's shears go back to the slots they now stand for */
                uint32_t sx = 0;
                for (int j = 0; j < K; ++j)
                    if (st.store_flip & (1u << j)) sx ^= st.xb[j];
                CHECK(sx == st.store_xor && st.store_flip < (1u << K), "stage %d: store_xor disagrees with store_flip", s);
                for (int r = 0; r < (1 << K); ++r) tile[ebase | roff(r ^ (int)st.store_flip)] = a[r];
            }
            for (uint32_t e = 0; e < tile_size; ++e)
                if (seen[e] != 1) {
                    CHECK(false, "stage %d covers tile element %u %d times", s, e, seen[e]);
                    break;
                }
        }
        for (uint32_t e = 0; e < tile_size; ++e) amp[base | choff[e >> L] | (e & lowmask)] = tile[e];
    }
}

static Gate random_gate(std::mt19937_64 &rng, int n, int max_ctrl) {
    std::uniform_real_distribution<double> u(-3.14159, 3.14159);
    Gate g;
    g.target = (int)(rng() % n);
    g.ctrl_mask = 0;
    int n_ctrl = (int)(rng() % (max_ctrl + 1));
    for (int c = 0; c < n_ctrl; ++c) {
        int lane = (int)(rng() % n);
        if (lane != g.target) g.ctrl_mask |= 1ull << lane;
    }
    for (int i = 0; i < 8; ++i) g.m[i] = 0.;
    const int kind = (int)(rng() % 7);
    const double a = u(rng), b = u(rng), c = u(rng);
    switch (kind) {
    case 0: /* general unitary */
    case 1:
        g.m[0] = std::cos(a) * std::cos(b);  g.m[1] = std::cos(a) * std::sin(b);
        g.m[2] = -std::sin(a) * std::cos(c); g.m[3] = -std::sin(a) * std::sin(c);
        g.m[4] = std::sin(a) * std::cos(-c); g.m[5] = std::sin(a) * std::sin(-c);
        g.m[6] = std::cos(a) * std::cos(-b); g.m[7] = std::cos(a) * std::sin(-b);
        break;
    case 2: /* phase (d0 == 1) */
        g.m[0] = 1.;
        g.m[6] = std::cos(a); g.m[7] = std::sin(a);
        break;
    case 3: /* general diagonal */
        g.m[0] = std::cos(a); g.m[1] = std::sin(a);
        g.m[6] = std::cos(b); g.m[7] = std::sin(b);
        break;
    case 4: /* X */
        g.m[2] = 1.; g.m[4] = 1.;
        break;
    case 6: { /* parity diagonal exp(i a Z x ... x Z) over 2..4 lanes, target = the lowest */
        uint64_t lanes = 0;
        const int want = 2 + (int)(rng() % 3);
        for (int t = 0; t < 8 && __builtin_popcountll(lanes) < want && __builtin_popcountll(lanes) < n; ++t) lanes |= 1ull << (rng() % n);
        g.parity = lanes;
        g.target = __builtin_ctzll(lanes);
        g.ctrl_mask &= ~lanes;
        g.m[0] = std::cos(a); g.m[1] = std::sin(a);
        g.m[6] = std::cos(a); g.m[7] = -std::sin(a);
        ++g_parity_gates;
        break;
    }
    case 5: /* Y-like anti-diagonal */
        g.m[2] = std::cos(a); g.m[3] = std::sin(a);
        g.m[4] = std::cos(b); g.m[5] = std::sin(b);
        break;
    }
    return g;
}

/* QFT-like circuits: H on lane i, then controlled phases between lane i and every other lane (in
 * either direction, in a shuffled order, some twice), sprinkled with random gates */
template <typename real>
static void run_qft_case(int n, int T, int L, int K, int max_cost, bool textbook, int noise, uint64_t seed) {
    std::mt19937_64 rng(seed);
    std::vector<cd> ref(1ull << n), amp;
    std::normal_distribution<double> nd;
    for (auto &v : ref) v = cd(nd(rng), nd(rng));
    amp = ref;
    std::vector<Gate> queue;
    int n_gates = 0, n_merged = 0;
    auto submit = [&](const Gate &g) {
        apply_direct(ref, n, g);
        n_merged += enqueue_gate(queue, g, true) ? 1 : 0;
        ++n_gates;
    };
    const double r = std::sqrt(0.5);
    auto had = [&](int lane) {
        Gate g;
        for (int e = 0; e < 8; ++e) g.m[e] = 0.;
        g.m[0] = g.m[2] = g.m[4] = r, g.m[6] = -r;
        g.target = lane, g.ctrl_mask = 0;
        return g;
    };
    auto cp = [&](int ctrl, int target, double phi) {
        Gate g;
        for (int e = 0; e < 8; ++e) g.m[e] = 0.;
        g.m[0] = 1., g.m[6] = std::cos(phi), g.m[7] = std::sin(phi);
        g.target = target, g.ctrl_mask = 1ull << ctrl;
        return g;
    };
    for (int i = 0; i < n; ++i) {
        if (!textbook) submit(had(i));
        std::vector<int> others;
        if (textbook)
            for (int j = 0; j < i; ++j) others.push_back(j);
        else
            for (int j = i + 1; j < n; ++j) others.push_back(j);
        std::shuffle(others.begin(), others.end(), rng);
        for (int j : others) {
            const double phi = 3.141592653589793 / (double)(1 << std::abs(j - i));
            submit((rng() & 1) ? cp(j, i, phi) : cp(i, j, phi));
            if (rng() % 7 == 0) submit(cp(j, i, 0.3));
            if (noise && (int)(rng() % 10) < noise) submit(random_gate(rng, n, 2));
        }
        if (textbook) submit(had(i));
    }
    PlanConfig cfg;
    cfg.fp32 = sizeof(real) == 4;
    cfg.K = K;
    cfg.max_cost = max_cost;
    cfg.T = T;
    cfg.L = L;
    cfg.max_ops = 32;
    cfg.shear = true;
    cfg.row_lanes = cfg.fp32 ? 4 : 3;
    cfg.max_groups = QGB_MAX_GROUPS;
    cfg.L = std::max(cfg.L, cfg.row_lanes);
    g_expect_tma = n > cfg.row_lanes && std::min(T, n) >= cfg.row_lanes;
    static PassProgram<real> prog;
    int n_pass = 0;
    const long fans_before = g_fan_ops;
    while (!queue.empty()) {
        PlanStats st;
        plan_pass<real>(queue, n, cfg, prog, st);
        CHECK(st.gates_in_pass > 0 && n_pass < 100000, "planner made no progress");
        if (st.gates_in_pass <= 0 || n_pass >= 100000) return;
        CHECK(prog.n_fans <= QGB_MAX_FANS && prog.n_fan_terms <= QGB_MAX_FAN_TERMS, "fan limits exceeded");
        emulate_pass<real>(prog, amp);
        ++n_pass;
    }
    double err = 0., scale = 0.;
    for (size_t i = 0; i < ref.size(); ++i) {
        err = std::max(err, std::abs(ref[i] - amp[i]));
        scale = std::max(scale, std::abs(ref[i]));
    }
    const double tol = sizeof(real) == 4 ? 2e-4 : 1e-11;
    CHECK(err <= tol * scale, "qft-like n=%d T=%d L=%d: rel err %.3g", n, T, L, err / scale);
    std::printf("ok %s qft-like%s n=%2d T=%2d L=%d gates=%4d merged=%4d passes=%3d fan ops=%3ld relerr=%.2e\n",
                sizeof(real) == 4 ? "f32" : "f64", textbook ? " (textbook order)" : "", n, T, L, n_gates, n_merged, n_pass,
                g_fan_ops - fans_before, err / scale);
}

template <typename real>
static void run_case(int n, int T, int L, int n_gates, int max_ctrl, int max_ops, int max_stages,
                     bool merge, uint64_t seed, bool tma = false, int reg_bits = 0, int max_cost = 0, bool shear = false) {
    std::mt19937_64 rng(seed);
    std::vector<cd> ref(1ull << n), amp;
    std::normal_distribution<double> nd;
    for (auto &v : ref) v = cd(nd(rng), nd(rng));
    amp = ref;
    std::vector<Gate> queue;
    int n_merged = 0;
    for (int i = 0; i < n_gates; ++i) {
        Gate g = random_gate(rng, n, max_ctrl);
        apply_direct(ref, n, g);
        n_merged += enqueue_gate(queue, g, merge) ? 1 : 0;
    }
    PlanConfig cfg;
    cfg.fp32 = sizeof(real) == 4;
    cfg.K = reg_bits > 0 ? reg_bits : (cfg.fp32 ? 4 : 3);
    if (max_cost > 0) cfg.max_cost = max_cost;
    cfg.T = T;
    cfg.L = L;
    cfg.max_ops = max_ops;
    cfg.max_stages = max_stages;
    cfg.warp_local = tma;
    cfg.shear = shear;
    if (tma) { /* what engine.cu sets for the TMA-staged kernel */
        cfg.row_lanes = cfg.fp32 ? 4 : 3;
        cfg.max_groups = QGB_MAX_GROUPS;
        cfg.L = std::max(cfg.L, cfg.row_lanes);
    }
    g_expect_tma = tma && n > cfg.row_lanes && std::min(T, n) >= cfg.row_lanes;
    static PassProgram<real> prog;
    int n_pass = 0, n_exec = 0, n_deferred = 0;
    while (!queue.empty()) {
        const size_t before = queue.size();
        PlanStats st;
        plan_pass<real>(queue, n, cfg, prog, st);
        /* (a pass may hand back one phase gate per sheared gate it executed: count gates, not queue length) */
        CHECK(st.gates_in_pass > 0 && n_pass < 100000, "planner made no progress");
        if (st.gates_in_pass <= 0 || n_pass >= 100000) return;
        CHECK(queue.size() + st.gates_in_pass >= before, "gate accounting");
        CHECK(prog.n_ops <= max_ops && prog.n_stages <= std::max(1, max_stages), "limits exceeded");
        emulate_pass<real>(prog, amp);
        ++n_pass;
        n_exec += st.gates_in_pass;
        n_deferred += (int)queue.size() - ((int)before - st.gates_in_pass);
        g_residual_ops += st.residual_ops;
        for (int s = 0; s < prog.n_stages; ++s) g_flipped_stages += prog.stage[s].store_flip != 0;
        CHECK(shear || (st.shear_ops == 0 && st.residual_ops == 0), "shear ops planned without the option");
    }
    CHECK(n_exec + n_merged == n_gates + n_deferred, "executed %d + merged %d != %d + %d deferred phases", n_exec, n_merged, n_gates, n_deferred);
    double err = 0., scale = 0.;
    for (size_t i = 0; i < ref.size(); ++i) {
        err = std::max(err, std::abs(ref[i] - amp[i]));
        scale = std::max(scale, std::abs(ref[i]));
    }
    const double tol = sizeof(real) == 4 ? 2e-4 : 1e-11; /* op matrices are rounded to `real` */
    CHECK(err <= tol * scale, "n=%d T=%d L=%d gates=%d: rel err %.3g", n, T, L, n_gates, err / scale);
    std::printf("ok %s n=%2d T=%2d L=%d gates=%4d passes=%3d merged=%3d relerr=%.2e%s%s\n",
                sizeof(real) == 4 ? "f32" : "f64", n, T, L, n_gates, n_pass, n_merged, err / scale, tma ? " tma" : "",
                shear ? " shear" : "");
}

int main() {
    uint64_t seed = 1;
    for (int n : {4, 5, 8, 10, 12}) {
        for (int T : {5, 7, 9, 12}) {
            for (int L : {1, 3, 5}) {
                if (T < 4) continue;
                run_case<double>(n, T, L, 150, 3, QGB_MAX_OPS, QGB_MAX_STAGES, true, seed++);
                run_case<float>(n, T, L, 60, 2, QGB_MAX_OPS, QGB_MAX_STAGES, true, seed++);
            }
        }
    }
    /* tile shapes constrained to what a 5-dimensional tensor map can describe (TMA staging) */
    for (int n : {8, 10, 13, 14}) {
        for (int T : {8, 9, 11}) {
            for (int L : {1, 4, 6}) {
                run_case<double>(n, T, L, 200, 3, QGB_MAX_OPS, QGB_MAX_STAGES, true, seed++, true);
                run_case<float>(n, std::max(T, 9), L, 80, 2, QGB_MAX_OPS, QGB_MAX_STAGES, true, seed++, true);
            }
        }
    }
    /* the engine's defaults: 16 amplitudes per thread in both precisions, 16 KiB tiles, at most 32
     * ops and a summed cost of 24 per pass */
    for (int n : {10, 12, 14}) {
        run_case<double>(n, 10, 5, 300, 3, 32, QGB_MAX_STAGES, true, seed++, true, 4, 24);
        run_case<double>(n, 11, 3, 300, 9, 32, QGB_MAX_STAGES, true, seed++, true, 4, 32);
        run_case<float>(n, 11, 6, 120, 3, 32, QGB_MAX_STAGES, true, seed++, true, 4, 24);
    }
    /* dense gates as three in-place shears (OP_SHEAR): relabelled registers, folded phases */
    for (int n : {9, 10, 12, 14}) {
        for (int ctrl : {0, 1, 3}) {
            run_case<double>(n, 10, 5, 300, ctrl, 32, QGB_MAX_STAGES, true, seed++, true, 4, 24, true);
            run_case<double>(n, 9, 3, 250, ctrl, 32, QGB_MAX_STAGES, false, seed++, true, 3, 48, true);
            run_case<float>(n, 11, 6, 100, ctrl, 32, QGB_MAX_STAGES, true, seed++, true, 4, 24, true);
        }
        run_case<double>(n, 10, 5, 300, 2, 6, 3, true, seed++, true, 4, 0, true);
    }
    /* phase fans (OP_FAN): QFT-like circuits in both gate orders, clean and with random gates mixed in */
    for (int n : {9, 12, 14, 16}) {
        for (int noise : {0, 2}) {
            run_qft_case<double>(n, 10, 4, 4, 24, false, noise, seed++);
            run_qft_case<double>(n, 10, 4, 4, 24, true, noise, seed++);
            run_qft_case<float>(n, 11, 5, 4, 21, false, noise, seed++);
            run_qft_case<double>(n, 9, 3, 3, 1 << 20, true, noise, seed++);
        }
    }
    g_expect_tma = false;
    run_case<double>(11, 8, 3, 400, 2, 16, 3, true, seed++, false, 0, 0, true);
    /* limits and no-merge paths */
    run_case<double>(10, 7, 2, 300, 3, 5, QGB_MAX_STAGES, false, seed++);
    run_case<double>(10, 7, 2, 300, 3, QGB_MAX_OPS, 2, true, seed++);
    run_case<double>(11, 8, 3, 400, 9, 16, 3, true, seed++);
    run_case<float>(11, 9, 4, 200, 4, 7, 2, false, seed++);
    if (g_failures) {
        std::printf("%d FAILURES\n", g_failures);
        return 1;
    }
    std::printf("multiplexed ops seen: thread-bit %ld, register-bit %ld, outside-tile %ld\n", g_mux_thr, g_mux_reg,
                g_mux_out);
    if (g_mux_thr == 0 || g_mux_reg == 0 || g_mux_out == 0) {
        std::printf("FAIL: a multiplexer placement was never exercised\n");
        return 1;
    }
    std::printf("shear ops %ld, direct 2x2 ops %ld, stages stored through a relabelling %ld, residual phase ops %ld\n",
                g_shear, g_direct, g_flipped_stages, g_residual_ops);
    std::printf("parity diagonals %ld, of which split over register and thread bits in a stage %ld\n", g_parity_gates, g_parity_split);
    if (g_parity_gates == 0 || g_parity_split == 0) {
        std::printf("FAIL: parity diagonals were never exercised\n");
        return 1;
    }
    if (g_shear == 0 || g_direct == 0 || g_flipped_stages == 0 || g_residual_ops == 0) {
        std::printf("FAIL: a shear path was never exercised\n");
        return 1;
    }
    std::printf("phase fans %ld: terms on thread bits %ld, outside the tile %ld, on register bits %ld\n", g_fan_ops, g_fan_thr,
                g_fan_out, g_fan_reg);
    std::printf("fan bodies: masked %ld, every register %ld, half of the registers by a clear / set bit %ld / %ld\n", g_fan_codes[0],
                g_fan_codes[1], g_fan_codes[2], g_fan_codes[3]);
    if (g_fan_ops == 0 || g_fan_thr == 0 || g_fan_out == 0 || g_fan_reg == 0 || g_fan_codes[1] == 0 || g_fan_codes[2] == 0 ||
        g_fan_codes[3] == 0) {
        std::printf("FAIL: a phase-fan term placement was never exercised\n");
        return 1;
    }
    std::printf("warp-local stage transitions checked: %ld\n", g_warp_local);
    if (g_warp_local == 0) {
        std::printf("FAIL: no warp-local stage transition was planned\n");
        return 1;
    }
    std::printf("ALL OK\n");
    return 0;
}
