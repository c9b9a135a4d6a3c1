/*
 * planner_stats.cpp — offline view of what the flush planner does with the bench circuit
 * (random U3 + CX ladder, bench.py): merged gates, passes, ops per pass by kind.  CPU only;
 * used to tune planner options before spending GPU time.  Usage: planner_stats n depth T L K
 */
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "planner.h"

using namespace qgb;

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 30, depth = argc > 2 ? atoi(argv[2]) : 200;
    const int T = argc > 3 ? atoi(argv[3]) : 11, L = argc > 4 ? atoi(argv[4]) : 5, K = argc > 5 ? atoi(argv[5]) : 3;
    const int max_ops = argc > 6 ? atoi(argv[6]) : QGB_MAX_OPS, max_stages = argc > 7 ? atoi(argv[7]) : QGB_MAX_STAGES;
    const int max_cost = argc > 8 ? atoi(argv[8]) : (1 << 30), shear = argc > 9 ? atoi(argv[9]) : 0, tma = argc > 10 ? atoi(argv[10]) : 0;
    std::mt19937_64 rng(1234);
    std::uniform_real_distribution<double> u(0., 6.283185307179586);
    std::vector<Gate> queue;
    long submitted = 0, merged = 0;
    const bool qft = getenv("QGB_STATS_QFT") != nullptr; /* QFT instead of the bench circuit (depth ignored) */
    for (int i = 0; qft && i < n; ++i) {
        Gate h;
        for (int e = 0; e < 8; ++e) h.m[e] = 0;
        const double r = std::sqrt(0.5);
        h.m[0] = h.m[2] = h.m[4] = r, h.m[6] = -r;
        h.target = i, h.ctrl_mask = 0;
        merged += enqueue_gate(queue, h, true);
        ++submitted;
        for (int j = i + 1; j < n; ++j) {
            Gate g;
            for (int e = 0; e < 8; ++e) g.m[e] = 0;
            const double phi = 3.141592653589793 / std::ldexp(1., j - i);
            g.m[0] = 1, g.m[6] = std::cos(phi), g.m[7] = std::sin(phi);
            g.target = i, g.ctrl_mask = 1ull << j;
            merged += enqueue_gate(queue, g, true);
            ++submitted;
        }
    }
    for (int d = 0; !qft && d < depth; ++d) {
        for (int i = 0; i < n; ++i) {
            const double th = u(rng), ph = u(rng), la = u(rng);
            Gate g;
            g.target = i;
            g.ctrl_mask = 0;
            const double c = std::cos(th / 2), s = std::sin(th / 2);
            g.m[0] = c; g.m[1] = 0;
            g.m[2] = -std::cos(la) * s; g.m[3] = -std::sin(la) * s;
            g.m[4] = std::cos(ph) * s; g.m[5] = std::sin(ph) * s;
            g.m[6] = std::cos(ph + la) * c; g.m[7] = std::sin(ph + la) * c;
            merged += enqueue_gate(queue, g, true);
            ++submitted;
        }
        for (int i = d % 2; i < n - 1; i += 2) {
            Gate g;
            for (int e = 0; e < 8; ++e) g.m[e] = 0;
            g.m[2] = 1; g.m[4] = 1;
            g.target = i + 1;
            g.ctrl_mask = 1ull << i;
            merged += enqueue_gate(queue, g, true);
            ++submitted;
        }
    }
    std::printf("submitted %ld merged %ld queued %zu\n", submitted, merged, queue.size());
    PlanConfig cfg;
    cfg.T = T; cfg.L = L; cfg.K = K; cfg.fp32 = false; cfg.max_ops = max_ops; cfg.max_stages = max_stages;
    cfg.max_cost = max_cost; cfg.shear = shear != 0;
    if (tma) { cfg.row_lanes = 3; cfg.max_groups = QGB_MAX_GROUPS; cfg.L = std::max(cfg.L, 3); }
    static PassProgram<double> prog;
    long fans = 0, fan_terms = 0;
    long passes = 0, ops = 0, stages = 0, gen = 0, swp = 0, diag = 0, mthr = 0, mreg = 0, mout = 0, cthr = 0, shr = 0, resid = 0, conflict = 0;
    double worst = 0.;
    while (!queue.empty()) {
        PlanStats st;
        plan_pass<double>(queue, n, cfg, prog, st);
        ++passes;
        ops += prog.n_ops;
        fans += prog.n_fans, fan_terms += prog.n_fan_terms;
        resid += st.residual_ops;
        for (int s = 0; s < prog.n_stages; ++s) {
            const Stage &sg = prog.stage[s];
            if (sg.op_end == sg.op_begin) continue;
            /* complex128: the 8 lanes of a quarter warp (thread bits 0..2) must differ in the three
             * bank-selecting slot bits, which tile bits b and b + 3 (b < 3) both move */
            int cls = 0;
            for (int i = 0; i < 3 && i < prog.T - prog.K; ++i)
                if (sg.W[i] < 6) cls |= 1 << (sg.W[i] % 3);
            conflict += cls != 7;
        }
        for (int s = 0; s < prog.n_stages; ++s) stages += prog.stage[s].op_end > prog.stage[s].op_begin;
        for (int o = 0; o < prog.n_ops; ++o) {
            const Op<double> &op = prog.op[o];
            gen += op.kind == OP_GEN;
            shr += op.kind == OP_SHEAR;
            if (op.kind == OP_SHEAR)
                for (int e = 0; e < 6; e += 2) {
                    worst = std::max(worst, std::hypot((double)op.m[e], (double)op.m[e + 1]));
                    worst = std::max(worst, std::hypot((double)op.m1[e], (double)op.m1[e + 1]));
                }
            swp += op.kind == OP_SWAP;
            diag += op.kind == OP_DIAG || op.kind == OP_DIAG_OUT;
            mthr += (op.arm & ARM_MUX_THR) != 0;
            mreg += (op.arm & ARM_MUX_REG) != 0;
            mout += (op.arm & ARM_MUX_OUT) != 0;
            cthr += op.cmt != 0;
        }
    }
    std::printf("shear %ld (largest coefficient %.2f)  residual phase ops %ld  stages with a 2-way bank conflict %ld\n", shr, worst, resid, conflict);
    std::printf("passes %ld  ops/pass %.1f  stages/pass %.1f  submitted gates/pass %.1f\n", passes, (double)ops / passes,
                (double)stages / passes, (double)submitted / passes);
    std::printf("ops: gen %ld (mux thread %ld, register %ld, outside %ld)  swap %ld  diag %ld  thread-controlled %ld\n", gen,
                mthr, mreg, mout, swp, diag, cthr);
    std::printf("phase fans %ld with %ld terms\n", fans, fan_terms);
    return 0;
}
