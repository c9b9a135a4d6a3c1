// reads segments of gates ("G target ctrlmask m0..m7" lines, "F" = flush) and plans each segment with the engine's defaults
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "planner.h"
using namespace qgb;
static long g_low = 0, g_hi = 0, g_hil = 0;
int main(int argc, char **argv) {
    const int n = atoi(argv[1]);
    const bool verbose = argc > 3;
    FILE *f = fopen(argv[2], "r");
    PlanConfig cfg; cfg.T = 10; cfg.L = 4; cfg.K = 4; cfg.max_ops = 32; cfg.max_cost = 24; cfg.shear = true; cfg.row_lanes = 3; cfg.max_groups = QGB_MAX_GROUPS;
    static PassProgram<double> prog;
    std::vector<Gate> queue;
    char line[1024];
    long total_passes = 0, total_gates = 0, seg = 0, submitted = 0;
    auto flush = [&]() {
        long passes = 0, ops = 0, dense = 0;
        std::vector<int> per_pass;
        const size_t q0 = queue.size();
        while (!queue.empty()) {
            PlanStats st; plan_pass<double>(queue, n, cfg, prog, st);
            ++passes; ops += prog.n_ops; dense += st.shear_ops + st.direct_ops;
            per_pass.push_back(st.shear_ops + st.direct_ops);
            if (verbose) {
                int cost = 0, swp = 0, dg = 0; unsigned long long targets = 0;
                for (int o = 0; o < prog.n_ops; ++o) {
                    const auto &op = prog.op[o];
                    if (op.kind == OP_SHEAR) cost += 3; else if (op.kind == OP_GEN) cost += 4; else cost += 1;
                    swp += op.kind == OP_SWAP; dg += op.kind == OP_DIAG || op.kind == OP_DIAG_OUT;
                }
                for (int s2 = 0; s2 < prog.n_stages; ++s2) {
                    const Stage &sg = prog.stage[s2];
                    for (int o = sg.op_begin; o < sg.op_end; ++o) {
                        const auto &op = prog.op[o];
                        if (op.kind == OP_SHEAR || op.kind == OP_GEN || op.kind == OP_SWAP) targets |= 1ull << prog.tile_lane[sg.R[op.bit]];
                    }
                }
                int hi = __builtin_popcountll(targets >> 4); int lowops = 0, hiops = 0; for (int s2 = 0; s2 < prog.n_stages; ++s2) { const Stage &sg = prog.stage[s2]; for (int o = sg.op_begin; o < sg.op_end; ++o) { const auto &op = prog.op[o]; if (op.kind == OP_SHEAR || op.kind == OP_GEN) { if (prog.tile_lane[sg.R[op.bit]] < 4) ++lowops; else ++hiops; } } } g_low += lowops; g_hi += hiops; g_hil += hi;
                printf("   pass %ld: dense %d swap %d diag %d cost %d high target lanes %d stages %d queue left %zu\n", passes, st.shear_ops + st.direct_ops, swp, dg, cost, hi, prog.n_stages, queue.size());
            }
        }
        printf("segment %ld: submitted %ld queued %zu passes %ld dense ops %ld (%.2f per pass)", seg, submitted, q0, passes, dense, passes ? (double)dense / passes : 0.);
        if (verbose) { printf("  ["); for (int v : per_pass) printf("%d ", v); printf("]"); }
        printf("\n");
        total_passes += passes; total_gates += submitted; ++seg; submitted = 0;
    };
    while (fgets(line, sizeof line, f)) {
        if (line[0] == 'F') { flush(); continue; }
        if (line[0] != 'G') continue;
        Gate g; unsigned long long cm; int t;
        sscanf(line + 1, "%d %llu %lf %lf %lf %lf %lf %lf %lf %lf", &t, &cm, &g.m[0], &g.m[1], &g.m[2], &g.m[3], &g.m[4], &g.m[5], &g.m[6], &g.m[7]);
        g.target = t; g.ctrl_mask = cm;
        enqueue_gate(queue, g, true);
        ++submitted;
    }
    if (!queue.empty()) flush();
    printf("dense ops on low lanes %ld, on high lanes %ld, high target lanes %ld\n", g_low, g_hi, g_hil);
    printf("TOTAL passes %ld for %ld submitted gates in %ld segments\n", total_passes, total_gates, seg);
}
