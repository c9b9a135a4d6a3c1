/*
 * planner.h — host-side flush planner: deferred gate queue -> pass programs.
 * Pure C++ (no CUDA); tests/native/planner_emul.cpp compiles it on the CPU box.
 */
#pragma once
#include <vector>

#include "program.h"

namespace qgb {

struct PlanConfig {
    int T = 11;          /* tile lanes (clamped to n_lanes)                               */
    int L = 5;           /* low contiguous lanes forced into every tile                   */
    int K = 3;           /* register bits per stage                                       */
    int max_ops = QGB_MAX_OPS;
    int max_stages = QGB_MAX_STAGES;
    int lookahead = 4096; /* how far past the first blocked gate the planner searches      */
    bool fp32 = false;    /* selects the shared-memory bank classes used to order W[]      */
    int max_cost = 1 << 30; /* cap on the summed op cost of a pass (see op_cost)            */
    /* TMA staging: lanes of one 128-byte shared-memory row (3 complex128 / 4 complex64) and the
     * most tensor-map groups a tile may need (0: any tile shape, cp.async staging) */
    int row_lanes = 0;
    int max_groups = 0;
    /* give runs of consecutive stages the same warp-index bits (tile bits none of them uses as a
     * register bit) so that a stage transition needs a warp barrier only (Stage::warp_local) */
    bool warp_local = false;
    /* lower dense 2x2 gates as three in-place complex shears (OP_SHEAR, program.h) where the
     * factorisation is well conditioned (every coefficient <= shear_max_coef in magnitude);
     * the rest stays a direct 2x2 (OP_GEN) */
    bool shear = false;
    double shear_max_coef = 12.;
    int fan_cost = 2;     /* cost of a phase fan against max_cost (sheared 2x2 = 3, diagonal = 1) */
};

struct PlanStats {
    int gates_in_pass = 0;
    int ops_in_pass = 0;
    int stages_in_pass = 0;
    int shear_ops = 0;    /* dense gates lowered as OP_SHEAR                                  */
    int direct_ops = 0;   /* dense gates lowered as OP_GEN (direct 2x2)                       */
    int residual_ops = 0; /* diagonal ops added for phases no later gate could absorb          */
    int fan_ops = 0;      /* phase fans (OP_FAN)                                               */
};

/* merge `g` into the queue: an uncontrolled gate folds into the previous gate on the same
 * lane when nothing in between touches that lane; a controlled gate folds into an
 * identical-signature gate directly before it; with `fans`, controlled phases that share a lane
 * collect into one phase fan (Gate::fan).  Returns true when merged (queue size unchanged). */
bool enqueue_gate(std::vector<Gate> &queue, const Gate &g, bool merge, bool fans = true);

/* Plan ONE pass from the front of `queue` for a state vector of n_lanes lanes.  Executed
 * gates are removed from `queue` (order of the remaining gates is preserved).  Always
 * consumes at least one gate when the queue is not empty. */
template <typename real>
void plan_pass(std::vector<Gate> &queue, int n_lanes, const PlanConfig &cfg,
               PassProgram<real> &prog, PlanStats &stats);

} // namespace qgb
