/*
 * program.h — the "pass program" shared by the host planner and the device kernels.
 *
 * A deferred run of gates on one state vector is cut into PASSES.  One pass is one
 * read + one write of the state vector (2 * 2^n * sizeof(complex) algorithmic bytes):
 * every CTA stages a TILE of 2^T amplitudes in shared memory, applies the pass's gates
 * to it and writes it back.  The tile is the sub-cube spanned by T "tile lanes": the low
 * L lanes (contiguous in memory -> coalesced / bulk transfers) plus T-L arbitrary lanes.
 * Inside a pass the gates are grouped into STAGES; in a stage each thread keeps 2^K
 * amplitudes (the K "register bits" of the stage) in registers and applies the stage's
 * gates without touching shared memory.
 *
 * This header is plain C++ (no CUDA types) so tests can compile the planner alone.
 * Reference semantics being reproduced: one applyGate / applyControlledGate per entry,
 * qgate/simulator/src/CPUQubitProcessor.cpp:307-362.
 */
#pragma once
#include <stdint.h>

#include <vector>

namespace qgb {

enum OpKind : int {
    OP_GEN = 0,      /* full 2x2 on register bit `bit` (index into the stage's R[])          */
    OP_DIAG = 1,     /* multiply by d1 where the target bit is 1, by d0 elsewhere:           */
                     /* m[0..1] = d0, m[2..3] = d1.  A phase gate (d0 == 1) is encoded with  */
                     /* the target folded into the controls and d0 = d1 = the phase.         */
    OP_DIAG_OUT = 3, /* diag on a lane OUTSIDE the tile (`bit` = state-vector lane): the CTA  */
                     /* picks d0 or d1 from its base index                                    */
    OP_SWAP = 5,     /* exchange the two amplitudes of register bit `bit` (X, CX, CCX ...)    */
    OP_FAN = 7,      /* phase fan: a product of controlled-phase gates that share one lane (the */
                     /* hub): amplitudes whose hub bit is 1 are multiplied by the product of    */
                     /* e^{i phi_j} over the fan's lanes j that are 1.  The hub is lowered as a  */
                     /* control (cmt / regmask / ctrl_out); the lanes j are PassProgram::fan_term */
                     /* entries, sorted by where lane j lives in the op's stage: thread bits (one */
                     /* factor per thread, from two small tables the kernel builds once per CTA), */
                     /* lanes outside the tile (one factor per tile, from two tables in device    */
                     /* memory built before the pass), register bits (Fan::reg: one factor per     */
                     /* register bit, applied along a Gray-code walk over the registers).  A QFT's */
                     /* n - 1 - i controlled phases onto lane i are ONE op (Gate::fan).            */
    OP_SHEAR = 6,    /* 2x2 on register bit `bit` as THREE complex shears, in place:           */
                     /*   q0 += y q1;  q1 += g q0;  q0 += x q1        (m = y, g, x, 0)          */
                     /* = N q for the unit-determinant matrix N = [[1,x],[0,1]] [[1,0],[g,1]]   */
                     /* [[1,y],[0,1]].  The planner writes a gate matrix M as phi * P^s * N     */
                     /* (P = exchange of the two outputs): phi goes into a later gate or a      */
                     /* diagonal op, P is a RELABELLING of the stage's registers (Stage::      */
                     /* store_flip) — 12 FMAs per pair instead of 16 and no temporaries.       */
};

/* Op::arm bits: GEN on register bit j = 1 << j, SWAP on register bit j = 16 << j */
#define ARM_GEN(j) (1u << (j))
#define ARM_SWAP(j) (16u << (j))
#define ARM_DIAG_REG 256u   /* diagonal, target on a register bit                         */
#define ARM_DIAG_THR 512u   /* diagonal / phase, target on a thread bit or outside the tile */
#define ARM_MUX_THR 1024u   /* OP_GEN multiplexed by a thread bit (matrix chosen per thread)  */
#define ARM_MUX_REG 2048u   /* OP_GEN multiplexed by a register bit (matrix chosen per pair)  */
#define ARM_MUX_OUT 4096u   /* OP_GEN multiplexed by a lane outside the tile (per CTA)        */

#define QGB_MAX_TILE_LANES 14
#define QGB_MAX_REG_BITS 4
#define QGB_MAX_LANES 40
#define QGB_MAX_STAGES 40
#define QGB_MAX_OPS 48      /* (the fused pass kernel keeps one predicate bit per op: it runs at most 32) */
#define QGB_MAX_FANS 12     /* phase fans per pass                                                     */
#define QGB_MAX_FAN_TERMS 384
#define QGB_MAX_GROUPS 4    /* tensor-map dimensions above the 128-byte row (TMA staging)    */

/* One op of a stage.  The control predicate of an amplitude with tile index e = ebase | roff(r)
 * (thread part | register part) is split on the host: `cmt` is tested once per thread,
 * `regmask` once per register index and is the same for every thread. */
template <typename real>
struct Op {
    real m[8];            /* (re,im) of m00, m01, m10, m11 in the state precision          */
    int32_t kind;
    int32_t bit;
    uint32_t cmt;         /* controls on thread bits, tile-bit coordinates                 */
    uint32_t regmask;     /* bit r set: register index r satisfies the register-bit        */
                          /* controls (GEN / SWAP: tested for the pair's low index)        */
    uint32_t arm;         /* one-hot code path selector, see ARM_* (the kernel tests bits  */
                          /* instead of switching on kind/bit: no jump table, the op loop  */
                          /* stays on the uniform datapath)                                */
    uint32_t tsel;        /* OP_DIAG: tile-bit mask of the target lanes that are thread    */
                          /* bits: the thread takes d1 where an ODD number of them is set  */
                          /* OP_GEN + ARM_MUX_THR: tile-bit mask of the multiplexer lane   */
    uint32_t regsel;      /* OP_DIAG: register indices with an odd number of target bits   */
                          /* OP_GEN + ARM_MUX_REG: pairs (low index) whose mux bit is 1    */
    uint64_t ctrl_out;    /* controls outside the tile, state-vector index coordinates     */
    real m1[8];           /* multiplexed OP_GEN: the matrix where the multiplexer bit is 1 */
                          /* (ARM_MUX_OUT: `mux_out` = the state-vector lane, outside tile) */
    int32_t mux_out;
    int32_t code;         /* OPC_*: the specialised body kernels_tma.cu runs for this op          */
};

/* Op::code — one value per specialised op body of the TMA-staged kernel (switch on a uniform) */
#define OPC_GEN(j) (0 + (j))                      /* 2x2 on register bit j; matrix = m, or m1 where a  */
                                                  /* thread-bit / outside-tile multiplexer is 1        */
#define OPC_GEN_MASKED(j) (4 + (j))               /* the same under register-bit controls (regmask)    */
#define OPC_GEN_REGMUX(j, j2) (8 + 4 * (j) + (j2)) /* multiplexed by register bit j2: m / m1 per pair  */
#define OPC_SWAP(j) (24 + (j))
#define OPC_DIAG_REG 28
#define OPC_DIAG_THR 29
#define OPC_SHEAR(j) (32 + (j))                     /* three shears on register bit j (m, or m1    */
                                                    /* where a thread-bit / outside multiplexer is 1) */
#define OPC_SHEAR_MASKED(j) (36 + (j))              /* the same under register-bit controls (regmask) */
#define OPC_SHEAR_REGMUX(j, j2) (40 + 4 * (j) + (j2)) /* multiplexed by register bit j2               */
#define OPC_FAN 30                                  /* phase fan (Op::bit = index into PassProgram::fan) on the */
                                                    /* registers of regmask                                    */
#define OPC_FAN_ALL 31                              /* ... on every register (hub on a thread bit / outside)    */
#define OPC_FAN_REG(j, pol) (56 + 2 * (j) + (pol))  /* ... on the registers whose bit j equals pol (hub = a     */
                                                    /* register bit; pol = 0 after a relabelling of that bit)   */
#define OPC_COUNT 64

/* Shared-memory slot of tile element e: the 128-byte XOR swizzle TMA tensor maps produce
 * (byte address bits [6:4] ^= bits [9:7]).  Linear over GF(2), so
 * swizzle(a | b) == swizzle(a) ^ swizzle(b) for disjoint a, b. */
#if defined(__CUDACC__)
__host__ __device__
#endif
inline uint32_t tile_swizzle(uint32_t e, bool fp32) {
    return fp32 ? (e ^ (((e >> 4) & 7u) << 1)) : (e ^ ((e >> 3) & 7u));
}

struct Stage {
    int16_t op_begin, op_end;
    int8_t R[QGB_MAX_REG_BITS];           /* register bit j  <-> tile bit R[j], ascending   */
    int8_t W[QGB_MAX_TILE_LANES];         /* thread bit i    <-> tile bit W[i]              */
    int8_t warp_local;                    /* 1: the next stage reads only what the SAME warp   */
                                          /* wrote in this one (identical warp-index bits and   */
                                          /* the same union of lane + register bits): the TMA   */
                                          /* kernel then synchronises the warp, not the CTA     */
    uint16_t sro[1 << QGB_MAX_REG_BITS];  /* swizzled tile offset of register index r       */
    uint32_t xb[QGB_MAX_REG_BITS];        /* swizzled BYTE offset of register bit j: the slot of */
                                          /* register r is base ^ XOR of xb[j] over the bits of r */
    /* Register relabelling accumulated by the stage's OP_SHEAR ops (exchanged outputs): at the end
     * of the stage register r holds the tile element of register index r ^ store_flip, i.e. it is
     * stored at slot ^ store_xor (store_xor = XOR of xb[j] over the bits of store_flip).  The
     * planner has already expressed every later op of the stage in relabelled registers. */
    uint32_t store_flip;
    uint32_t store_xor;
};

template <typename real>
struct PassProgram {
    int32_t n_lanes;      /* lanes of the (local) state vector                             */
    int32_t T;            /* tile lanes                                                    */
    int32_t L;            /* low contiguous lanes inside the tile                          */
    int32_t K;            /* register bits per stage                                       */
    int32_t n_stages;
    int32_t n_ops;
    int8_t tile_lane[QGB_MAX_TILE_LANES];   /* tile bit p <-> lane, ascending, [p] = p for p < L */
    int8_t rest_lane[QGB_MAX_LANES];        /* the n_lanes - T other lanes, ascending        */
    /* TMA staging (kernels_tma.cu): the lanes above the 128-byte row (row_lanes of them) are cut,
     * ascending, into groups of [grp_t tile lanes][grp_r other lanes] starting at lane grp_start;
     * group d is dimension d + 1 of the pass's tensor map (box 2^grp_t of 2^(grp_t + grp_r)).
     * n_groups == 0: the tile shape does not fit a tensor map (cp.async staging only). */
    int32_t row_lanes;
    int32_t n_groups;
    int8_t grp_start[QGB_MAX_GROUPS], grp_t[QGB_MAX_GROUPS], grp_r[QGB_MAX_GROUPS];
    /* ops that look at lanes OUTSIDE the tile (controls, multiplexer, diagonal target): the
     * TMA-staged kernel turns them into two per-tile bit masks over the ops instead of testing
     * every op of every tile (needs n_ops <= 32) */
    int32_t n_direct;     /* OP_GEN ops (direct 2x2, the fallback of OP_SHEAR): 0 -> the leaner kernel variant */
    int32_t n_out;
    struct OutRef {
        uint64_t ctrl_mask;   /* the op runs only in tiles whose origin has these bits set       */
        uint64_t sel_mask;    /* the op takes m1 / d1 in tiles whose origin has an ODD number of */
                              /* these bits set (one bit: multiplexer or diagonal target outside */
                              /* the tile; several: a parity diagonal, Gate::parity)             */
        int16_t op;           /* op index                                                       */
        int16_t pad_[3];
    } out[QGB_MAX_OPS];
    Stage stage[QGB_MAX_STAGES];
    Op<real> op[QGB_MAX_OPS];
    /* phase fans (OP_FAN, Op::bit = fan number): fan f multiplies the amplitudes the op's control
     * predicate admits (hub bit 1) by the product of its terms whose bit is 1.  Terms
     * [first, first + n_thr) test bit `bit` of the THREAD number (thread bit i <-> tile bit W[i] of
     * the op's stage), the next n_out terms test state-vector lane `bit` of the tile's origin.  Terms
     * on register bits are folded into reg[b] = the factor of the registers whose PHYSICAL bit b is
     * set (the identity where the fan has no such term; a term on a bit the stage's shears relabelled
     * becomes base *= t, reg[b] = 1 / t) and `base`, a factor of every amplitude the fan touches (the
     * kernel folds it into its thread tables).  Factors are kept in double in both precisions. */
    int32_t n_fans;
    int32_t n_fan_terms;
    struct Fan {
        int16_t first, n_thr, n_out, n_reg; /* n_reg: register bits that carry a term (bit mask) */
        double reg[QGB_MAX_REG_BITS][2];
        double base[2];
    } fan[QGB_MAX_FANS];
    struct FanTerm {
        double re, im;
        int32_t bit;
        int32_t pad_;
    } fan_term[QGB_MAX_FAN_TERMS];
};

/* Cut the lanes [row_lanes, n) of tile-lane set S into tensor-map groups (see PassProgram).
 * Returns the number of groups (may exceed QGB_MAX_GROUPS: the arrays are filled up to that). */
inline int tile_groups(uint64_t S, int n, int row_lanes, int8_t *start, int8_t *t_bits, int8_t *r_bits) {
    int count = 0;
    auto emit = [&](int s, int t, int r) {
        if (count < QGB_MAX_GROUPS) {
            start[count] = (int8_t)s;
            t_bits[count] = (int8_t)t;
            r_bits[count] = (int8_t)r;
        }
        ++count;
    };
    int lane = row_lanes;
    while (lane < n) {
        int s = lane, t = 0, r = 0;
        while (lane < n && ((S >> lane) & 1ull)) ++t, ++lane;
        while (lane < n && !((S >> lane) & 1ull)) ++r, ++lane;
        while (t > 8) { /* box dimensions are capped at 256 */
            emit(s, 8, 0);
            s += 8;
            t -= 8;
        }
        while (t + r > 30) { /* global dimensions must stay below 2^32 */
            const int take = 30 - t;
            emit(s, t, take);
            s += t + take;
            r -= take;
            t = 0;
        }
        emit(s, t, r);
    }
    return count;
}

/* one queued gate, always kept in double (glue.cpp:382-405 builds the matrix in double and
 * the processor casts to the state precision at apply time, CPUQubitProcessor.cpp:312). */
struct Gate {
    double m[8];
    int32_t target;
    uint64_t ctrl_mask;
    /* multiplexed form, made by the host-side merging only (planner.cpp enqueue_gate): with
     * mux >= 0 the gate applies m where lane `mux` is 0 and m1 where it is 1 (ctrl_mask == 0).
     * U(t) . CX(c->t) . V(t) collapses into ONE such gate: m = U V, m1 = U X V. */
    int32_t mux = -1;
    double m1[8] = {0., 0., 0., 0., 0., 0., 0., 0.};
    /* parity diagonal (exp(i theta Z x Z x ... x Z)): with parity != 0 the gate
     * is diagonal over ALL lanes of `parity` (target = the lowest of them): amplitudes whose index
     * has an even number of those bits set are multiplied by (m[0], m[1]), the others by
     * (m[6], m[7]).  An ordinary diagonal gate is the case parity == 1 << target. */
    uint64_t parity = 0;
    /* phase fan (made by the host-side merging only, planner.cpp enqueue_gate): with a non-empty
     * `fan` the gate is the product over the terms of the controlled phase CP(target, lane; e^{i phi})
     * — amplitudes with bit `target` (the hub) AND bit `lane` set are multiplied by (re, im).
     * m is the identity, ctrl_mask 0. */
    struct FanTerm {
        int32_t lane;
        double re, im;
    };
    std::vector<FanTerm> fan;
};

inline uint64_t gate_fan_lanes(const Gate &g) {
    uint64_t lanes = 0;
    for (const Gate::FanTerm &t : g.fan) lanes |= 1ull << t.lane;
    return lanes;
}
inline uint64_t gate_diag_lanes(const Gate &g) {
    return (g.parity ? g.parity : (1ull << g.target)) | gate_fan_lanes(g);
}

inline bool gate_is_diag(const Gate &g) {
    if (!g.fan.empty()) return true;
    if (g.mux >= 0) return false; /* multiplexed gates are only formed when the result is dense */
    return g.m[2] == 0. && g.m[3] == 0. && g.m[4] == 0. && g.m[5] == 0.;
}
inline bool gate_is_antidiag(const Gate &g) {
    return g.m[0] == 0. && g.m[1] == 0. && g.m[6] == 0. && g.m[7] == 0.;
}

} // namespace qgb
