/*
 * ref_shim.cpp — TEST INFRASTRUCTURE, not product code.
 *
 * Puts the UNMODIFIED reference CPU runtime (compiled in place from
 * /root/reference/qgate/simulator/src by oracle/Makefile) behind the C ABI of
 * include/qgate_b200.h, so the same Python host code and the same tests can
 * drive either the reference or the CUDA engine.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load
 * the resulting oracle/_ref/libqgate_ref_cpu.so.
 *
 * This file plays the role of the reference's glue.cpp + cpuext.cpp (argument
 * unpacking, bounds checks, the all-empty get_states shortcut) without any
 * CPython dependency.  It contains no reference source; it only #includes the
 * reference's public headers at build time.
 *   Interfaces.h:7-81          the 4 abstract classes called below
 *   cpuext.cpp:9-70            which concrete classes are instantiated per dtype
 *   glue.cpp:447-506           get_states range checks + empty-qstates shortcut
 */
#include "qgate_b200.h"

#include "Interfaces.h"
#include "GateMatrix.h"
#include "Misc.h"
#include "CPUQubitStates.h"
#include "CPUQubitProcessor.h"
#include "CPUQubitsStatesGetter.h"
#include "CPUSamplingPool.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <map>
#include <string>
#include <vector>

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

#define QGB_TRY try {
#define QGB_CATCH                                                   \
    } catch (const std::exception &e) {                             \
        return fail(QGB_ERR_RUNTIME, e.what());                     \
    } catch (...) {                                                 \
        return fail(QGB_ERR_RUNTIME, "unknown C++ exception");      \
    }                                                               \
    return QGB_OK;

qgate::QubitStates *QS(qgb_handle h) { return reinterpret_cast<qgate::QubitStates *>(h); }
qgate::QubitProcessor *QP(qgb_handle h) { return reinterpret_cast<qgate::QubitProcessor *>(h); }
qgate::QubitsStatesGetter *QG(qgb_handle h) { return reinterpret_cast<qgate::QubitsStatesGetter *>(h); }
qgate::SamplingPool *SP(qgb_handle h) { return reinterpret_cast<qgate::SamplingPool *>(h); }

qgate::QubitStatesList toList(const qgb_handle *hs, int n) {
    qgate::QubitStatesList list;
    for (int i = 0; i < n; ++i)
        list.push_back(QS(hs[i]));
    return list;
}

qgate::IdListList toTables(const int *lane_tables, const int *n_per, int n_qstates) {
    qgate::IdListList tables;
    const int *p = lane_tables;
    for (int i = 0; i < n_qstates; ++i) {
        tables.emplace_back(p, p + n_per[i]);
        p += n_per[i];
    }
    return tables;
}

void toMatrix(const double *m, qgate::Matrix2x2C64 &mat) {
    mat(0, 0) = std::complex<double>(m[0], m[1]);
    mat(0, 1) = std::complex<double>(m[2], m[3]);
    mat(1, 0) = std::complex<double>(m[4], m[5]);
    mat(1, 1) = std::complex<double>(m[6], m[7]);
}

int buildMatrix(int gate_id, const double *a, int n_args, int adjoint, qgate::Matrix2x2C64 &mat) {
    static const int nargs[QGB_N_GATES] = {3, 2, 1, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 0};
    if (gate_id < 0 || gate_id >= QGB_N_GATES)
        return fail(QGB_ERR_RUNTIME, "Unknown gate type.");
    if (n_args != nargs[gate_id])
        return fail(QGB_ERR_INVALID, "wrong number of gate arguments.");
    switch (gate_id) {
    case QGB_GATE_U:     qgate::U_mat(mat, a[0], a[1], a[2]); break;
    case QGB_GATE_U2:    qgate::U2_mat(mat, a[0], a[1]); break;
    case QGB_GATE_U1:    qgate::U1_mat(mat, a[0]); break;
    case QGB_GATE_ID:    qgate::ID_mat(mat); break;
    case QGB_GATE_X:     qgate::X_mat(mat); break;
    case QGB_GATE_Y:     qgate::Y_mat(mat); break;
    case QGB_GATE_Z:     qgate::Z_mat(mat); break;
    case QGB_GATE_H:     qgate::H_mat(mat); break;
    case QGB_GATE_S:     qgate::S_mat(mat); break;
    case QGB_GATE_T:     qgate::T_mat(mat); break;
    case QGB_GATE_RX:    qgate::RX_mat(mat, a[0]); break;
    case QGB_GATE_RY:    qgate::RY_mat(mat, a[0]); break;
    case QGB_GATE_RZ:    qgate::RZ_mat(mat, a[0]); break;
    case QGB_GATE_EXPII: qgate::ExpiI_mat(mat, a[0]); break;
    case QGB_GATE_EXPIZ: qgate::ExpiZ_mat(mat, a[0]); break;
    case QGB_GATE_SH:    qgate::SH_mat(mat); break;
    }
    if (adjoint)
        qgate::adjoint(&mat);
    return QGB_OK;
}

qgb_stats g_stats;

} // namespace

/* ---- sharding support: what qgate_b200/dist.py needs from a local runtime.  The reference has
 * no such entry points; these few loops are this shim's own (they let the world_size-2 gloo
 * tests drive the sharding host logic over the reference's gate / reduce / readout code). */

namespace {

template <class real>
bool dataPtr(qgate::QubitStates *qs, void **ptr, int64_t *bytes) {
    qgate_cpu::CPUQubitStates<real> *c = dynamic_cast<qgate_cpu::CPUQubitStates<real> *>(qs);
    if (c == NULL)
        return false;
    *ptr = c->getPtr();
    *bytes = (int64_t)(sizeof(real) * 2) << c->getNLanes();
    return true;
}

bool anyDataPtr(qgb_handle h, void **ptr, int64_t *bytes) {
    return dataPtr<double>(QS(h), ptr, bytes) || dataPtr<float>(QS(h), ptr, bytes);
}

std::map<qgb_handle, void *> g_alt;

template <class real>
void joinShard(std::complex<real> *dst, int n_dst_lanes, const qgb_handle *src, int n_src,
               int n_product_lanes, int64_t offset) {
    std::vector<const std::complex<real> *> ptr(n_src);
    std::vector<int> shift(n_src), lanes(n_src);
    int sh = 0;
    for (int k = n_src - 1; k >= 0; --k) {
        void *p; int64_t bytes;
        anyDataPtr(src[k], &p, &bytes);
        ptr[k] = static_cast<const std::complex<real> *>(p);
        lanes[k] = QS(src[k])->getNLanes();
        shift[k] = sh;
        sh += lanes[k];
    }
    for (int64_t d = 0; d < (1LL << n_dst_lanes); ++d) {
        int64_t i = offset + d;
        std::complex<real> v(0, 0);
        if (i < (1LL << n_product_lanes)) {
            int last = n_src - 1;
            v = ptr[last][(i >> shift[last]) & ((1LL << lanes[last]) - 1)];
            for (int k = last - 1; k >= 0; --k)
                v = ptr[k][(i >> shift[k]) & ((1LL << lanes[k]) - 1)] * v;
        }
        dst[d] = v;
    }
}

/* cumulative array in double + upper_bound, for pools made in two steps or from a vector */
struct ShimPool : qgate::SamplingPool {
    std::vector<double> cum;
    std::vector<int> empty;      /* ascending */
    bool fp32, finalized;
    void sample(qgate::QstateIdx *obs, int n, const double *r) {
        for (int i = 0; i < n; ++i) {
            double v = fp32 ? (double)(float)r[i] : r[i];
            int64_t idx = std::upper_bound(cum.begin(), cum.end(), v) - cum.begin();
            if (idx > (int64_t)cum.size() - 1) idx = (int64_t)cum.size() - 1;
            for (size_t k = 0; k < empty.size(); ++k) {
                int64_t lo = idx & ((1LL << empty[k]) - 1);
                idx = ((idx - lo) << 1) | lo;
            }
            obs[i] = idx;
        }
    }
};

} // namespace

extern "C" {

const char *qgb_last_error(void) { return g_last_error.c_str(); }
const char *qgb_backend_name(void) { return "reference-cpu"; }
int qgb_abi_version(void) { return 1; }

int qgb_devices_initialize(const int *, int, int, int64_t) { return QGB_OK; }
int qgb_devices_clear(void) { return QGB_OK; }
int qgb_device_count(int *count) { *count = 0; return QGB_OK; }
int qgb_set_stream(uint64_t) { return QGB_OK; }

int qgb_gate_matrix(int gate_id, const double *args, int n_args, int adjoint, double *mat8) {
    QGB_TRY
    qgate::Matrix2x2C64 mat;
    int rc = buildMatrix(gate_id, args, n_args, adjoint, mat);
    if (rc != QGB_OK)
        return rc;
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 2; ++c) {
            mat8[(r * 2 + c) * 2] = mat(r, c).real();
            mat8[(r * 2 + c) * 2 + 1] = mat(r, c).imag();
        }
    QGB_CATCH
}

int qgb_qstates_new(int prec, qgb_handle *out) {
    QGB_TRY
    qgate::QubitStates *qs = NULL;
    if (prec == QGB_PREC_FP64)
        qs = new qgate_cpu::CPUQubitStates<double>();
    else if (prec == QGB_PREC_FP32)
        qs = new qgate_cpu::CPUQubitStates<float>();
    else
        return fail(QGB_ERR_INVALID, "unknown precision.");
    *out = reinterpret_cast<qgb_handle>(qs);
    QGB_CATCH
}

int qgb_qstates_delete(qgb_handle h) {
    QGB_TRY
    if (g_alt.count(h)) {
        free(g_alt[h]);
        g_alt.erase(h);
    }
    delete QS(h);
    QGB_CATCH
}

int qgb_qstates_deallocate(qgb_handle h) {
    QGB_TRY
    QS(h)->deallocate();
    QGB_CATCH
}

int qgb_qstates_get_n_lanes(qgb_handle h, int *n_lanes) {
    QGB_TRY
    *n_lanes = QS(h)->getNLanes();
    QGB_CATCH
}

int qgb_qproc_new(int prec, qgb_handle *out) {
    QGB_TRY
    qgate::QubitProcessor *qp = NULL;
    if (prec == QGB_PREC_FP64)
        qp = new qgate_cpu::CPUQubitProcessor<double>();
    else if (prec == QGB_PREC_FP32)
        qp = new qgate_cpu::CPUQubitProcessor<float>();
    else
        return fail(QGB_ERR_INVALID, "unknown precision.");
    *out = reinterpret_cast<qgb_handle>(qp);
    QGB_CATCH
}

int qgb_qproc_delete(qgb_handle h) {
    QGB_TRY
    delete QP(h);
    QGB_CATCH
}

int qgb_qproc_synchronize(qgb_handle h) {
    QGB_TRY
    QP(h)->synchronize();
    QGB_CATCH
}

int qgb_qproc_reset(qgb_handle h) {
    QGB_TRY
    QP(h)->reset();
    QGB_CATCH
}

int qgb_qproc_initialize_qstates(qgb_handle qp, qgb_handle qs, int n_lanes) {
    QGB_TRY
    QP(qp)->initializeQubitStates(*QS(qs), n_lanes);
    QGB_CATCH
}

int qgb_qproc_reset_qstates(qgb_handle qp, qgb_handle qs) {
    QGB_TRY
    QP(qp)->resetQubitStates(*QS(qs));
    QGB_CATCH
}

int qgb_qproc_calc_probability(qgb_handle qp, qgb_handle qs, int lane, double *prob) {
    QGB_TRY
    *prob = QP(qp)->calcProbability(*QS(qs), lane);
    QGB_CATCH
}

int qgb_qproc_join(qgb_handle qp, qgb_handle dst, const qgb_handle *src, int n_src, int n_new) {
    QGB_TRY
    QP(qp)->join(*QS(dst), toList(src, n_src), n_new);
    QGB_CATCH
}

int qgb_qproc_decohere(qgb_handle qp, int value, double prob, qgb_handle qs, int lane) {
    QGB_TRY
    QP(qp)->decohere(value, prob, *QS(qs), lane);
    QGB_CATCH
}

int qgb_qproc_decohere_and_separate(qgb_handle qp, int value, double prob, qgb_handle qs0,
                                    qgb_handle qs1, qgb_handle qs, int lane) {
    QGB_TRY
    QP(qp)->decohereAndSeparate(value, prob, *QS(qs0), *QS(qs1), *QS(qs), lane);
    QGB_CATCH
}

int qgb_qproc_apply_reset(qgb_handle qp, qgb_handle qs, int lane) {
    QGB_TRY
    QP(qp)->applyReset(*QS(qs), lane);
    QGB_CATCH
}

int qgb_qproc_apply_gate(qgb_handle qp, const double *mat8, qgb_handle qs, int lane) {
    QGB_TRY
    qgate::Matrix2x2C64 mat;
    toMatrix(mat8, mat);
    QP(qp)->applyGate(mat, *QS(qs), lane);
    g_stats.gates_submitted += 1;
    g_stats.gate_amp_updates += 1LL << QS(qs)->getNLanes();
    QGB_CATCH
}

int qgb_qproc_apply_controlled_gate(qgb_handle qp, const double *mat8, qgb_handle qs,
                                    const int *ctrl, int n_ctrl, int target) {
    QGB_TRY
    qgate::Matrix2x2C64 mat;
    toMatrix(mat8, mat);
    qgate::IdList controls(ctrl, ctrl + n_ctrl);
    QP(qp)->applyControlledGate(mat, *QS(qs), controls, target);
    g_stats.gates_submitted += 1;
    g_stats.gate_amp_updates += 1LL << (QS(qs)->getNLanes() - n_ctrl);
    QGB_CATCH
}

int qgb_qproc_apply_gate_typed(qgb_handle qp, int gate_id, const double *args, int n_args,
                               int adjoint, qgb_handle qs, const int *ctrl, int n_ctrl,
                               int target) {
    QGB_TRY
    qgate::Matrix2x2C64 mat;
    int rc = buildMatrix(gate_id, args, n_args, adjoint, mat);
    if (rc != QGB_OK)
        return rc;
    if (n_ctrl == 0) {
        QP(qp)->applyGate(mat, *QS(qs), target);
    } else {
        qgate::IdList controls(ctrl, ctrl + n_ctrl);
        QP(qp)->applyControlledGate(mat, *QS(qs), controls, target);
    }
    g_stats.gates_submitted += 1;
    g_stats.gate_amp_updates += 1LL << (QS(qs)->getNLanes() - n_ctrl);
    QGB_CATCH
}

/* The "next" rows on the reference runtime = what the reference's front end
 * expands them into (qgate/model/expand.py:14-90), applied gate by gate. */
int qgb_qproc_apply_swap(qgb_handle qp, qgb_handle qs, int lane_a, int lane_b) {
    const int a[1] = {lane_a}, b[1] = {lane_b};
    int rc = qgb_qproc_apply_gate_typed(qp, 4 /* X */, NULL, 0, 0, qs, b, 1, lane_a);
    if (rc == QGB_OK) rc = qgb_qproc_apply_gate_typed(qp, 4, NULL, 0, 0, qs, a, 1, lane_b);
    if (rc == QGB_OK) rc = qgb_qproc_apply_gate_typed(qp, 4, NULL, 0, 0, qs, b, 1, lane_a);
    return rc;
}

int qgb_qproc_apply_pauli_expi(qgb_handle qp, qgb_handle qs, double theta, const int *lanes, const int *paulis,
                               int n, const int *ctrl, int n_ctrl) {
    /* basis changes, a CX chain onto the last non-identity factor, ExpiZ there, everything undone */
    std::vector<int> active;
    for (int i = 0; i < n; ++i)
        if (paulis[i] != 0) active.push_back(i);
    int rc = QGB_OK;
    for (size_t k = 0; k < active.size() && rc == QGB_OK; ++k) {
        const int i = active[k];
        if (paulis[i] == 2) rc = qgb_qproc_apply_gate_typed(qp, 8 /* S */, NULL, 0, 1, qs, NULL, 0, lanes[i]);
        if (rc == QGB_OK && paulis[i] != 3) rc = qgb_qproc_apply_gate_typed(qp, 7 /* H */, NULL, 0, 0, qs, NULL, 0, lanes[i]);
    }
    if (active.empty()) {
        int free_lane = 0;
        for (;; ++free_lane) {
            bool used = false;
            for (int c = 0; c < n_ctrl; ++c) used = used || ctrl[c] == free_lane;
            if (!used) break;
        }
        return qgb_qproc_apply_gate_typed(qp, 13 /* ExpiI */, &theta, 1, 0, qs, ctrl, n_ctrl, free_lane);
    }
    const int last = lanes[active.back()];
    for (size_t k = 0; k + 1 < active.size() && rc == QGB_OK; ++k) {
        const int c[1] = {lanes[active[k]]};
        rc = qgb_qproc_apply_gate_typed(qp, 4 /* X */, NULL, 0, 0, qs, c, 1, last);
    }
    if (rc == QGB_OK) rc = qgb_qproc_apply_gate_typed(qp, 14 /* ExpiZ */, &theta, 1, 0, qs, ctrl, n_ctrl, last);
    for (size_t k = active.size() - 1; k-- > 0 && rc == QGB_OK;) {
        const int c[1] = {lanes[active[k]]};
        rc = qgb_qproc_apply_gate_typed(qp, 4, NULL, 0, 0, qs, c, 1, last);
    }
    for (size_t k = active.size(); k-- > 0 && rc == QGB_OK;) {
        const int i = active[k];
        if (paulis[i] != 3) rc = qgb_qproc_apply_gate_typed(qp, 7, NULL, 0, 0, qs, NULL, 0, lanes[i]);
        if (rc == QGB_OK && paulis[i] == 2) rc = qgb_qproc_apply_gate_typed(qp, 8, NULL, 0, 0, qs, NULL, 0, lanes[i]);
    }
    return rc;
}

int qgb_qproc_apply_gates_batch(qgb_handle qp, qgb_handle qs, const qgb_gate_op *ops, int64_t n_ops) {
    for (int64_t i = 0; i < n_ops; ++i) {
        const qgb_gate_op &op = ops[i];
        int ctrl[64], n_ctrl = 0;
        for (int lane = 0; lane < 64; ++lane)
            if ((op.ctrl_mask >> lane) & 1ull) ctrl[n_ctrl++] = lane;
        int rc;
        if (op.gate_id == QGB_GATE_MATRIX)
            rc = n_ctrl ? qgb_qproc_apply_controlled_gate(qp, op.mat8, qs, ctrl, n_ctrl, op.target)
                        : qgb_qproc_apply_gate(qp, op.mat8, qs, op.target);
        else {
            static const int n_args[16] = {3, 2, 1, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 0};
            if (op.gate_id < 0 || op.gate_id > 15) return fail(QGB_ERR_RUNTIME, "Unknown gate type.");
            rc = qgb_qproc_apply_gate_typed(qp, op.gate_id, op.args, n_args[op.gate_id], op.adjoint, qs, ctrl, n_ctrl,
                                            op.target);
        }
        if (rc != QGB_OK) return rc;
    }
    return QGB_OK;
}

int qgb_getter_new(int prec, qgb_handle *out) {
    QGB_TRY
    qgate::QubitsStatesGetter *g = NULL;
    if (prec == QGB_PREC_FP64)
        g = new qgate_cpu::CPUQubitsStatesGetter<double>();
    else if (prec == QGB_PREC_FP32)
        g = new qgate_cpu::CPUQubitsStatesGetter<float>();
    else
        return fail(QGB_ERR_INVALID, "unknown precision.");
    *out = reinterpret_cast<qgb_handle>(g);
    QGB_CATCH
}

int qgb_getter_delete(qgb_handle h) {
    QGB_TRY
    delete QG(h);
    QGB_CATCH
}

int qgb_getter_get_states(qgb_handle getter, void *array, int64_t array_offset, int mathop,
                          const int *lane_tables, const int *n_per, int64_t empty_lane_mask,
                          const qgb_handle *qstates_list, int n_qstates, int n_ext_lanes,
                          int64_t n_states, int64_t start, int64_t step) {
    QGB_TRY
    /* range checks of glue.cpp:465-478 */
    int64_t space = 1LL << n_ext_lanes;
    if (start < 0 || space <= start)
        return fail(QGB_ERR_INVALID, "value out of range");
    int64_t end = start + step * (n_states - 1);
    if (end < 0 || space <= end)
        return fail(QGB_ERR_INVALID, "value out of range");
    if (mathop != QGB_MATHOP_NULL && mathop != QGB_MATHOP_PROB)
        return fail(QGB_ERR_RUNTIME, "unknown math op.");

    if (n_qstates == 0) {
        /* no qstates at all: |0...0> (glue.cpp:481-496).  The item size cannot be
         * derived from the (absent) qstates, so the getter's precision decides. */
        bool fp64 = dynamic_cast<qgate_cpu::CPUQubitsStatesGetter<double> *>(QG(getter)) != NULL;
        size_t real_size = fp64 ? sizeof(double) : sizeof(float);
        size_t item = (mathop == QGB_MATHOP_NULL) ? 2 * real_size : real_size;
        char *dst = static_cast<char *>(array) + array_offset * item;
        qgate::fillZeros(dst, (qgate::QstateSize)(n_states * item));
        if (start == 0 && n_states > 0) {
            if (fp64) *reinterpret_cast<double *>(dst) = 1.;
            else *reinterpret_cast<float *>(dst) = 1.f;
        }
        return QGB_OK;
    }
    qgate::IdListList tables = toTables(lane_tables, n_per, n_qstates);
    QG(getter)->getStates(array, array_offset, (qgate::MathOp)mathop, tables.data(),
                          empty_lane_mask, toList(qstates_list, n_qstates), n_states, start, step);
    g_stats.d2h_bytes += 0;
    QGB_CATCH
}

int qgb_getter_prepare_prob_array(qgb_handle getter, void *prob, const int *lane_tables,
                                  const int *n_per, const qgb_handle *qstates_list, int n_qstates,
                                  int n_lanes, int n_hidden) {
    QGB_TRY
    QG(getter)->prepareProbArray(prob, toTables(lane_tables, n_per, n_qstates),
                                 toList(qstates_list, n_qstates), n_lanes, n_hidden);
    QGB_CATCH
}

int qgb_getter_create_sampling_pool(qgb_handle getter, const int *lane_tables, const int *n_per,
                                    const qgb_handle *qstates_list, int n_qstates, int n_lanes,
                                    int n_hidden, const int *empty_lanes, int n_empty,
                                    qgb_handle *pool) {
    QGB_TRY
    qgate::IdList empties(empty_lanes, empty_lanes + n_empty);
    qgate::SamplingPool *sp =
        QG(getter)->createSamplingPool(toTables(lane_tables, n_per, n_qstates),
                                       toList(qstates_list, n_qstates), n_lanes, n_hidden, empties);
    *pool = reinterpret_cast<qgb_handle>(sp);
    QGB_CATCH
}

int qgb_pool_sample(qgb_handle pool, int64_t *obs, int n_samples, const double *randnum) {
    QGB_TRY
    static_assert(sizeof(qgate::QstateIdx) == sizeof(int64_t), "index width");
    SP(pool)->sample(reinterpret_cast<qgate::QstateIdx *>(obs), n_samples, randnum);
    QGB_CATCH
}

int qgb_pool_sample_sequential(qgb_handle, int64_t *, int, const double *) {
    return fail(QGB_ERR_RUNTIME, "the reference pool has no sequential sampler (its front end re-runs the circuit per shot).");
}

int qgb_pool_delete(qgb_handle pool) {
    QGB_TRY
    delete SP(pool);
    QGB_CATCH
}

int qgb_qstates_data_ptr(qgb_handle h, uint64_t *ptr, int64_t *bytes) {
    QGB_TRY
    void *p; int64_t b;
    if (!anyDataPtr(h, &p, &b))
        return fail(QGB_ERR_RUNTIME, "not a CPU qstates.");
    *ptr = reinterpret_cast<uint64_t>(p);
    if (bytes) *bytes = b;
    QGB_CATCH
}

int qgb_qstates_alt_buffer(qgb_handle h, uint64_t *ptr) {
    QGB_TRY
    void *p; int64_t b;
    if (!anyDataPtr(h, &p, &b))
        return fail(QGB_ERR_RUNTIME, "not a CPU qstates.");
    if (!g_alt.count(h)) g_alt[h] = malloc((size_t)b);
    *ptr = reinterpret_cast<uint64_t>(g_alt[h]);
    QGB_CATCH
}

int qgb_qstates_flip(qgb_handle h) {
    QGB_TRY
    /* the reference object owns its array: exchange the contents instead of the pointers */
    void *p; int64_t b;
    if (!anyDataPtr(h, &p, &b) || !g_alt.count(h))
        return fail(QGB_ERR_RUNTIME, "qstates has no spare buffer.");
    std::vector<char> tmp((size_t)b);
    std::memcpy(tmp.data(), p, (size_t)b);
    std::memcpy(p, g_alt[h], (size_t)b);
    std::memcpy(g_alt[h], tmp.data(), (size_t)b);
    QGB_CATCH
}

int qgb_qproc_calc_norm(qgb_handle, qgb_handle h, double *norm) {
    QGB_TRY
    void *p; int64_t b;
    double acc = 0.;
    if (dataPtr<double>(QS(h), &p, &b)) {
        const double *v = static_cast<const double *>(p);
        for (int64_t i = 0; i < b / 8; ++i) acc += v[i] * v[i];
    } else if (dataPtr<float>(QS(h), &p, &b)) {
        const float *v = static_cast<const float *>(p);
        for (int64_t i = 0; i < b / 4; ++i) acc += (double)v[i] * v[i];
    } else
        return fail(QGB_ERR_RUNTIME, "not a CPU qstates.");
    *norm = acc;
    QGB_CATCH
}

int qgb_qproc_join_shard(qgb_handle, qgb_handle dst, const qgb_handle *src, int n_src, int n_new,
                         int n_total, int64_t offset) {
    QGB_TRY
    void *p; int64_t b;
    int n_dst = QS(dst)->getNLanes();
    if (dataPtr<double>(QS(dst), &p, &b))
        joinShard<double>(static_cast<std::complex<double> *>(p), n_dst, src, n_src, n_total - n_new, offset);
    else if (dataPtr<float>(QS(dst), &p, &b))
        joinShard<float>(static_cast<std::complex<float> *>(p), n_dst, src, n_src, n_total - n_new, offset);
    else
        return fail(QGB_ERR_RUNTIME, "not a CPU qstates.");
    QGB_CATCH
}

int qgb_getter_create_sampling_pool_partial(qgb_handle getter, const int *lane_tables, const int *n_per,
                                            const qgb_handle *qstates_list, int n_qstates, int n_lanes,
                                            int n_hidden, qgb_handle *pool, double *local_total) {
    QGB_TRY
    bool fp64 = dynamic_cast<qgate_cpu::CPUQubitsStatesGetter<double> *>(QG(getter)) != NULL;
    ShimPool *sp = new ShimPool();
    sp->fp32 = !fp64;
    sp->finalized = false;
    int64_t n = 1LL << n_lanes;
    sp->cum.resize((size_t)n);
    if (fp64) {
        QG(getter)->prepareProbArray(sp->cum.data(), toTables(lane_tables, n_per, n_qstates),
                                     toList(qstates_list, n_qstates), n_lanes, n_hidden);
    } else {
        std::vector<float> tmp((size_t)n);
        QG(getter)->prepareProbArray(tmp.data(), toTables(lane_tables, n_per, n_qstates),
                                     toList(qstates_list, n_qstates), n_lanes, n_hidden);
        for (int64_t i = 0; i < n; ++i) sp->cum[i] = tmp[i];
    }
    double acc = 0.;
    for (int64_t i = 0; i < n; ++i) { acc += sp->cum[i]; sp->cum[i] = acc; }
    *local_total = acc;
    *pool = reinterpret_cast<qgb_handle>(static_cast<qgate::SamplingPool *>(sp));
    QGB_CATCH
}

int qgb_pool_finalize(qgb_handle pool, double offset, double total) {
    QGB_TRY
    ShimPool *sp = dynamic_cast<ShimPool *>(SP(pool));
    if (sp == NULL || sp->finalized)
        return fail(QGB_ERR_RUNTIME, "sampling pool is already finalized.");
    double norm = 1. / total;
    for (size_t i = 0; i < sp->cum.size(); ++i) sp->cum[i] = (offset + sp->cum[i]) * norm;
    sp->finalized = true;
    QGB_CATCH
}

int qgb_pool_from_prob_array(int prec, const double *prob, int n_lanes, const int *empty_lanes,
                             int n_empty, qgb_handle *pool) {
    QGB_TRY
    ShimPool *sp = new ShimPool();
    sp->fp32 = prec == QGB_PREC_FP32;
    sp->cum.assign(prob, prob + (1LL << n_lanes));
    double acc = 0.;
    for (size_t i = 0; i < sp->cum.size(); ++i) { acc += sp->cum[i]; sp->cum[i] = acc; }
    if (!(std::fabs(acc - 1.) <= 0.05)) {
        delete sp;
        return fail(QGB_ERR_RUNTIME, "error in probability sum is beyond 0.05.");
    }
    for (size_t i = 0; i < sp->cum.size(); ++i) sp->cum[i] *= 1. / acc;
    sp->empty.assign(empty_lanes, empty_lanes + n_empty);
    std::sort(sp->empty.begin(), sp->empty.end());
    sp->finalized = true;
    *pool = reinterpret_cast<qgb_handle>(static_cast<qgate::SamplingPool *>(sp));
    QGB_CATCH
}

/* Shim-only (not in include/qgate_b200.h; tests bind it by name): the reference's OWN pool,
 * qgate_cpu::CPUSamplingPool<V> (CPUSamplingPool.cpp:8-63), built from a caller-supplied
 * probability vector — what pins the engine's reference-compatible scan. */
int qgb_ref_pool_from_prob_array(int prec, const double *prob, int n_lanes, const int *empty_lanes,
                                 int n_empty, qgb_handle *pool) {
    QGB_TRY
    const size_t n = (size_t)1 << n_lanes;
    qgate::IdList empties(empty_lanes, empty_lanes + n_empty);
    qgate::SamplingPool *sp;
    if (prec == QGB_PREC_FP32) {
        float *p = static_cast<float *>(malloc(sizeof(float) * n)); /* the pool frees it */
        for (size_t i = 0; i < n; ++i) p[i] = (float)prob[i];
        sp = new qgate_cpu::CPUSamplingPool<float>(p, n_lanes, empties);
    } else {
        double *p = static_cast<double *>(malloc(sizeof(double) * n));
        for (size_t i = 0; i < n; ++i) p[i] = prob[i];
        sp = new qgate_cpu::CPUSamplingPool<double>(p, n_lanes, empties);
    }
    *pool = reinterpret_cast<qgb_handle>(sp);
    QGB_CATCH
}

int qgb_qstates_ipc_export(qgb_handle, void *, int64_t *) { return fail(QGB_ERR_RUNTIME, "no CUDA IPC in the CPU shim."); }
int qgb_ipc_open(const void *, uint64_t *) { return fail(QGB_ERR_RUNTIME, "no CUDA IPC in the CPU shim."); }
int qgb_ipc_close(uint64_t) { return fail(QGB_ERR_RUNTIME, "no CUDA IPC in the CPU shim."); }
int qgb_qstates_exchange_p2p(qgb_handle, const uint64_t *, int, const int *, int) {
    return fail(QGB_ERR_RUNTIME, "no peer-memory exchange in the CPU shim.");
}
int qgb_pool_trim_exported(void) { return QGB_OK; }
int qgb_qstates_ipc_export_alt(qgb_handle, void *, int64_t *) { return fail(QGB_ERR_RUNTIME, "no CUDA IPC in the CPU shim."); }
int qgb_qstates_exchange_push(qgb_handle, const uint64_t *, int, const int *, int) {
    return fail(QGB_ERR_RUNTIME, "no peer-memory exchange in the CPU shim.");
}

int qgb_stats_get(qgb_stats *out) { *out = g_stats; return QGB_OK; }
int qgb_stats_reset(void) { std::memset(&g_stats, 0, sizeof(g_stats)); return QGB_OK; }
int qgb_qproc_flush(qgb_handle, qgb_handle) { return QGB_OK; }
int qgb_set_option(const char *, int64_t) { return QGB_OK; }

} // extern "C"
