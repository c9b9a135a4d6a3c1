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

#include <cstring>
#include <exception>
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

int qgb_pool_delete(qgb_handle pool) {
    QGB_TRY
    delete SP(pool);
    QGB_CATCH
}

int qgb_stats_get(qgb_stats *out) { *out = g_stats; return QGB_OK; }
int qgb_stats_reset(void) { std::memset(&g_stats, 0, sizeof(g_stats)); return QGB_OK; }
int qgb_qproc_flush(qgb_handle, qgb_handle) { return QGB_OK; }
int qgb_set_option(const char *, int64_t) { return QGB_OK; }

} // extern "C"
