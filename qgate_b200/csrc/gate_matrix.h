/* gate_matrix.h — the 16 single-qubit matrices of qgate's gate set (see gate_matrix.cpp). */
#pragma once

namespace qgb {

/* number of arguments gate `gate_id` (QGB_GATE_* of include/qgate_b200.h) takes; -1 if unknown */
int gate_matrix_n_args(int gate_id);
/* m receives (re, im) of m00, m01, m10, m11; false for an unknown gate id */
bool gate_matrix(int gate_id, const double *args, int adjoint, double *m);

} // namespace qgb
