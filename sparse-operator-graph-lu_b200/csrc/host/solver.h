// Drop-in mirror of the reference's library entry (solver.h:20-24): same namespace, names
// and argument meaning.  Differences, both deliberate: the returned x is malloc'ed (free
// it with soglu_free / std::free; the reference returns arena memory owned by memutil),
// and a NULL return means an error whose text is in soglu_last_error().
#pragma once
#include "../../../include/soglu.h"

namespace SOGLU {
int iniData();
double* solveLU(int dim, int valcount, bool symmetric, int* index_i, int* index_j, double* vals, double* b);
}  // namespace SOGLU
