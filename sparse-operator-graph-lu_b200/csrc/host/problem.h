// A planned problem: everything the host front-end produces before the device hot path
// starts (the state the reference keeps in data::*, GOrder::* and config::* statics).
#pragma once
#include <string>
#include <vector>

#include "../../../include/soglu.h"
#include "soglu_host.h"

namespace soglu {

struct Problem {
    int dim = 0;
    bool symmetric = false;
    int n_ext = 0;
    Coo a;                         // original matrix (as read)
    std::vector<double> b;         // original rhs
    std::vector<int> pi, pj;       // expanded + permuted COO (data::indexi/indexj)
    std::vector<double> pv;        // data::vals
    std::vector<double> b_perm;    // permuted rhs padded with 1.0 to n_ext (data::b)
    Config cfg;
    Ordering ord;
    Plan plan;
    double flops = 0;
    double t_reorder = 0, t_plan = 0;
    std::string log;
};

void set_error(const std::string& s);
int prepare_problem(Problem& P, int dim, int64_t nnz, bool symmetric, const int* ii, const int* jj, const double* vv,
                    const double* b);

}  // namespace soglu
