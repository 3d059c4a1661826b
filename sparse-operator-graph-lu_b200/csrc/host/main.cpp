// ./solve <file.mtx> -- same command line and console output as the reference CLI
// (main.cpp:34-66).  Two documented departures: <base>_x.mtx receives the actual solution
// at full precision (the reference writes the forward-substituted rhs data::b at 6 digits,
// main.cpp:61 / mtx.cpp:137), and a failure prints the library error text.
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "soglu_host.h"
#include "solver.h"

int main(int argc, char* argv[]) {
    SOGLU::iniData();
    if (argc <= 1 || std::string(argv[1]).find(".mtx") == std::string::npos) {
        std::cout << "usage: ./solve filename.mtx" << std::endl;
        return 0;
    }
    std::string fname = argv[1];
    std::string filebase = fname.substr(0, fname.find(".mtx"));
    soglu::Coo a;
    if (soglu::read_mtx(fname, a) == 0) {
        std::cout << "Can not open file" << '\n';
        return 0;
    }
    std::vector<double> b;
    soglu::read_array(filebase + "_b.mtx", a.n, b);
    double* x = SOGLU::solveLU(a.n, (int)a.v.size(), a.symmetric, a.i.data(), a.j.data(), a.v.data(), b.data());
    if (!x) return 1;
    std::cout << "max rhs error:" << soglu::check_result(a, b.data(), x) << std::endl;
    soglu::write_array(filebase + "_x.mtx", x, a.n);
    if (a.n > 3) std::cout << x[0] << " " << x[1] << " " << x[2] << " ";
    std::cout << "...";
    if (a.n > 3) std::cout << " " << x[a.n - 3] << " " << x[a.n - 2] << " " << x[a.n - 1] << " ";
    std::cout << '\n';
    std::free(x);
    return 0;
}
