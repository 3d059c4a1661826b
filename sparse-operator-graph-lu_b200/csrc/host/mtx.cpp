// MatrixMarket coordinate reader / array reader / residual check.
// Behaviour follows the reference reader (mtx.cpp:44-215) including its quirks, because
// the COO *order* and which entries survive feed the bit-exact planner:
//   * "symmetric" is detected by substring in the first line            (mtx.cpp:59-60)
//   * lines of <= 3 characters or containing '%' are skipped            (mtx.cpp:63-64, 92-93)
//   * the first surviving line is "rows cols nnz"; non-square -> min    (mtx.cpp:68-76)
//   * entry lines of >= 1000 characters are skipped                     (mtx.cpp:94)
//   * explicit zeros and out-of-range indices are dropped               (mtx.cpp:102-105)
//   * array files: '%' lines skipped, first other line is the size line (mtx.cpp:154-160),
//     missing values become 1.0                                         (mtx.cpp:177-179)
#include "soglu_host.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

namespace soglu {

long read_mtx(const std::string& path, Coo& out) {
    out = Coo();
    std::ifstream f(path.c_str());
    if (!f.is_open()) return 0;
    std::string line;
    std::getline(f, line);
    if (line.find("symmetric") != std::string::npos) out.symmetric = true;
    f.clear();
    f.seekg(0, f.beg);

    long rows = 0, declared = 0;
    bool have_size = false;
    while (std::getline(f, line)) {
        if (line.length() <= 3) continue;
        if (line.find('%') != std::string::npos) continue;
        char *p1, *p2;
        const char* s = line.c_str();
        rows = std::strtol(s, &p1, 10);
        long cols = std::strtol(p1, &p2, 10);
        declared = std::strtol(p2, nullptr, 10);
        if (rows != cols && rows > cols) rows = cols;
        have_size = true;
        break;
    }
    if (!have_size) return 0;
    out.n = (int)rows;
    out.i.reserve(declared);
    out.j.reserve(declared);
    out.v.reserve(declared);
    long count = 0;
    while (std::getline(f, line)) {
        if (line.length() <= 3) continue;
        if (line.find('%') != std::string::npos) continue;
        if (line.length() >= 1000) continue;
        char *p1, *p2;
        const char* s = line.c_str();
        long r = std::strtol(s, &p1, 10);
        long c = std::strtol(p1, &p2, 10);
        double val = std::strtod(p2, nullptr);
        if (val == 0) continue;
        if (r > rows || c > rows) continue;
        if (count >= declared) break;  // the reference would overrun its arrays here
        out.i.push_back((int)r - 1);
        out.j.push_back((int)c - 1);
        out.v.push_back(val);
        count++;
    }
    return count;
}

long read_array(const std::string& path, int dim, std::vector<double>& b) {
    b.assign(dim, 1.0);
    std::ifstream f(path.c_str());
    long count = 0;
    if (f.is_open()) {
        std::string line;
        while (std::getline(f, line)) {
            if (line.find('%') != std::string::npos) continue;
            break;  // size line
        }
        while (std::getline(f, line)) {
            if (count >= dim) break;
            if (line.find('%') != std::string::npos) continue;
            if (line.length() >= 1000) continue;
            b[count++] = std::strtod(line.c_str(), nullptr);
        }
    }
    return count;
}

bool write_array(const std::string& path, const double* a, int dim) {
    FILE* fp = std::fopen(path.c_str(), "w");
    if (!fp) return false;
    std::fprintf(fp, "%%%%MatrixMarket matrix array real general\n%d 1\n", dim);
    for (int i = 0; i < dim; i++) std::fprintf(fp, "%.17g\n", a[i]);
    std::fclose(fp);
    return true;
}

double check_result(const Coo& a, const double* b, const double* x) {
    std::vector<double> ax(a.n, 0.0);
    size_t nnz = a.v.size();
    for (size_t k = 0; k < nnz; k++) {
        ax[a.i[k]] += a.v[k] * x[a.j[k]];
        if (a.symmetric && a.i[k] != a.j[k]) ax[a.j[k]] += a.v[k] * x[a.i[k]];
    }
    double mx = 0;
    for (int i = 0; i < a.n; i++) {
        double t = std::fabs(b[i] - ax[i]);
        if (t > mx || std::isnan(t) || std::isinf(t)) mx = t;
    }
    return mx;
}

void Config::set(int dim) {
    mSize = dim;
    blockSize = 64;
    int multiple = dim / blockSize, r = 1;
    while (multiple > 0) { multiple /= 2; r *= 2; }
    blockRows = r;
    blockRowsL2 = 2 << (__builtin_popcount(blockRows - 1) / 2);
    blockSizeL2 = blockSize * blockRows / blockRowsL2;
}

}  // namespace soglu
