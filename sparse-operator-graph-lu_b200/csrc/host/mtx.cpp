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
#include <algorithm>
#include <chrono>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace soglu {

namespace {
// Whole file in memory with every '\n' replaced by '\0', plus the start offset of each line: the reference parses
// each std::getline() line with strtol / strtod on its c_str(), and a NUL-terminated slice of the buffer behaves
// exactly like that (a field missing at the end of a line must not be taken from the next line).
struct Lines {
    std::vector<char> buf;
    std::vector<size_t> start;     // start[k] .. start[k + 1] - 1 = line k (without its terminator)
    size_t length(size_t k) const { return start[k + 1] - start[k] - 1; }
    const char* c_str(size_t k) const { return buf.data() + start[k]; }
    bool has_percent(size_t k) const { return std::memchr(c_str(k), '%', length(k)) != nullptr; }
};
bool slurp_lines(const std::string& path, Lines& L) {
    FILE* fp = std::fopen(path.c_str(), "rb");
    if (!fp) return false;
    std::fseek(fp, 0, SEEK_END);
    const long sz = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    L.buf.resize((size_t)(sz > 0 ? sz : 0) + 1);
    const size_t got = sz > 0 ? std::fread(L.buf.data(), 1, (size_t)sz, fp) : 0;
    std::fclose(fp);
    L.buf.resize(got + 1);
    const bool open_tail = got > 0 && L.buf[got - 1] != '\n';   // last line without a newline still counts (getline)
    L.buf[got] = '\n';
    const size_t n = got + (open_tail ? 1 : 0);
    // line starts: count per chunk, prefix, fill (two parallel passes over the buffer)
    int nth = 1;
#ifdef _OPENMP
    nth = omp_get_max_threads();
#endif
    if (n < (size_t(1) << 20)) nth = 1;
    const size_t chunk = (n + nth - 1) / nth;
    std::vector<size_t> cnt(nth + 1, 0);
#pragma omp parallel for schedule(static, 1) num_threads(nth)
    for (int t = 0; t < nth; t++) {
        size_t c = 0;
        for (size_t i = t * chunk, e = std::min(n, i + chunk); i < e; i++) c += L.buf[i] == '\n';
        cnt[t + 1] = c;
    }
    for (int t = 0; t < nth; t++) cnt[t + 1] += cnt[t];
    L.start.resize(cnt[nth] + 1);
    L.start[0] = 0;
#pragma omp parallel for schedule(static, 1) num_threads(nth)
    for (int t = 0; t < nth; t++) {
        size_t k = cnt[t];
        for (size_t i = t * chunk, e = std::min(n, i + chunk); i < e; i++)
            if (L.buf[i] == '\n') { L.buf[i] = '\0'; L.start[++k] = i + 1; }
    }
    return true;
}
}  // namespace

long read_mtx(const std::string& path, Coo& out) {
    out = Coo();
    Lines L;
    const bool timing = std::getenv("SOGLU_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[mtx] %-20s %8.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    };
    if (!slurp_lines(path, L)) return 0;
    lap("read + split lines");
    const size_t nl = L.start.size() - 1;
    if (nl > 0 && std::string(L.c_str(0), L.length(0)).find("symmetric") != std::string::npos) out.symmetric = true;

    long rows = 0, declared = 0;
    size_t first = nl;      // first entry line
    bool have_size = false;
    for (size_t k = 0; k < nl; k++) {
        if (L.length(k) <= 3) continue;
        if (L.has_percent(k)) continue;
        char *p1, *p2;
        rows = std::strtol(L.c_str(k), &p1, 10);
        long cols = std::strtol(p1, &p2, 10);
        declared = std::strtol(p2, nullptr, 10);
        if (rows != cols && rows > cols) rows = cols;
        first = k + 1;
        have_size = true;
        break;
    }
    if (!have_size) return 0;
    out.n = (int)rows;
    // entry lines in parallel chunks; every chunk keeps the reference's filters, the chunks are concatenated in file
    // order and cut at `declared` entries (where the serial reader stops)
    int nth = 1;
#ifdef _OPENMP
    nth = omp_get_max_threads();
#endif
    const size_t n_lines = nl - first;
    if (n_lines < 100000) nth = 1;
    const size_t chunk = (n_lines + nth - 1) / std::max(nth, 1);
    std::vector<std::vector<int>> ci(nth), cj(nth);
    std::vector<std::vector<double>> cv(nth);
    long below_one = 0;
#pragma omp parallel for schedule(static, 1) num_threads(nth)
    for (int t = 0; t < nth; t++) {
        const size_t lo = first + t * chunk, hi = std::min(nl, lo + chunk);
        if (lo >= hi) continue;
        ci[t].reserve(hi - lo); cj[t].reserve(hi - lo); cv[t].reserve(hi - lo);
        for (size_t k = lo; k < hi; k++) {
            const size_t len = L.length(k);
            if (len <= 3) continue;
            if (L.has_percent(k)) continue;
            if (len >= 1000) continue;
            char *p1, *p2;
            long r = std::strtol(L.c_str(k), &p1, 10);
            long c = std::strtol(p1, &p2, 10);
            double val = std::strtod(p2, nullptr);
            if (val == 0) continue;
            if (r > rows || c > rows) continue;
            if (r < 1 || c < 1) { __atomic_fetch_add(&below_one, 1L, __ATOMIC_RELAXED); continue; }   // the reference would index [-1]
            ci[t].push_back((int)r - 1);
            cj[t].push_back((int)c - 1);
            cv[t].push_back(val);
        }
    }
    lap("parse");
    if (below_one) std::fprintf(stderr, "%s: %ld entries with a row or column index below 1 were dropped (MatrixMarket indices are 1-based)\n", path.c_str(), below_one);
    std::vector<size_t> off(nth + 1, 0);
    for (int t = 0; t < nth; t++) off[t + 1] = off[t] + cv[t].size();
    const size_t keep = std::min<size_t>(off[nth], declared > 0 ? (size_t)declared : 0);   // the reference would overrun its arrays here
    out.i.resize(keep); out.j.resize(keep); out.v.resize(keep);
#pragma omp parallel for schedule(static, 1) num_threads(nth)
    for (int t = 0; t < nth; t++)
        for (size_t q = 0, e = cv[t].size(); q < e && off[t] + q < keep; q++) {
            out.i[off[t] + q] = ci[t][q]; out.j[off[t] + q] = cj[t][q]; out.v[off[t] + q] = cv[t][q];
        }
    lap("concatenate");
    return (long)keep;
}

long read_array(const std::string& path, int dim, std::vector<double>& b) {
    b.assign(dim, 1.0);
    Lines L;
    if (!slurp_lines(path, L)) return 0;
    const size_t nl = L.start.size() - 1;
    size_t k = 0;
    for (; k < nl; k++) {
        if (L.has_percent(k)) continue;
        k++;
        break;  // size line
    }
    long count = 0;
    for (; k < nl; k++) {
        if (count >= dim) break;
        if (L.has_percent(k)) continue;
        if (L.length(k) >= 1000) continue;
        b[count++] = std::strtod(L.c_str(k), nullptr);
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
