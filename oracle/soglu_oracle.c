/* TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product
 * (sparse-operator-graph-lu_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file.
 *
 * CPU restatement ("port") of the numeric hot path of hotlei/sparse-operator-graph-LU:
 * sequential, scalar, dense 64x64 FP64 blocks, one function per reference kernel.  It
 * executes the reference's flat operation list exactly in list order, i.e. the order
 * BlockPlanner::calculate walks the stages (BlockPlanner.cpp:376-651), and then the block
 * forward/back substitution of BlockPlanner::solve (BlockPlanner.cpp:834-862).
 *
 * Differences from the reference's own kernels, all below the 1e-10 parity tolerance
 * (SURVEY.md section 8c validated the dense restatement to <= 8e-14 against the reference):
 *   - no 8x8 sub-block bitmaps, hence no dropping of |v| <= 1e-18 results
 *     (MatrixStdDouble.cpp:622-631);
 *   - plain double accumulation where the reference uses x87 long double
 *     (MatrixStdDouble.cpp:2714, 2744, 2762, 2860).
 *
 * Parity pin: tests/test_oracle.py checks this file against the golden vectors recorded
 * from the UNMODIFIED reference (tests/golden/, made by tools/make_golden.py through
 * oracle/_ref/ref_harness) and, when oracle/_ref exists, against live runs of it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define B 64
#define BB (B * B)

enum { OP_LU = 1, OP_LOWERINV = 2, OP_UPPERINV = 3, OP_SUB = 4, OP_MUL = 8, OP_MULNEG = 9, OP_LLT = 10, OP_MULT = 11 };

typedef struct {
    int64_t n_ids;
    double** blk; /* NULL until produced, like data::blockstorage (data.h:47) */
} oracle_t;

static double* get_or_zero(oracle_t* o, int32_t id) {
    if (!o->blk[id]) o->blk[id] = (double*)calloc(BB, sizeof(double));
    return o->blk[id];
}

/* C += A*B / C -= A*B  (blockMulOneAvxBlock / ...Neg, MatrixStdDouble.cpp:2043-2507, 237-701) */
static void k_mul(const double* a, const double* b, double* c, double sign) {
    for (int i = 0; i < B; i++)
        for (int k = 0; k < B; k++) {
            double aik = sign * a[i * B + k];
            if (aik == 0.0) continue;
            for (int j = 0; j < B; j++) c[i * B + j] += aik * b[k * B + j];
        }
}
/* C += A*B^T  (mat_mult, MatrixStdDouble.cpp:786-1095, 719-768) */
static void k_mult(const double* a, const double* b, double* c) {
    for (int i = 0; i < B; i++)
        for (int j = 0; j < B; j++) {
            double s = 0;
            for (int k = 0; k < B; k++) s += a[i * B + k] * b[j * B + k];
            c[i * B + j] += s;
        }
}
/* R = S2 - S1, missing source = zero  (mat_sub / mat_copy / mat_neg, MatrixStdDouble.cpp:2948-3121;
 * dispatch BlockPlanner.cpp:553-572) */
static void k_sub(const double* s2, const double* s1, double* r) {
    for (int i = 0; i < BB; i++) r[i] = (s2 ? s2[i] : 0.0) - (s1 ? s1[i] : 0.0);
}
/* Doolittle LU without pivoting, unit-diagonal L, |u_ii| < 1e-9 clamped sign-preserving
 * (ludcmpSimple, MatrixStdDouble.cpp:2711-2784) */
static void k_lu(const double* a, double* l, double* u) {
    memset(l, 0, BB * sizeof(double));
    memset(u, 0, BB * sizeof(double));
    for (int i = 0; i < B; i++) {
        for (int j = i; j < B; j++) {
            double s = 0;
            for (int k = 0; k < i; k++) s += l[i * B + k] * u[k * B + j];
            u[i * B + j] = a[i * B + j] - s;
        }
        if (u[i * B + i] < 1e-9 && u[i * B + i] > -1e-9) u[i * B + i] = (u[i * B + i] < 0) ? -1e-9 : 1e-9;
        double inv = 1.0 / u[i * B + i];
        for (int j = i + 1; j < B; j++) {
            double s = 0;
            for (int k = 0; k < i; k++) s += l[j * B + k] * u[k * B + i];
            l[j * B + i] = inv * (a[j * B + i] - s);
        }
        l[i * B + i] = 1.0;
    }
}
/* Cholesky on the lower triangle, pivot < 1e-20 clamped (lltdcmpSimple, MatrixStdDouble.cpp:2629-2668) */
static void k_llt(const double* a, double* l) {
    memset(l, 0, BB * sizeof(double));
    for (int j = 0; j < B; j++) {
        double s = 0;
        for (int k = 0; k < j; k++) s += l[j * B + k] * l[j * B + k];
        double p = a[j * B + j] - s;
        if (p < 1e-20) p = 1e-20;
        p = sqrt(p);
        l[j * B + j] = p;
        p = 1.0 / p;
        for (int i = j + 1; i < B; i++) {
            double t = 0;
            for (int k = 0; k < j; k++) t += l[i * B + k] * l[j * B + k];
            l[i * B + j] = p * (a[i * B + j] - t);
        }
    }
}
/* Y = L^-1 by forward substitution, general diagonal (inv_lower, MatrixStdDouble.cpp:2787-2802) */
static void k_inv_lower(const double* l, double* y) {
    memset(y, 0, BB * sizeof(double));
    for (int i = 0; i < B; i++) {
        double q = 1.0 / l[i * B + i];
        y[i * B + i] = q;
        for (int j = 0; j < i; j++) {
            double s = 0;
            for (int k = j; k < i; k++) s += l[i * B + k] * y[k * B + j];
            y[i * B + j] = -s * q;
        }
    }
}
/* Y = U^-1 by back elimination on the identity (inv_upper, MatrixStdDouble.cpp:2829-2866) */
static void k_inv_upper(const double* u, double* y) {
    memset(y, 0, BB * sizeof(double));
    for (int i = 0; i < B; i++) y[i * B + i] = 1.0;
    for (int j = B - 1; j >= 0; j--) {
        for (int i = j - 1; i >= 0; i--) {
            double scale = u[i * B + j] / u[j * B + j];
            for (int k = j; k < B; k++) y[i * B + k] -= y[j * B + k] * scale;
        }
        double rate = 1.0 / u[j * B + j];
        for (int i = j; i < B; i++) y[j * B + i] *= rate;
    }
}

/* ---- public API ---------------------------------------------------------------------- */
void* oracle_create(int64_t n_ids) {
    oracle_t* o = (oracle_t*)calloc(1, sizeof(oracle_t));
    o->n_ids = n_ids;
    o->blk = (double**)calloc((size_t)n_ids, sizeof(double*));
    return o;
}
void oracle_destroy(void* h) {
    oracle_t* o = (oracle_t*)h;
    if (!o) return;
    for (int64_t i = 0; i < o->n_ids; i++) free(o->blk[i]);
    free(o->blk);
    free(o);
}
/* iniBlockStorage result (BlockPlanner.cpp:1492-1539): dense 64x64 row-major per input block */
void oracle_set_inputs(void* h, int64_t n_input, const int32_t* ids, const double* dense) {
    oracle_t* o = (oracle_t*)h;
    for (int64_t k = 0; k < n_input; k++) memcpy(get_or_zero(o, ids[k]), dense + k * BB, BB * sizeof(double));
}
/* BlockPlanner::calculate: the dispatch switch of BlockPlanner.cpp:458-592 over the list in order.
 * Returns 0, or the 1-based index of the first op that cannot be executed. */
int64_t oracle_factor(void* h, int64_t n_ops, const int32_t* src, const int32_t* src2, const uint8_t* op, const int32_t* result,
                      const int32_t* result2) {
    oracle_t* o = (oracle_t*)h;
    for (int64_t i = 0; i < n_ops; i++) {
        const double* a = src[i] > 0 ? o->blk[src[i]] : NULL;
        const double* b = src2[i] > 0 ? o->blk[src2[i]] : NULL;
        double* r = get_or_zero(o, result[i]); /* accumulation targets start as zero (430-445) */
        switch (op[i]) {
            case OP_MUL: if (!a || !b) return i + 1; k_mul(a, b, r, 1.0); break;
            case OP_MULNEG: if (!a || !b) return i + 1; k_mul(a, b, r, -1.0); break;
            case OP_MULT: if (!a || !b) return i + 1; k_mult(a, b, r); break;
            case OP_SUB: k_sub(b, a, r); break; /* blk[result] = blk[src2] - blk[src] */
            case OP_LU: if (!a) return i + 1; k_lu(a, r, get_or_zero(o, result2[i])); break;
            case OP_LLT: if (!a) return i + 1; k_llt(a, r); break;
            case OP_LOWERINV: if (!a) return i + 1; k_inv_lower(a, r); break;
            case OP_UPPERINV: if (!a) return i + 1; k_inv_upper(a, r); break;
            default: return i + 1;
        }
    }
    return 0;
}
int oracle_get_block(void* h, int32_t id, double* out) {
    oracle_t* o = (oracle_t*)h;
    if (id <= 0 || id >= o->n_ids || !o->blk[id]) return 1;
    memcpy(out, o->blk[id], BB * sizeof(double));
    return 0;
}

/* BlockPlanner::solve (BlockPlanner.cpp:834-862): y = L^-1 b dividing by the stored diagonal
 * (lowerSolver 736-768), x = U^-1 y (upperSolver 800-832) or L^-T y (upperSolverT 769-797).
 * The quadtree recursion visits block rows in order, so the restatement is a plain loop over
 * block rows with the off-diagonal blocks applied before the diagonal solve. */
int oracle_solve(void* h, int64_t nL, const int32_t* L, int64_t nU, const int32_t* U, int32_t n_rows, int symmetric,
                 const double* b_ext, double* x_ext) {
    oracle_t* o = (oracle_t*)h;
    int64_t n = (int64_t)n_rows * B;
    double* y = (double*)malloc(n * sizeof(double));
    double* r = (double*)malloc(n * sizeof(double));
    memcpy(r, b_ext, n * sizeof(double));
    int32_t* diagL = (int32_t*)calloc(n_rows, sizeof(int32_t));
    int32_t* diagU = (int32_t*)calloc(n_rows, sizeof(int32_t));
    for (int64_t k = 0; k < nL; k++) if (L[3 * k + 1] == L[3 * k + 2]) diagL[L[3 * k + 1]] = L[3 * k];
    for (int64_t k = 0; k < nU; k++) if (U[3 * k + 1] == U[3 * k + 2]) diagU[U[3 * k + 1]] = U[3 * k];
    int rc = 0;
    /* forward: for each block row, subtract L_ij y_j (j < i) then solve the diagonal block */
    for (int32_t i = 0; i < n_rows && !rc; i++) {
        for (int64_t k = 0; k < nL; k++) {
            if (L[3 * k + 1] != i || L[3 * k + 2] >= i) continue;
            const double* m = o->blk[L[3 * k]];
            if (!m) { rc = 2; break; }
            const double* yj = y + (int64_t)L[3 * k + 2] * B;
            for (int a = 0; a < B; a++) {
                double s = 0;
                for (int c = 0; c < B; c++) s += m[a * B + c] * yj[c];
                r[(int64_t)i * B + a] -= s;
            }
        }
        const double* d = diagL[i] ? o->blk[diagL[i]] : NULL;
        if (!d) { rc = 3; break; }
        for (int a = 0; a < B; a++) {
            double s = 0;
            for (int c = 0; c < a; c++) s += d[a * B + c] * y[(int64_t)i * B + c];
            y[(int64_t)i * B + a] = (r[(int64_t)i * B + a] - s) / d[a * B + a];
        }
    }
    /* backward */
    memcpy(r, y, n * sizeof(double));
    for (int32_t i = n_rows - 1; i >= 0 && !rc; i--) {
        if (!symmetric) {
            for (int64_t k = 0; k < nU; k++) {
                if (U[3 * k + 1] != i || U[3 * k + 2] <= i) continue;
                const double* m = o->blk[U[3 * k]];
                if (!m) { rc = 4; break; }
                const double* xj = x_ext + (int64_t)U[3 * k + 2] * B;
                for (int a = 0; a < B; a++) {
                    double s = 0;
                    for (int c = 0; c < B; c++) s += m[a * B + c] * xj[c];
                    r[(int64_t)i * B + a] -= s;
                }
            }
            const double* d = diagU[i] ? o->blk[diagU[i]] : NULL;
            if (!d) { rc = 5; break; }
            for (int a = B - 1; a >= 0; a--) {
                double s = 0;
                for (int c = a + 1; c < B; c++) s += d[a * B + c] * x_ext[(int64_t)i * B + c];
                x_ext[(int64_t)i * B + a] = (r[(int64_t)i * B + a] - s) / d[a * B + a];
            }
        } else {
            /* x_i = L_ii^-T (y_i - sum_{j>i} L_ji^T x_j) */
            for (int64_t k = 0; k < nL; k++) {
                if (L[3 * k + 2] != i || L[3 * k + 1] <= i) continue;
                const double* m = o->blk[L[3 * k]];
                if (!m) { rc = 4; break; }
                const double* xj = x_ext + (int64_t)L[3 * k + 1] * B;
                for (int a = 0; a < B; a++) {
                    double s = 0;
                    for (int c = 0; c < B; c++) s += m[c * B + a] * xj[c];
                    r[(int64_t)i * B + a] -= s;
                }
            }
            const double* d = diagL[i] ? o->blk[diagL[i]] : NULL;
            if (!d) { rc = 5; break; }
            for (int a = B - 1; a >= 0; a--) {
                double s = 0;
                for (int c = a + 1; c < B; c++) s += d[c * B + a] * x_ext[(int64_t)i * B + c];
                x_ext[(int64_t)i * B + a] = (r[(int64_t)i * B + a] - s) / d[a * B + a];
            }
        }
    }
    free(y); free(r); free(diagL); free(diagU);
    return rc;
}
