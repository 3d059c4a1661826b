// Host front-end of soglu-b200: MatrixMarket reader, size configuration, GPS
// bandwidth-reducing ordering and the two-level block planner.  Everything in this
// header produces INTEGER output (permutation, block ids, operation lists, stages)
// that must equal the reference's bit for bit; the numeric hot path lives behind
// include/soglu.h on the GPU.
//
// Reference seams (file:line in hotlei/sparse-operator-graph-LU):
//   mtx.cpp:44-215        reader / writer / residual check      -> soglu::Mtx*
//   config.cpp:30-49      block-size configuration              -> soglu::Config
//   GPSOrder.cpp:41-635   GPS ordering, sortInBlock             -> soglu::gps_*
//   BlockPlanner.cpp:73-374, 865-1577   symbolic planner         -> soglu::Planner
//   solver.cpp:37-184     driver                                -> SOGLU::solveLU (solver.h)
#pragma once
#include <cstdint>
#include <string>
#include <cstdlib>
#include <vector>

#include "../bigalloc.h"

namespace soglu {

// ---- operation codes: same numbering as the reference enum (operation.h:20) ----------
enum BlockOp : uint8_t {
    OP_INV = 0, OP_LU = 1, OP_LOWERINV = 2, OP_UPPERINV = 3, OP_SUB = 4, OP_ADD = 5, OP_NEG = 6,
    OP_COPY = 7, OP_MUL = 8, OP_MULNEG = 9, OP_LLT = 10, OP_MULT = 11, OP_NOOP = 12
};

// ---- MatrixMarket I/O (mtx.cpp) -------------------------------------------------------
struct Coo {
    int n = 0;                 // matrix dimension (min(rows, cols) like the reference)
    bool symmetric = false;    // "symmetric" appears in the banner line
    std::vector<int> i, j;     // 0-based
    std::vector<double> v;
};
// returns number of entries read, 0 if the file cannot be opened (mtx.cpp:44-117)
long read_mtx(const std::string& path, Coo& out);
// reads an "array" file into b[0..dim); missing tail (or missing file) is filled with 1.0
// (mtx.cpp:145-182).  Returns the number of values actually read.
long read_array(const std::string& path, int dim, std::vector<double>& b);
// full-precision array writer (the reference prints 6 digits, mtx.cpp:130-143)
bool write_array(const std::string& path, const double* a, int dim);
// max_i |b - A x| in the original ordering, symmetric-aware (mtx.cpp:183-215)
double check_result(const Coo& a, const double* b, const double* x);

// ---- size configuration (config.cpp:41-49) --------------------------------------------
struct Config {
    int mSize = 0;
    int blockSize = 64;
    int blockRows = 0;      // power of two, 2^(floor(log2(dim/64))+1)
    int blockRowsL2 = 0;    // 2 << (log2(blockRows)/2)
    int blockSizeL2 = 0;    // 64*blockRows/blockRowsL2
    void set(int dim);
};

// ---- GPS ordering (GPSOrder.cpp) --------------------------------------------------------
struct Ordering {
    // Same (swapped) naming as the reference, GPSOrder.cpp:448-449:
    //   newOrder[k]     = original index of the row that became row k   (new -> old)
    //   reverseOrder[i] = new index of original row i                   (old -> new)
    std::vector<int> newOrder, reverseOrder;
    // summary line the reference prints (GPSOrder.cpp:161-175)
    int levels = 0, width = 0, lastLevelCount = 0, accounted = 0, startNode = 0;
};
// Permutes idx_i/idx_j in place, rebuilds b as the permuted rhs padded with 1.0 to
// blockRows*64 (GPSOrder.cpp:425-447), fills ord.  `vals` is only read by sort_in_block.
void gps_reorder(int dim, std::vector<int>& idx_i, std::vector<int>& idx_j, std::vector<double>& b,
                 const Config& cfg, Ordering& ord);
void sort_in_block(int dim, std::vector<int>& idx_i, std::vector<int>& idx_j, const std::vector<double>& vals,
                   std::vector<double>& b, Ordering& ord);

// ---- planner output -----------------------------------------------------------------------
struct Op {                 // one DAG node; mirrors struct operation (operation.h:37-52)
    int32_t src, src2, result, result2;
    int32_t stage, seq, group;
    uint8_t op;
};
struct BlockRef { int32_t id, brow, bcol; };

// the op lists (10^8 entries at n = 10^6) live in huge-page backed vectors, see bigalloc.h
using OpVec = BigVec<Op>;

// Dense values of the input blocks: one malloc, zeroed in parallel (first touch by all threads).
class BlockValues {
    double* p_ = nullptr;
    size_t n_ = 0;
  public:
    BlockValues() = default;
    BlockValues(const BlockValues&) = delete;
    BlockValues& operator=(const BlockValues&) = delete;
    BlockValues(BlockValues&& o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
    BlockValues& operator=(BlockValues&& o) noexcept { if (this != &o) { std::free(p_); p_ = o.p_; n_ = o.n_; o.p_ = nullptr; o.n_ = 0; } return *this; }
    ~BlockValues() { std::free(p_); }
    bool alloc_zero(size_t n_blocks);   // false when out of memory
    double* data() { return p_; }
    const double* data() const { return p_; }
    size_t size() const { return n_; }
    double& operator[](size_t i) { return p_[i]; }
};

struct Plan {
    Config cfg;
    bool symmetric = false;
    // coarse (L2) pass, kept for the bit-exact checks
    OpVec coarse_ops;
    int coarse_storage = 0;
    int coarse_emitted = 0;          // op count before pruning
    // fine pass
    OpVec ops;                       // sorted by (stage, group, result, src, seq)
    int fine_emitted = 0;
    int storage = 0;                 // data::storageCount: block ids are 1..storage-1, 0 = none
    std::vector<int32_t> stage;      // per block id (data::stage)
    std::vector<int32_t> laststage;  // per block id (data::laststage)
    std::vector<BlockRef> inputs;    // blocks filled by iniBlockStorage, quadtree (Z) order
    // values of the input blocks, duplicate-free: input block id (1..n_input, ids are handed out in allocation
    // order), position row*64+col inside the block, value; the identity padding is included
    BigVec<int32_t> entry_block, entry_pos;
    BigVec<double> entry_val;
    bool dense_inputs(BlockValues& out) const;   // dense 64x64 row-major per input block (index id-1), on demand
    std::vector<BlockRef> L, U;      // factor leaves with block coordinates, quadtree order
    std::vector<int32_t> brow, bcol; // per block id: coordinates of the quadtree slot (or -1)
    std::string log;                 // the lines the reference prints while planning
};

// Runs both planning passes (solver.cpp:50-100) on the permuted COO.
// Returns 0 on success; non-zero with plan.log holding the reason otherwise.
int build_plan(const Config& cfg, bool symmetric, const std::vector<int>& idx_i, const std::vector<int>& idx_j,
               const std::vector<double>& vals, Plan& plan, bool keep_values = true);

// dense-block FLOP convention of SURVEY.md section 8(d)
double factor_flops(const OpVec& ops);

}  // namespace soglu
