// Two-level symbolic block planner.
//
// Emits the DAG of 64x64 block operations for a recursive 2x2 block LU (or LL^T) with
// explicit triangular inverses, first on coarse "L2" blocks, then expanded to 64-blocks,
// each pass followed by liveness pruning, ASAP stage numbering and the reference's
// total ordering.  The emitted lists (ops, ids, stages, sequence/group numbers) equal the
// reference's bit for bit; tests/test_planner_golden.py checks that against dumps of the
// unmodified reference.
//
// Reference: BlockPlanner.cpp:865-939 (LU), 941-989 (LLT), 1019-1092 (triangular
// inverses), 1126-1283 (sub/copy/neg/mul), 1305-1460 (two-level expansion), 1462-1577
// (block storage), 73-374 (blockPlan); solver.cpp:50-100 (pass order); memutil.cpp:297-299
// and operation.cpp:107-118 (sequence counter side effects).
//
// Own design: the quadtree lives in a chunked pool addressed by 32-bit indices (24-byte
// nodes instead of 48), operations are 32-byte PODs in one vector, the stage sort is a
// counting sort on stage followed by parallel per-stage sorts (the comparator is a total
// order because sequence numbers are unique, so any correct sort reproduces the list).
#include "soglu_host.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <sstream>
#include <chrono>
#include <cstdlib>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace soglu {
namespace {

using NodeId = int32_t;  // 0 == NULL
struct Node {
    NodeId sub[4];   // b11 b12 b21 b22
    int32_t block;   // block id (leaf), 0 = none
    int32_t level;   // 0 = leaf
};

class NodePool {
    static constexpr int SHIFT = 18;
    static constexpr int CHUNK = 1 << SHIFT;
    std::vector<std::unique_ptr<Node[]>> chunks_;
    int64_t count_ = 1;  // slot 0 reserved as NULL
  public:
    NodePool() { chunks_.emplace_back(new Node[CHUNK]()); }
    Node& operator[](NodeId i) { return chunks_[(uint32_t)i >> SHIFT][i & (CHUNK - 1)]; }
    NodeId make(int level) {
        if ((count_ >> SHIFT) >= (int64_t)chunks_.size()) chunks_.emplace_back(new Node[CHUNK]());
        NodeId id = (NodeId)count_++;
        Node& n = (*this)[id];
        n.sub[0] = n.sub[1] = n.sub[2] = n.sub[3] = 0;
        n.block = 0;
        n.level = level;
        return id;
    }
    int64_t size() const { return count_; }
};

// Emission buffer: fixed-size chunks, so appending 10^8 ops never reallocates or copies.
class OpStore {
    static constexpr int SHIFT = 20;
    static constexpr int64_t CHUNK = int64_t(1) << SHIFT;
    std::vector<Op*> chunks_;
    int64_t n_ = 0;
  public:
    OpStore() = default;
    OpStore(const OpStore&) = delete;
    OpStore& operator=(const OpStore&) = delete;
    ~OpStore() { clear(); }
    void clear() {
        for (Op* c : chunks_) big_free(c, CHUNK * sizeof(Op));
        chunks_.clear();
        n_ = 0;
    }
    void push_back(const Op& o) {
        if ((n_ >> SHIFT) >= (int64_t)chunks_.size()) chunks_.push_back(static_cast<Op*>(big_alloc(CHUNK * sizeof(Op))));
        chunks_[n_ >> SHIFT][n_ & (CHUNK - 1)] = o;
        n_++;
    }
    Op& operator[](int64_t i) { return chunks_[i >> SHIFT][i & (CHUNK - 1)]; }
    const Op& operator[](int64_t i) const { return chunks_[i >> SHIFT][i & (CHUNK - 1)]; }
    int64_t size() const { return n_; }
};

// SOGLU_TIMING=1 prints the planner's phase times to stderr (diagnostics only)
struct PhaseTimer {
    bool on = std::getenv("SOGLU_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(const char* what) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[plan] %-28s %8.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    }
};

inline int ilog2(int v) { return __builtin_popcount(v - 1); }  // v is a power of two

struct Planner {
    NodePool T;
    OpStore graph;        // ops in emission order
    int storage = 1;      // data::storageCount (id 0 = none)
    int seq = 0;          // operation::seq
    bool planL2 = false;
    NodeId blocks = 0;    // input quadtree (data::blocks)
    int blockRows = 0, blockSize = 0, mSize = 0;
    std::vector<int32_t> stage, laststage;
    std::ostringstream log;

    // ---- storage ------------------------------------------------------------------
    int new_block() { return storage++; }                       // appendBlockStorage
    void emit(int src, int src2, BlockOp op, int res, int res2, int group) {  // memutil::newoperation
        Op o;
        o.src = src; o.src2 = src2; o.result = res; o.result2 = res2;
        o.stage = 0; o.seq = seq++; o.group = group; o.op = (uint8_t)op;
        graph.push_back(o);
    }
    // quadtree lookup / insert by (block row, block col), MSB first (matrix.cpp:34-88)
    int32_t tree_get(NodeId root, int bi, int bj) {
        NodeId cur = root;
        for (int step = T[root].level - 1; step >= 0; step--) {
            int q = (((bi >> step) & 1) << 1) | ((bj >> step) & 1);
            cur = T[cur].sub[q];
            if (!cur) return 0;
        }
        return T[cur].block;
    }
    void tree_set(NodeId root, int bi, int bj, int32_t val) {
        NodeId cur = root;
        for (int step = T[root].level - 1; step >= 0; step--) {
            int q = (((bi >> step) & 1) << 1) | ((bj >> step) & 1);
            NodeId nx = T[cur].sub[q];
            if (!nx) {
                nx = T.make(T[cur].level - 1);
                T[cur].sub[q] = nx;
            }
            cur = nx;
        }
        T[cur].block = val;
    }
    NodeId child(NodeId c, int q) {   // result-side child, created on demand
        NodeId s = T[c].sub[q];
        if (!s) {
            s = T.make(T[c].level - 1);
            T[c].sub[q] = s;
        }
        return s;
    }
    int32_t leaf_block(NodeId c) {    // result leaf id, allocated on first write
        if (T[c].block <= 0) T[c].block = new_block();
        return T[c].block;
    }

    // ---- symbolic block algebra ------------------------------------------------------
    // C += A*B (BlockPlanner.cpp:1258-1283); variant selects B indexing and the op code
    enum MulKind { MUL, MULNEG, MULT };
    template <MulKind K>
    void mul(NodeId a, NodeId b, NodeId c, int group) {
        if (!a || !b) return;
        if (T[a].level == 0) {
            int ab = T[a].block, bb = T[b].block;
            if (ab > 0 && bb > 0) {
                int cb = leaf_block(c);
                emit(ab, bb, K == MUL ? OP_MUL : (K == MULNEG ? OP_MULNEG : OP_MULT), cb, 0, group);
            }
            return;
        }
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 2; j++)
                for (int k = 0; k < 2; k++) {
                    NodeId as = T[a].sub[i * 2 + k];
                    NodeId bs = T[b].sub[K == MULT ? j * 2 + k : k * 2 + j];
                    if (as && bs) mul<K>(as, bs, child(c, i * 2 + j), group);
                }
    }
    // D = A (BlockPlanner.cpp:1158-1179): leaf op is sub with src = 0
    void copy(NodeId a, NodeId c, int group) {
        if (T[a].level == 0) {
            if (T[a].block > 0) { int cb = leaf_block(c); emit(0, T[a].block, OP_SUB, cb, 0, group); }
            return;
        }
        for (int q = 0; q < 4; q++)
            if (T[a].sub[q]) copy(T[a].sub[q], child(c, q), group);
    }
    // D = -B (BlockPlanner.cpp:1180-1201): leaf op is sub with src2 = 0
    void neg(NodeId b, NodeId c, int group) {
        if (T[b].level == 0) {
            if (T[b].block > 0) { int cb = leaf_block(c); emit(T[b].block, 0, OP_SUB, cb, 0, group); }
            return;
        }
        for (int q = 0; q < 4; q++)
            if (T[b].sub[q]) neg(T[b].sub[q], child(c, q), group);
    }
    // D = A - B (BlockPlanner.cpp:1126-1156): leaf op is sub(src = B, src2 = A)
    void sub(NodeId a, NodeId b, NodeId c, int group) {
        if (T[a].level == 0) {
            if (T[a].block > 0 || T[b].block > 0) {
                int cb = leaf_block(c);
                emit(T[b].block, T[a].block, OP_SUB, cb, 0, group);
            }
            return;
        }
        for (int q = 0; q < 4; q++) {
            NodeId as = T[a].sub[q], bs = T[b].sub[q];
            if (as && bs) sub(as, bs, child(c, q), group);
            else if (as) copy(as, child(c, q), group);
            else if (bs) neg(bs, child(c, q), group);
        }
    }
    // Y = L^-1 for block lower-triangular L; a3 = already available inverse of L00
    // (BlockPlanner.cpp:1019-1054)
    void inv_lower(NodeId l, NodeId y, NodeId a3, int group) {
        if (T[l].level == 0) {
            if (T[l].block > 0) {
                T[y].block = new_block();
                emit(T[l].block, 0, OP_LOWERINV, T[y].block, 0, group);
            }
            return;
        }
        NodeId a = T[l].sub[0], c = T[l].sub[2], d = T[l].sub[3];
        int lv = T[a].level;
        NodeId d1 = T.make(lv);
        inv_lower(d, d1, 0, group);
        T[y].sub[3] = d1;
        NodeId d1c = T.make(lv);
        mul<MUL>(d1, c, d1c, group);
        NodeId a1 = a3;
        if (!a1) {
            a1 = T.make(lv);
            inv_lower(a, a1, 0, group);
        }
        T[y].sub[0] = a1;
        NodeId c21 = T.make(lv);
        mul<MULNEG>(d1c, a1, c21, group);
        T[y].sub[2] = c21;
    }
    // Y = U^-1 (BlockPlanner.cpp:1056-1092)
    void inv_upper(NodeId u, NodeId y, NodeId a3, int group) {
        if (T[u].level == 0) {
            if (T[u].block > 0) {
                T[y].block = new_block();
                emit(T[u].block, 0, OP_UPPERINV, T[y].block, 0, group);
            }
            return;
        }
        NodeId a = T[u].sub[0], b = T[u].sub[1], d = T[u].sub[3];
        int lv = T[a].level;
        NodeId a2 = a3;
        if (!a2) {
            a2 = T.make(lv);
            inv_upper(a, a2, 0, group);
        }
        T[y].sub[0] = a2;
        NodeId a1b = T.make(lv);
        mul<MUL>(a2, b, a1b, group);
        NodeId d1 = T.make(lv);
        inv_upper(d, d1, 0, group);
        T[y].sub[3] = d1;
        NodeId a1bd1 = T.make(lv);
        mul<MULNEG>(a1b, d1, a1bd1, group);
        T[y].sub[1] = a1bd1;
    }
    // A = L*U, no pivoting.  l3/u3 (optional) receive L00^-1 / U00^-1 so the caller can
    // reuse them (BlockPlanner.cpp:865-939)
    void lu(NodeId a, NodeId l, NodeId u, NodeId l3, NodeId u3, int group) {
        if (T[a].level == 0) {
            if (T[a].block > 0) {
                if (T[l].block <= 0) T[l].block = new_block();
                if (T[u].block <= 0) T[u].block = new_block();
                emit(T[a].block, 0, OP_LU, T[l].block, T[u].block, group);
            }
            return;
        }
        int lv = T[a].level - 1;
        NodeId l1 = T.make(lv), u1 = T.make(lv);
        NodeId l2 = l3, u2 = u3;
        if (!l2) { l2 = T.make(lv); u2 = T.make(lv); }
        NodeId hl1 = 0, hu1 = 0;
        if (T[a].level > 1) { hl1 = T.make(lv - 1); hu1 = T.make(lv - 1); }
        lu(T[a].sub[0], l1, u1, hl1, hu1, group);
        inv_lower(l1, l2, hl1, group);
        inv_upper(u1, u2, hu1, group);
        T[l].sub[0] = l1;
        T[u].sub[0] = u1;

        NodeId u12 = T.make(lv);
        NodeId b = T[a].sub[1];
        if (b) {
            mul<MUL>(l2, b, u12, group);
            T[u].sub[1] = u12;
        }
        NodeId dsub = T.make(lv);
        NodeId d = T[a].sub[3];
        NodeId c = T[a].sub[2];
        if (c) {
            NodeId l21 = T.make(lv);
            mul<MUL>(c, u2, l21, group);
            T[l].sub[2] = l21;
            NodeId prod = T.make(lv);
            mul<MUL>(l21, u12, prod, group);
            sub(d, prod, dsub, group);
        } else {
            dsub = d;
        }
        NodeId l1b = T.make(lv), u1b = T.make(lv);
        NodeId dl1 = 0, du1 = 0;
        if (T[a].level > 1) { dl1 = T.make(lv - 1); du1 = T.make(lv - 1); }
        lu(dsub, l1b, u1b, dl1, du1, group);
        T[l].sub[3] = l1b;
        T[u].sub[3] = u1b;
    }
    // A = L*L^T (BlockPlanner.cpp:941-989)
    void llt(NodeId a, NodeId l, NodeId l3, int hn, int group) {
        if (T[a].level == 0) {
            if (T[a].block > 0) {
                if (T[l].block <= 0) T[l].block = new_block();
                emit(T[a].block, 0, OP_LLT, T[l].block, 0, group);
            }
            return;
        }
        int lv = T[a].level - 1;
        int h = hn / 2;
        NodeId l1 = T.make(lv);
        NodeId l2 = l3 ? l3 : T.make(lv);
        NodeId hl1 = 0;
        if (T[a].level > 1) hl1 = T.make(lv - 1);
        llt(T[a].sub[0], l1, hl1, h, group);
        inv_lower(l1, l2, hl1, group);
        T[l].sub[0] = l1;
        NodeId dsub = T.make(lv);
        NodeId d = T[a].sub[3];
        NodeId c = T[a].sub[2];
        if (c) {
            NodeId ca1 = T.make(lv), ca2 = T.make(lv);
            mul<MULT>(c, l2, ca1, group);
            T[l].sub[2] = ca1;
            mul<MULT>(ca1, ca1, ca2, group);
            sub(d, ca2, dsub, group);
        } else {
            dsub = d;
        }
        NodeId l1b = T.make(lv);
        NodeId dl1 = 0;
        if (h > 1) dl1 = T.make(lv - 1);
        llt(dsub, l1b, dl1, h, group);
        T[l].sub[3] = l1b;
    }

    // ---- liveness, staging, ordering (blockPlan, BlockPlanner.cpp:73-374) ----------
    void mark_inputs(NodeId a, std::vector<int32_t>& lastuse) {
        if (!a) return;
        Node& n = T[a];
        if (n.level == 0) {
            if (n.block > 0) { stage[n.block] = 1; laststage[n.block] = 1; lastuse[n.block] = 0; }
            return;
        }
        for (int q = 0; q < 4; q++) mark_inputs(T[a].sub[q], lastuse);
    }
    void mark_outputs(NodeId a, std::vector<int32_t>& lastuse, int marker) {
        if (!a) return;
        Node& n = T[a];
        if (n.level == 0) {
            if (n.block > 0) { laststage[n.block] = marker; lastuse[n.block] = marker; }
            return;
        }
        for (int q = 0; q < 4; q++) mark_outputs(T[a].sub[q], lastuse, marker);
    }

    void block_plan(NodeId saved, NodeId savedu, OpVec& sorted) {
        PhaseTimer pt;
        const int64_t nops = (int64_t)graph.size();
        std::vector<int32_t> lastuse(storage, 0);
        stage.assign(storage, 0);
        laststage.assign(storage, 0);
        stage[0] = 1;
        mark_inputs(blocks, lastuse);
        if (saved) mark_outputs(saved, lastuse, storage);
        if (savedu) mark_outputs(savedu, lastuse, storage);

        // backward liveness sweep (135-176)
        for (int64_t i = nops - 1; i >= 0; i--) {
            const Op& o = graph[i];
            int stg = 0;
            if (o.result > 0) {
                stg = o.result;
                if (o.result2 > 0 && o.result2 > stg) stg = o.result2;
            }
            if (o.result > 0 && o.result2 > 0 && lastuse[o.result] == 0 && lastuse[o.result2] == 0) continue;
            if (o.result > 0 && o.result2 == 0 && lastuse[o.result] == 0) continue;
            if (stg > 0) {
                if (o.src > 0 && lastuse[o.src] < stg) lastuse[o.src] = stg;
                if (o.src2 > 0) {
                    if (o.src2 >= storage) continue;
                    if (lastuse[o.src2] < stg) lastuse[o.src2] = stg;
                }
            }
        }
        pt.lap("liveness sweep");
        // dead-op pruning (178-221): decided here, carried out by the stage scatter below
        int64_t waste = 0;
#pragma omp parallel for schedule(static) reduction(+ : waste) if (nops > 200000)
        for (int64_t i = 0; i < nops; i++) {
            const Op& o = graph[i];
            if (lastuse[o.result] > 0) continue;
            if (o.result2 > 0 && lastuse[o.result2] > 0) continue;
            ++waste;
        }
        const bool prune = planL2 || (uint64_t)waste > (uint64_t)nops / 25;
        const int64_t n = prune ? nops - waste : nops;
        if (prune) seq += 2 * (int)n;  // operation::sett (++seq) + memutil::newoperation (seq++) per kept op
        log << "reduced ops to: " << n << "\n";
        auto dead = [&](const Op& o) { return prune && !(lastuse[o.result] > 0 || (o.result2 > 0 && lastuse[o.result2] > 0)); };
        pt.lap("prune decision");

        // forward ASAP stage numbering over the kept ops (225-267); dead ops get stage -1
        int maxstg = 1;
        for (int64_t i = 0; i < nops; i++) {
            Op& o = graph[i];
            if (dead(o)) { o.stage = -1; continue; }
            int stg = 0;
            if (o.src > 0) {
                stg = stage[o.src] + 1;
                if (stg < 2) stg = 2;
                if (o.src2 > 0 && stage[o.src2] + 1 > stg) stg = stage[o.src2] + 1;
            } else {
                if (o.src2 > 0 && stage[o.src2] + 1 > stg) stg = stage[o.src2] + 1;
                if (stg < 2) stg = 2;
            }
            o.stage = stg;
            if (o.result > 0 && stage[o.result] < stg) stage[o.result] = stg;
            if (o.result2 > 0 && stage[o.result2] < stg) stage[o.result2] = stg;
            if (stg > maxstg) maxstg = stg;
        }
        // every single-result writer moves to its result's final stage (270-276)
#pragma omp parallel for schedule(static) if (nops > 200000)
        for (int64_t i = 0; i < nops; i++) {
            Op& o = graph[i];
            if (o.stage < 0 || o.result2 > 0) continue;
            if (o.result > 0 && stage[o.result] > o.stage) o.stage = stage[o.result];
        }
        pt.lap("stage numbering");
        // total order (stage, group, result, src, seq) (operation.cpp:180-196): a parallel
        // counting sort on the stage (which also drops the dead ops), then one sort per stage
        std::vector<int64_t> start((size_t)maxstg + 2, 0);   // first op of every stage in the sorted list
        {
            int nth = 1;
#ifdef _OPENMP
            nth = omp_get_max_threads();
#endif
            if (nops < (1 << 16)) nth = 1;
            const int64_t chunk = (nops + nth - 1) / nth;
            const size_t S = (size_t)maxstg + 1;
            std::vector<int64_t> cnt((size_t)nth * S, 0);   // [thread][stage]
#pragma omp parallel for schedule(static, 1) num_threads(nth)
            for (int t = 0; t < nth; t++) {
                int64_t* c = cnt.data() + (size_t)t * S;
                const int64_t lo = t * chunk, hi = std::min(nops, lo + chunk);
                for (int64_t i = lo; i < hi; i++)
                    if (graph[i].stage >= 0) c[graph[i].stage]++;
            }
            for (size_t st = 0; st < S; st++) {
                int64_t run = start[st];
                for (int t = 0; t < nth; t++) {
                    int64_t c = cnt[(size_t)t * S + st];
                    cnt[(size_t)t * S + st] = run;
                    run += c;
                }
                start[st + 1] = run;
            }
            sorted.clear();
            sorted.resize(n);
#pragma omp parallel for schedule(static, 1) num_threads(nth)
            for (int t = 0; t < nth; t++) {
                int64_t* pos = cnt.data() + (size_t)t * S;
                const int64_t lo = t * chunk, hi = std::min(nops, lo + chunk);
                for (int64_t i = lo; i < hi; i++)
                    if (graph[i].stage >= 0) sorted[pos[graph[i].stage]++] = graph[i];
            }
            graph.clear();
            pt.lap("stage scatter");
            auto less = [](const Op& x, const Op& y) {
                if (x.group != y.group) return x.group < y.group;
                if (x.result != y.result) return x.result < y.result;
                if (x.src != y.src) return x.src < y.src;
                return x.seq < y.seq;
            };
#pragma omp parallel for schedule(dynamic, 16) if (nops > 200000)
            for (int st = 0; st <= maxstg; st++)
                if (start[st + 1] - start[st] > 1) std::sort(sorted.begin() + start[st], sorted.begin() + start[st + 1], less);
        }
        pt.lap("sort");
        // split stages wider than 8000 ops (335-354): the reference counts ops per stage in one serial scan and
        // bumps the stage number after every 8001st op; with the stage boundaries known that is a prefix sum over
        // the stages.  Then the last reading stage per block (356-371), a maximum.
        {
            std::vector<int32_t> jump((size_t)maxstg + 2, 0);
            for (int st = 0; st <= maxstg; st++) {
                const int64_t c = start[st + 1] - start[st];
                jump[st + 1] = jump[st] + (c > 0 ? (int32_t)((c - 1) / 8001) : 0);
            }
            int32_t* last = laststage.data();
            auto raise = [last](int32_t id, int32_t v) {
                int32_t cur = __atomic_load_n(last + id, __ATOMIC_RELAXED);
                while (cur < v && !__atomic_compare_exchange_n(last + id, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
            };
#pragma omp parallel for schedule(dynamic, 64) if (n > 200000)
            for (int st = 0; st <= maxstg; st++) {
                const int64_t lo = start[st], hi = start[st + 1];
                for (int64_t i = lo; i < hi; i++) {
                    Op& o = sorted[i];
                    o.stage += jump[st] + (int32_t)((i - lo) / 8001);
                    if (o.src > 0) raise(o.src, o.stage);
                    if (o.src2 > 0) raise(o.src2, o.stage);
                }
            }
        }
        pt.lap("split + laststage");
    }

    void collect(NodeId m, int r0, int c0, int n, std::vector<BlockRef>& out) {
        if (!m) return;
        if (T[m].level == 0) {
            if (T[m].block > 0) out.push_back({T[m].block, r0, c0});
            return;
        }
        int h = n / 2;
        for (int q = 0; q < 4; q++) collect(T[m].sub[q], r0 + (q >> 1) * h, c0 + (q & 1) * h, h, out);
    }
};

struct Cell { int row, col; double val; };

}  // namespace

bool BlockValues::alloc_zero(size_t n_blocks) {
    std::free(p_);
    n_ = n_blocks * 4096;
    p_ = n_ ? static_cast<double*>(std::malloc(n_ * sizeof(double))) : nullptr;
    if (n_ && !p_) { n_ = 0; return false; }
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < (int64_t)n_blocks; b++) std::memset(p_ + (size_t)b * 4096, 0, 4096 * sizeof(double));
    return true;
}

bool Plan::dense_inputs(BlockValues& V) const {
    if (!V.alloc_zero(inputs.size())) return false;
    const int64_t n = (int64_t)entry_val.size();
    for (int64_t k = 0; k < n; k++) V[(size_t)(entry_block[k] - 1) * 4096 + entry_pos[k]] = entry_val[k];
    return true;
}

double factor_flops(const OpVec& ops) {
    double f = 0;
    for (const Op& o : ops) {
        switch (o.op) {
            case OP_MUL: case OP_MULNEG: case OP_MULT: f += 524288.0; break;
            case OP_LU: f += 174763.0; break;
            case OP_LLT: case OP_LOWERINV: case OP_UPPERINV: f += 87381.0; break;
            case OP_SUB: f += 4096.0; break;
            default: break;
        }
    }
    return f;
}

int build_plan(const Config& cfg, bool symmetric, const std::vector<int>& idx_i, const std::vector<int>& idx_j,
               const std::vector<double>& vals, Plan& plan, bool keep_values) {
    plan = Plan();
    plan.cfg = cfg;
    plan.symmetric = symmetric;
    if (cfg.mSize < 64) {
        plan.log = "matrix dimension below one 64-row block is not supported by the block planner\n";
        return 1;
    }
    Planner P;
    PhaseTimer pt;
    const size_t nnz = idx_i.size();
    P.mSize = cfg.mSize;
    if (symmetric) P.log << "symmetric\n";

    // ===== coarse pass (solver.cpp:56-82) =================================================
    P.planL2 = true;
    P.blockSize = cfg.blockSizeL2;
    P.blockRows = cfg.blockRowsL2;
    P.storage = 1;
    {
        P.blocks = P.T.make(ilog2(P.blockRows));
        int exp = ilog2(P.blockSize);
        for (size_t k = 0; k < nnz; k++) {          // first-touch order (1482-1490)
            int bi = idx_i[k] >> exp, bj = idx_j[k] >> exp;
            if (P.tree_get(P.blocks, bi, bj) == 0) P.tree_set(P.blocks, bi, bj, P.new_block());
        }
        for (int bi = cfg.mSize / P.blockSize; bi < P.blockRows; bi++)   // identity padding (1520-1528)
            if (P.tree_get(P.blocks, bi, bi) == 0) P.tree_set(P.blocks, bi, bi, P.new_block());
    }
    const int nL2 = P.blockRows;
    NodeId bl = P.T.make(ilog2(nL2)), bu = P.T.make(ilog2(nL2));
    if (symmetric) P.llt(P.blocks, bl, 0, nL2, 0);
    else P.lu(P.blocks, bl, bu, 0, 0, 0);
    plan.coarse_emitted = (int)P.graph.size();
    P.log << "blocks: " << P.blockRows << " blockSize: " << P.blockSize << " inputSize: " << cfg.mSize
          << " extend: " << P.blockRows * P.blockSize << " op count: " << P.graph.size() << " storage: " << P.storage << "\n";
    P.block_plan(bl, symmetric ? 0 : bu, plan.coarse_ops);
    plan.coarse_storage = P.storage;
    pt.lap("coarse pass");

    // ===== expansion to 64-blocks (copyOperatorL2, 1332-1460) ==============================
    const NodeId coarse_blocks = P.blocks;
    const int coarse_storage = P.storage;
    P.planL2 = false;
    P.blockSize = cfg.blockSize;
    P.blockRows = cfg.blockRows;
    const int scaleL2 = cfg.blockSizeL2 / cfg.blockSize;
    const int levelL2 = ilog2(scaleL2);
    NodeId bl2 = P.T.make(ilog2(P.blockRows)), bu2 = P.T.make(ilog2(P.blockRows));
    std::vector<NodeId> L2(coarse_storage, 0);   // coarse block id -> fine sub-quadtree
    P.storage = 1;
    std::vector<BlockRef> in_alloc;              // fine input blocks in allocation order
    {
        P.blocks = P.T.make(ilog2(P.blockRows));
        std::vector<Cell> cells(nnz);
        for (size_t k = 0; k < nnz; k++) { cells[k].row = idx_i[k]; cells[k].col = idx_j[k]; cells[k].val = vals[k]; }
        // same comparator and element count as the reference (operation.cpp:86-92,
        // BlockPlanner.cpp:1498): ties (cells of one block) are ordered by introsort's moves,
        // which matters only for duplicate (row, col) entries -- last write wins.
        std::sort(cells.begin(), cells.end(), [](Cell x, Cell y) {
            if ((x.row >> 6) == (y.row >> 6)) return (x.col >> 6) < (y.col >> 6);
            return (x.row >> 6) < (y.row >> 6);
        });
        pt.lap("  cell sort");
        // ids in first-touch order (BlockPlanner.cpp:1498-1519).  The values are kept as a duplicate-free entry
        // list (input id, position in the 64x64 block, value): the cells of one block are contiguous after the
        // sort, a repeated (row, col) keeps its last value exactly like the reference's scatter (1510).
        auto block_of = [&](int bi, int bj) {
            int id = P.tree_get(P.blocks, bi, bj);
            if (id == 0) {
                id = P.new_block();
                P.tree_set(P.blocks, bi, bj, id);
                in_alloc.push_back({id, bi, bj});
            }
            return id;
        };
        if (keep_values) {
            plan.entry_block.reserve(nnz + 64);
            plan.entry_pos.reserve(nnz + 64);
            plan.entry_val.reserve(nnz + 64);
        }
        {
            int last_bi = -1, last_bj = -1, last_id = 0;
            std::vector<int64_t> seen(4096, -1);      // position -> entry index, valid for entries >= block_first
            int64_t block_first = 0;
            for (size_t k = 0; k < nnz; k++) {
                int bi = cells[k].row >> 6, bj = cells[k].col >> 6;
                if (bi != last_bi || bj != last_bj) {
                    last_id = block_of(bi, bj);
                    last_bi = bi; last_bj = bj;
                    block_first = (int64_t)plan.entry_val.size();
                }
                if (!keep_values) continue;
                const int pos = (cells[k].row & 63) * 64 + (cells[k].col & 63);
                if (seen[pos] >= block_first) { plan.entry_val[seen[pos]] = cells[k].val; continue; }
                seen[pos] = (int64_t)plan.entry_val.size();
                plan.entry_block.push_back(last_id);
                plan.entry_pos.push_back(pos);
                plan.entry_val.push_back(cells[k].val);
            }
        }
        for (int bi = cfg.mSize >> 6; bi < P.blockRows; bi++) block_of(bi, bi);   // identity padding (1520-1528)
        if (keep_values)
            for (int i = cfg.mSize; i < P.blockRows * 64; i++) {
                const int bi = i >> 6, ri = i & 63;
                plan.entry_block.push_back(P.tree_get(P.blocks, bi, bi));
                plan.entry_pos.push_back(ri * 64 + ri);
                plan.entry_val.push_back(1.0);
            }
        // input ids are 1..n_input in allocation order
    }
    pt.lap("fine input blocks");
    // bind coarse input ids to fine sub-quadtrees (matrixZoomSet, 1305-1315)
    {
        struct Fr { NodeId a, d; };
        std::vector<Fr> st;
        st.push_back({coarse_blocks, P.blocks});
        while (!st.empty()) {
            Fr f = st.back();
            st.pop_back();
            Node& a = P.T[f.a];
            if (a.level == 0) {
                if (a.block > 0) L2[a.block] = f.d;
                continue;
            }
            for (int q = 0; q < 4; q++)
                if (a.sub[q]) {
                    NodeId dq = f.d ? P.T[f.d].sub[q] : 0;
                    st.push_back({a.sub[q], dq});
                }
        }
    }
    const OpVec& graphL2 = plan.coarse_ops;
    P.seq += (int)graphL2.size();   // the reference re-allocates every coarse op (1359-1365)
    for (const Op& o : graphL2) {
        if (o.result > 0 && !L2[o.result]) L2[o.result] = P.T.make(levelL2);
        if (o.result2 > 0 && !L2[o.result2]) L2[o.result2] = P.T.make(levelL2);
        const int g = o.seq;
        NodeId s1 = o.src > 0 ? L2[o.src] : 0, s2 = o.src2 > 0 ? L2[o.src2] : 0;
        NodeId r1 = o.result > 0 ? L2[o.result] : 0, r2 = o.result2 > 0 ? L2[o.result2] : 0;
        switch (o.op) {
            case OP_LOWERINV: if (!s1) goto missing; P.inv_lower(s1, r1, 0, g); break;
            case OP_UPPERINV: if (!s1) goto missing; P.inv_upper(s1, r1, 0, g); break;
            case OP_LU:       if (!s1) goto missing; P.lu(s1, r1, r2, 0, 0, g); break;
            case OP_LLT:      if (!s1) goto missing; P.llt(s1, r1, 0, scaleL2, g); break;
            case OP_MUL:      P.mul<Planner::MUL>(s1, s2, r1, g); break;
            case OP_MULT:     P.mul<Planner::MULT>(s1, s2, r1, g); break;
            case OP_MULNEG:   P.mul<Planner::MULNEG>(s1, s2, r1, g); break;
            case OP_SUB:
                if (o.src > 0) {
                    if (!s1) goto missing;
                    if (o.src2 > 0) { if (!s2) goto missing; P.sub(s2, s1, r1, g); }
                    else P.neg(s1, r1, g);
                } else if (o.src2 > 0) {
                    if (!s2) goto missing;
                    P.copy(s2, r1, g);
                }
                break;
            default: break;
        }
        continue;
    missing:
        plan.log = P.log.str() + "planner: coarse block without fine data (structurally singular diagonal?)\n";
        return 2;
    }
    // stitch the fine L/U quadtrees (matrixZoomUpdate, 1316-1330)
    {
        struct Fr { NodeId l, d; };
        auto stitch = [&](NodeId lroot, NodeId droot) {
            std::vector<Fr> st;
            st.push_back({lroot, droot});
            while (!st.empty()) {
                Fr f = st.back();
                st.pop_back();
                int lv = P.T[f.l].level;
                for (int q = 0; q < 4; q++) {
                    NodeId ls = P.T[f.l].sub[q];
                    if (!ls) continue;
                    if (lv == 1) {
                        if (P.T[ls].block > 0) P.T[f.d].sub[q] = L2[P.T[ls].block];
                    } else if (lv > 1) {
                        NodeId nd = P.T.make(lv - 1);
                        P.T[f.d].sub[q] = nd;
                        st.push_back({ls, nd});
                    }
                }
            }
        };
        if (!symmetric) stitch(bu, bu2);
        stitch(bl, bl2);
    }
    pt.lap("fine emission");
    plan.fine_emitted = (int)P.graph.size();
    P.log << "blocks: " << P.blockRows << " blockSize: " << P.blockSize << " inputSize: " << cfg.mSize
          << " extend: " << P.blockRows * P.blockSize << " op count: " << P.graph.size() << " storage: " << P.storage << "\n";
    P.block_plan(bl2, bu2, plan.ops);
    pt.lap("fine block_plan");

    plan.storage = P.storage;
    plan.stage.swap(P.stage);
    plan.laststage.swap(P.laststage);
    P.collect(P.blocks, 0, 0, P.blockRows, plan.inputs);
    P.collect(bl2, 0, 0, P.blockRows, plan.L);
    P.collect(bu2, 0, 0, P.blockRows, plan.U);

    // block coordinates per id, derived from the op semantics (inputs carry theirs)
    plan.brow.assign(plan.storage, -1);
    plan.bcol.assign(plan.storage, -1);
    for (const BlockRef& r : plan.inputs) { plan.brow[r.id] = r.brow; plan.bcol[r.id] = r.bcol; }
    {
        // ops are sorted by stage, and a block's writers precede its readers in stage order
        auto setrc = [&](int id, int r, int c) { if (id > 0 && plan.brow[id] < 0) { plan.brow[id] = r; plan.bcol[id] = c; } };
        for (const Op& o : plan.ops) {
            switch (o.op) {
                case OP_LU: setrc(o.result, plan.brow[o.src], plan.bcol[o.src]); setrc(o.result2, plan.brow[o.src], plan.bcol[o.src]); break;
                case OP_LLT: case OP_LOWERINV: case OP_UPPERINV: setrc(o.result, plan.brow[o.src], plan.bcol[o.src]); break;
                case OP_SUB: { int s = o.src2 > 0 ? o.src2 : o.src; setrc(o.result, plan.brow[s], plan.bcol[s]); break; }
                case OP_MUL: case OP_MULNEG: setrc(o.result, plan.brow[o.src], plan.bcol[o.src2]); break;
                case OP_MULT: setrc(o.result, plan.brow[o.src], plan.brow[o.src2]); break;
                default: break;
            }
        }
    }
    pt.lap("collect + coordinates");
    plan.log = P.log.str();
    return 0;
}

}  // namespace soglu
