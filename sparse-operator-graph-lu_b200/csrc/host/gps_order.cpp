// GPS (Gibbs-Poole-Stockmeyer-like) bandwidth-reducing ordering.
//
// Re-implementation of the reference's GOrder (GPSOrder.cpp:41-635) on flat CSR arrays.
// The permutation must equal the reference's bit for bit, so the *decision sequence* is
// kept: which BFS roots are tried, in which order candidates are examined, how ties are
// broken.  Two places are sensitive to the sort algorithm because their keys tie
// (nodeSum ranks, GPSOrder.cpp:360, 378, 410): there we call libstdc++ std::sort on a
// sequence of the same length with a comparator returning the same booleans, which makes
// introsort perform the same moves as in the reference binary.
//
// Differences in mechanics (not in results): adjacency is CSR instead of per-row
// std::vector; the per-row std::set "tbd" is a sorted neighbour list with erase flags;
// candidate level structures are summarised (depth, width, last level) and only the
// winner is rebuilt in full, so 500 BFS trees are never resident at once.
#include "soglu_host.h"

#include <algorithm>
#include <set>
#include <sstream>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace soglu {
namespace {

constexpr int LIST100 = 100;
constexpr int LIST500 = 500;

struct Graph {
    int n = 0;
    std::vector<int64_t> ptr;   // n+1
    std::vector<int> adj;       // sorted, unique, no self loops (GPSOrder.cpp:279-338)
    std::vector<int> nodesum;   // sum of neighbour degrees (GPSOrder.cpp:452-459)
    bool pattern_symmetric = false;
    int deg(int v) const { return (int)(ptr[v + 1] - ptr[v]); }
};

void setup_edges(const std::vector<int>& ii, const std::vector<int>& jj, int n, Graph& g) {
    g.n = n;
    size_t nnz = ii.size();
    std::vector<int64_t> cnt(n + 1, 0);
    int64_t shadow = 0;
    for (size_t k = 0; k < nnz; k++)
        if (ii[k] != jj[k]) { cnt[ii[k] + 1]++; cnt[jj[k] + 1]++; shadow++; }
    for (int v = 0; v < n; v++) cnt[v + 1] += cnt[v];
    std::vector<int> raw(cnt[n]);
    std::vector<int64_t> fill(cnt.begin(), cnt.end() - 1);
    for (size_t k = 0; k < nnz; k++)
        if (ii[k] != jj[k]) { raw[fill[ii[k]]++] = jj[k]; raw[fill[jj[k]]++] = ii[k]; }
    g.ptr.assign(n + 1, 0);
    // sort + unique per row (order inside a row is ascending, so the two passes of the
    // reference collapse to one)
#pragma omp parallel for schedule(dynamic, 4096)
    for (int v = 0; v < n; v++) {
        std::sort(raw.begin() + cnt[v], raw.begin() + cnt[v + 1]);
        int64_t u = std::unique(raw.begin() + cnt[v], raw.begin() + cnt[v + 1]) - (raw.begin() + cnt[v]);
        g.ptr[v + 1] = u;
    }
    for (int v = 0; v < n; v++) g.ptr[v + 1] += g.ptr[v];
    g.adj.resize(g.ptr[n]);
#pragma omp parallel for schedule(static)
    for (int v = 0; v < n; v++)
        std::copy(raw.begin() + cnt[v], raw.begin() + cnt[v] + (g.ptr[v + 1] - g.ptr[v]), g.adj.begin() + g.ptr[v]);
    g.pattern_symmetric = (shadow == g.ptr[n]);
    g.nodesum.resize(n);
#pragma omp parallel for schedule(static)
    for (int v = 0; v < n; v++) {
        int s = 0;
        for (int64_t e = g.ptr[v]; e < g.ptr[v + 1]; e++) s += g.deg(g.adj[e]);
        g.nodesum[v] = s;
    }
}

// rooted level structure; levels appended for further components in ascending order of
// their smallest unseen node (getLevelList + addMissing, GPSOrder.cpp:185-277)
struct Levels {
    std::vector<int> order;      // nodes, level by level
    std::vector<int64_t> off;    // level l = order[off[l] .. off[l+1])
    int depth() const { return (int)off.size() - 1; }
    int width() const {
        int64_t w = 0;
        for (size_t l = 0; l + 1 < off.size(); l++) w = std::max(w, off[l + 1] - off[l]);
        return (int)w;
    }
};

void bfs_levels(const Graph& g, int root, Levels& out, std::vector<char>& seen) {
    int n = g.n;
    seen.assign(n, 0);
    out.order.clear();
    out.order.reserve(n);
    out.off.clear();
    int next_unseen = 0;
    int start = root;
    while (true) {
        out.off.push_back((int64_t)out.order.size());
        out.order.push_back(start);
        seen[start] = 1;
        int64_t lb = out.off.back(), le = (int64_t)out.order.size();
        while (true) {
            for (int64_t p = lb; p < le; p++) {
                int row = out.order[p];
                for (int64_t e = g.ptr[row]; e < g.ptr[row + 1]; e++) {
                    int col = g.adj[e];
                    if (seen[col]) continue;
                    seen[col] = 1;
                    out.order.push_back(col);
                }
            }
            if ((int64_t)out.order.size() == le) break;
            out.off.push_back(le);
            lb = le;
            le = (int64_t)out.order.size();
        }
        if ((int)out.order.size() >= n) break;
        while (next_unseen < n && seen[next_unseen]) next_unseen++;
        if (next_unseen >= n) break;
        start = next_unseen;
    }
    out.off.push_back((int64_t)out.order.size());
}

struct Summary {
    int depth = 0, width = 0;
    std::vector<int> last;
    bool valid = false;
};

void summarise(const Levels& lv, Summary& s) {
    s.depth = lv.depth();
    s.width = lv.width();
    s.last.assign(lv.order.begin() + lv.off[s.depth - 1], lv.order.begin() + lv.off[s.depth]);
    s.valid = true;
}

void bfs_many(const Graph& g, const int* roots, Summary* out, int count) {
#pragma omp parallel
    {
        Levels lv;
        std::vector<char> seen;
#pragma omp for schedule(dynamic, 1)
        for (int t = 0; t < count; t++) {
            bfs_levels(g, roots[t], lv, seen);
            summarise(lv, out[t]);
        }
    }
}

struct Pos { int value, index; };
inline bool pos_less(Pos x, Pos y) { return x.value < y.value; }

// numbering of one more level (reorderOneMoreLevel, GPSOrder.cpp:340-392)
struct Numbering {
    const Graph& g;
    std::vector<int> newnum;     // old -> new  (reference: newOrd)
    std::vector<int> oldof;      // new -> old  (reference: reverseOrd)
    std::vector<char> erased;    // per adjacency entry: removed from the row's tbd set
    std::vector<int> remaining;  // per row: entries left in tbd
    explicit Numbering(const Graph& gg) : g(gg), newnum(gg.n, -1), oldof(gg.n, -1), erased(gg.adj.size(), 0), remaining(gg.n) {
        for (int v = 0; v < g.n; v++) remaining[v] = g.deg(v);
    }
    void tbd_erase(int row, int what) {
        const int* b = g.adj.data() + g.ptr[row];
        const int* e = g.adj.data() + g.ptr[row + 1];
        const int* p = std::lower_bound(b, e, what);
        if (p != e && *p == what) {
            int64_t k = p - g.adj.data();
            if (!erased[k]) { erased[k] = 1; remaining[row]--; }
        }
    }
    void assign(int node, int& cursor) {
        newnum[node] = cursor;
        oldof[cursor] = node;
        cursor++;
        for (int64_t e = g.ptr[node]; e < g.ptr[node + 1]; e++) {
            int col = g.adj[e];
            if (newnum[col] >= 0 && remaining[col] > 0) tbd_erase(col, node);
        }
    }
};

}  // namespace

void gps_reorder(int dim, std::vector<int>& idx_i, std::vector<int>& idx_j, std::vector<double>& b,
                 const Config& cfg, Ordering& ord) {
    Graph g;
    setup_edges(idx_i, idx_j, dim, g);

    // getLeastConnected (GPSOrder.cpp:471-484)
    int startnode = 0;
    {
        int count = dim;
        for (int i = 0; i < dim; i++) {
            int d = g.deg(i);
            if (d < count) {
                if (d == 0) continue;
                startnode = i;
                count = d;
            }
        }
    }
    Levels work;
    std::vector<char> seen;
    bfs_levels(g, startnode, work, seen);
    Summary levels;
    summarise(work, levels);

    std::vector<int> ppnodeset(LIST500);
    std::vector<Summary> ppsum(LIST500);
    int ppnodecount = 0;
    std::set<int> Gset;
    for (int v : levels.last) Gset.insert(v);
    {
        int d0 = g.deg(startnode);
        for (int v = 0; v < dim; v++)
            if (g.deg(v) == d0) Gset.insert(v);
    }
    // pseudo-peripheral search (GPSOrder.cpp:521-582)
    while (true) {
        const std::vector<int> last = levels.last;
        int lastsize = (int)last.size();
        int lastskip = 0;
        if (lastsize > LIST100) {
            lastskip = (lastsize + LIST100 - 1) / LIST100;
            lastskip = (lastskip / 2 + 1) * 2 - 1;
        }
        ppnodecount = 0;
        int maxsize = levels.depth;
        for (int t = 0; t < lastsize; t++) {
            if (lastskip > 0 && t % lastskip != 0) continue;
            ppnodeset[ppnodecount] = last[t];
            ppsum[ppnodecount].valid = false;
            ppnodecount++;
        }
        bfs_many(g, ppnodeset.data(), ppsum.data(), ppnodecount);
        bool updated = false;
        for (int t = 0; t < ppnodecount; t++) {
            const Summary& lvl = ppsum[t];
            if (lvl.depth < maxsize) continue;
            for (int v : lvl.last) Gset.insert(v);
            if (lvl.depth > levels.depth) {
                levels = lvl;
                startnode = ppnodeset[t];
                updated = true;
                ppnodecount = 0;
                break;
            }
        }
        if (!updated) break;
    }

    // widen the candidate set with Gset (GPSOrder.cpp:585-612)
    int vdegree = g.deg(startnode);
    Gset.insert(dim - 1);
    Gset.insert(0);
    int nodeindex = ppnodecount;
    int nodecnt = ppnodecount;
    int nodeskip = 0;
    int cap = LIST500 - ppnodecount;
    if (cap > 0) {
        nodeskip = (int)((Gset.size() + cap - 1) / cap);
        nodeskip = (nodeskip / 2 + 1) * 2 - 1;
    }
    for (int ss : Gset) {
        if (g.deg(ss) > vdegree && ss != dim - 1 && ss != 0) continue;
        bool found = false;
        for (int q = 0; q < nodecnt; q++)
            if (ppnodeset[q] == ss) { found = true; break; }
        nodeindex++;
        if (nodeskip > 1 && nodeindex % nodeskip != 0) continue;
        if (found) continue;
        if (nodecnt >= LIST500) break;
        ppnodeset[nodecnt++] = ss;
    }
    bfs_many(g, ppnodeset.data() + ppnodecount, ppsum.data() + ppnodecount, nodecnt - ppnodecount);

    // first candidate maximising depth - width (GPSOrder.cpp:614-624)
    int minmax = -dim, best = -1;
    for (int i = 0; i < nodecnt; i++) {
        if (!ppsum[i].valid) continue;
        int eff = ppsum[i].depth - ppsum[i].width;
        if (minmax < eff) { minmax = eff; best = i; }
    }
    if (best < 0) {  // cannot happen for dim >= 1; keep the start node's structure
        ppnodeset[0] = startnode;
        best = 0;
    }
    bfs_levels(g, ppnodeset[best], work, seen);
    ord.levels = work.depth();
    ord.width = work.width();
    ord.lastLevelCount = (int)(work.off[work.depth()] - work.off[work.depth() - 1]);
    ord.accounted = (int)work.order.size();
    ord.startNode = work.order[0];

    // updatewithlevel (GPSOrder.cpp:394-450)
    Numbering num(g);
    std::vector<Pos> rank;
    int cursor = 0;
    {
        for (int64_t p = work.off[0]; p < work.off[1]; p++) rank.push_back({g.nodesum[work.order[p]], work.order[p]});
        std::sort(rank.begin(), rank.end(), pos_less);
        for (const Pos& r : rank) num.assign(r.index, cursor);
    }
    std::vector<Pos> prevsorted;
    for (int lev = 1; lev < work.depth(); lev++) {
        int64_t pb = work.off[lev - 1], pe = work.off[lev], cb = work.off[lev], ce = work.off[lev + 1];
        int start = cursor;
        prevsorted.clear();
        for (int64_t p = pb; p < pe; p++) prevsorted.push_back({num.newnum[work.order[p]], work.order[p]});
        // the reference sorts descending and consumes from the back: ascending new number
        // (keys are unique, so any sort gives the same sequence)
        std::sort(prevsorted.begin(), prevsorted.end(), pos_less);
        for (const Pos& pw : prevsorted) {
            int w = pw.index;
            if (num.remaining[w] == 0) continue;
            rank.clear();
            for (int64_t e = g.ptr[w]; e < g.ptr[w + 1]; e++) {
                if (num.erased[e]) continue;
                int node = g.adj[e];
                if (num.newnum[node] < 0) rank.push_back({g.nodesum[node], node});
            }
            if (rank.empty()) continue;
            std::sort(rank.begin(), rank.end(), pos_less);
            for (const Pos& r : rank) num.assign(r.index, cursor);
        }
        if (cursor - start != (int)(ce - cb)) {
            rank.clear();
            for (int64_t p = cb; p < ce; p++)
                if (num.newnum[work.order[p]] < 0) rank.push_back({g.nodesum[work.order[p]], work.order[p]});
            std::sort(rank.begin(), rank.end(), pos_less);
            for (const Pos& r : rank) num.assign(r.index, cursor);
        }
    }

    size_t nnz = idx_i.size();
    for (size_t k = 0; k < nnz; k++) {
        idx_i[k] = num.newnum[idx_i[k]];
        idx_j[k] = num.newnum[idx_j[k]];
    }
    int ext = cfg.blockRows * cfg.blockSize;
    std::vector<double> nb(ext, 1.0);
    for (int i = 0; i < dim; i++) nb[i] = b[num.oldof[i]];
    b.swap(nb);
    ord.newOrder = num.oldof;       // naming trap of the reference (GPSOrder.cpp:448-449)
    ord.reverseOrder = num.newnum;
}

// In-block re-sort by diagonal magnitude (GPSOrder.cpp:55-150).  Identity for inputs
// whose diagonals are all non-zero, but the composition with the GPS permutation is
// still applied exactly as the reference does.
void sort_in_block(int dim, std::vector<int>& idx_i, std::vector<int>& idx_j, const std::vector<double>& vals,
                   std::vector<double>& b, Ordering& ord) {
    const double TINY = 1e-8;
    const int bs = 64;
    std::vector<double> tmpv(dim, 0.0);
    std::vector<int> tmpi(dim, -1);
    size_t nnz = idx_i.size();
    for (size_t k = 0; k < nnz; k++)
        if (idx_i[k] == idx_j[k]) { tmpv[idx_i[k]] = vals[k]; tmpi[idx_i[k]] = idx_i[k]; }
    for (int i = 0; i < dim; i++)
        if (tmpi[i] == -1) tmpi[i] = i;
    for (int i = 0; i < dim; i += bs) {
        if (i + bs >= dim) continue;
        int step = bs / 2;
        int zcount = 0;
        for (int j = 0; j < step; j++) {
            if (tmpv[i + j] > TINY || tmpv[i + j] < -TINY) continue;
            zcount++;
        }
        if (zcount == 0) continue;
        while (step > 0) {
            for (int j = 0; j < step; j++) {
                int k = i + j, m = i + j + step;
                if (tmpv[m] > tmpv[k]) { std::swap(tmpv[m], tmpv[k]); std::swap(tmpi[m], tmpi[k]); }
            }
            step /= 2;
        }
    }
    std::vector<int> tmpn(dim, -1);
    for (int i = 0; i < dim; i++) tmpn[tmpi[i]] = i;
    for (size_t k = 0; k < nnz; k++) { idx_i[k] = tmpn[idx_i[k]]; idx_j[k] = tmpn[idx_j[k]]; }
    std::vector<double> bn(dim);
    for (int i = 0; i < dim; i++) bn[i] = b[tmpi[i]];
    for (int i = 0; i < dim; i++) b[i] = bn[i];
    if (ord.newOrder.empty()) {
        ord.newOrder.resize(dim);
        ord.reverseOrder.resize(dim);
        for (int i = 0; i < dim; i++) ord.newOrder[i] = ord.reverseOrder[i] = i;
    }
    std::vector<int> norder(dim);
    for (int i = 0; i < dim; i++) norder[i] = ord.newOrder[tmpi[i]];
    for (int i = 0; i < dim; i++) { ord.newOrder[i] = norder[i]; ord.reverseOrder[norder[i]] = i; }
}

}  // namespace soglu
