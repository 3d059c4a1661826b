// TEST INFRASTRUCTURE: host stress test of the shared high-priority ready queue (csrc/device/ready_queue.cuh, executor
// option hi_shared).  Real threads play the CTAs' scheduler lanes and run the SAME claim loop as the kernel on
// std::atomic queue arrays; a task "executes" for a random few hundred nanoseconds and then releases its successors
// exactly like the executor (decrement the successor group's counter; whoever reaches zero claims tail slots with
// one fetch_add and publishes the group's slices).  Random DAGs with random priority classes and group sizes; every
// task must run exactly once, after all its predecessors, and every thread must terminate.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../sparse-operator-graph-lu_b200/csrc/device/ready_queue.cuh"

struct Dag {
    int n = 0;
    std::vector<int> group_leader, group_size, hi;   // per task
    std::vector<std::vector<int>> succ;              // successor group leaders
    std::vector<int> n_deps;                         // per leader: number of predecessor TASKS (each decrements once)
};

struct HostQueues {
    std::atomic<int>* ready[2];
    std::atomic<int>* head[2];
    int head_hi() { return head[0]->load(std::memory_order_relaxed); }
    int ready_hi(int s) { return ready[0][s].load(std::memory_order_acquire); }
    bool cas_head_hi(int h) { int e = h; return head[0]->compare_exchange_strong(e, h + 1, std::memory_order_acq_rel); }
    int claim_lo() { return head[1]->fetch_add(1, std::memory_order_acq_rel); }
    int ready_lo(int s) { return ready[1][s].load(std::memory_order_acquire); }
};

static bool run_case(unsigned seed, int n_threads, int n_tasks, double hi_frac) {
    unsigned long long st = seed * 2654435761ull + 12345;
    auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return (unsigned)(st >> 33); };
    Dag D;
    // tasks in topological order; groups of 1, 2 or 4 consecutive slices share a dependency counter and a class
    while (D.n < n_tasks) {
        const int g = (rnd() % 4 == 0) ? ((rnd() & 1) ? 2 : 4) : 1;
        const int lead = D.n, h = ((rnd() % 1000) < hi_frac * 1000) ? 1 : 0;
        for (int k = 0; k < g; k++) { D.group_leader.push_back(lead); D.group_size.push_back(g); D.hi.push_back(h); D.succ.emplace_back(); D.n_deps.push_back(0); }
        D.n += g;
    }
    for (int t = 0; t < D.n; t++) {
        if (D.group_leader[t] != t) continue;
        // predecessors: up to 3 earlier groups, all slices of each (as the row-split compiler wires them)
        const int np = (t == 0) ? 0 : (int)(rnd() % 4);
        std::vector<int> preds;
        for (int k = 0; k < np; k++) {
            const int window = 1 + (int)(rnd() % 64);
            int p = t - 1 - (int)(rnd() % std::min(t, window));
            p = D.group_leader[p];
            bool dup = false;
            for (int q : preds) dup = dup || q == p;
            if (!dup) preds.push_back(p);
        }
        for (int p : preds)
            for (int k = 0; k < D.group_size[p]; k++) { D.succ[p + k].push_back(t); D.n_deps[t]++; }
    }
    int n_hi = 0;
    for (int t = 0; t < D.n; t++) n_hi += D.hi[t];
    const int n_lo = D.n - n_hi;
    std::vector<std::atomic<int>> ready_hi(std::max(1, n_hi)), ready_lo(std::max(1, n_lo)), dep(D.n), runs(D.n), done(D.n);
    for (auto& r : ready_hi) r.store(-1);
    for (auto& r : ready_lo) r.store(-1);
    std::atomic<int> head[2], tail[2];
    head[0] = head[1] = 0; tail[0] = tail[1] = 0;
    for (int t = 0; t < D.n; t++) { dep[t] = D.n_deps[t]; runs[t] = 0; done[t] = 0; }
    for (int t = 0; t < D.n; t++)
        if (D.group_leader[t] == t && D.n_deps[t] == 0)
            for (int k = 0; k < D.group_size[t]; k++) (D.hi[t] ? ready_hi : ready_lo)[tail[D.hi[t] ? 0 : 1]++].store(t + k);
    std::atomic<int> violations{0};
    auto worker = [&](unsigned wseed) {
        unsigned long long ws = wseed;
        HostQueues q{{ready_hi.data(), ready_lo.data()}, {&head[0], &head[1]}};
        auto issue = [&](int t) {
            if (runs[t].fetch_add(1) != 0) violations++;
            ws = ws * 6364136223846793005ull + 1442695040888963407ull;
            const auto until = std::chrono::steady_clock::now() + std::chrono::nanoseconds((ws >> 40) % 400);
            while (std::chrono::steady_clock::now() < until) {}
            done[t].store(1, std::memory_order_release);
            for (int s : D.succ[t]) {
                if (dep[s].fetch_sub(1, std::memory_order_acq_rel) == 1) {
                    const int cls = D.hi[s] ? 0 : 1, g = D.group_size[s];
                    const int pos = tail[cls].fetch_add(g, std::memory_order_acq_rel);
                    for (int k = 0; k < g; k++) (cls == 0 ? ready_hi : ready_lo)[pos + k].store(s + k, std::memory_order_release);
                }
            }
        };
        soglu::serve_shared_queues(q, n_hi, n_lo, issue);
    };
    std::vector<std::thread> th;
    for (int w = 0; w < n_threads; w++) th.emplace_back(worker, seed * 977u + w);
    for (auto& t : th) t.join();
    for (int t = 0; t < D.n; t++) if (runs[t] != 1) violations++;
    if (head[0] != n_hi || head[1] < n_lo) violations++;
    return violations == 0;
}

int main() {
    int bad = 0, cases = 0;
    for (unsigned seed = 1; seed <= 12; seed++)
        for (double hf : {0.0, 0.05, 0.5, 1.0}) {
            const int threads = 2 + (seed % 3) * 7;      // 2, 9, 16 scheduler lanes
            cases++;
            if (!run_case(seed, threads, 20000, hf)) { std::printf("FAILED seed %u hi_frac %.2f threads %d\n", seed, hf, threads); bad++; }
        }
    std::printf("%d cases, %d failed\n%s\n", cases, bad, bad ? "FAIL" : "OK");
    return bad ? 1 : 0;
}
