// Timed host model of the persistent executor (diagnostics; nothing on the product path calls it).
//
// A discrete-event replay of executor_kernel's scheduling on a compiled TaskGraph: n_ctas workers, each with a
// producer (claims the next slot of the FIFO ready queue, waits for the task published there, streams its operand
// pairs into a 3-stage ring) and a math side (consumes the stages, writes the result, releases the successors).
// Durations are parameters measured on a B200 with the executor's trace option.  The model
// answers "what would the factorisation time be if ..." questions on the CPU: a different task graph (chain
// splitting, row slicing), a faster diagonal kernel, a cheaper hand-over between tasks, another queue policy --
// and gives the critical path under the same durations, so a change can be judged before it is written for the GPU.
//
// Policies: 0 = what the executor does (one FIFO queue, slots pre-claimed with atomicAdd and waited on);
//           1 = ideal list scheduling (an idle CTA takes the ready task with the longest remaining path): the bound
//               any queue discipline can reach with these task durations;
//           2 = two FIFO queues: tasks whose slack is below hi_slack_us go to a high-priority queue that EVERY CTA
//               looks at (non-blocking, one extra CAS) before it pre-claims its next bulk slot; a CTA spinning on
//               its bulk slot serves the high-priority queue meanwhile.
#include "model.h"

#include <algorithm>
#include <deque>
#include <queue>
#include <vector>

namespace soglu {

namespace {
enum : int32_t { EV_CLAIM = 0, EV_FINISH = 1, EV_PUBLISH = 2 };
struct Ev {
    double t;
    int32_t kind;
    int32_t cta;    // EV_PUBLISH: bulk-queue slot (or -1)
    int32_t task;
    bool operator<(const Ev& o) const { return t > o.t; }   // min-heap
};

inline double stage_time(const Task& T, const ModelParams& M) { return model_stage_us(T, M); }
inline int n_stages(const Task& T) { return model_stages(T); }
inline double hop_time(const Task& T, const ModelParams& M) { return model_hop_us(T, M); }
inline int32_t leader_of(const TaskGraph& G, int32_t t) {
    const Task& T = G.tasks[t];
    if (T.type != T_GEMM) return t;
    const int rows16 = std::max(1, (T.flags >> TF_NROWS_SHIFT) & 7);
    return t - ((T.flags >> TF_ROW0_SHIFT) & 3) / rows16;
}
}  // namespace

ModelResult model_executor(const TaskGraph& G, const ModelParams& M) {
    ModelResult R;
    const int64_t nt = (int64_t)G.tasks.size();
    const int nseg = (int)G.seg_begin.size() - 1;
    const int W = M.n_ctas;
    R.n_tasks = nt;
    std::vector<int32_t> dep(nt);
    for (int64_t t = 0; t < nt; t++) dep[t] = G.tasks[t].n_deps;

    // longest paths under the model's durations (successor lists name group leaders; slices are alike)
    const int NG = std::max(1, G.n_owners);
    auto owner_of_task = [&](int32_t t) { return NG > 1 ? (int)G.task_owner[t] : 0; };
    std::vector<float> top(nt, 0.f), bot(nt, 0.f), hop(nt);
    for (int64_t t = 0; t < nt; t++) {
        const Task& T = G.tasks[t];
        double h = hop_time(T, M);
        if (NG > 1) {      // an operand in a peer's HBM
            const uint32_t me = (uint32_t)owner_of_task((int32_t)t);
            const bool two = (T.type == T_GEMM || T.type == T_SUB);
            bool remote = false;
            for (int p = 0, n = n_stages(T); p < n && !remote; p++) {
                const Pair& pr = G.pairs[T.pair_begin + p];
                remote = ((uint32_t)pr.a >> REF_SHIFT) != me || (two && ((uint32_t)pr.b >> REF_SHIFT) != me);
            }
            if (remote) h += M.t_load_remote - M.t_load;
        }
        hop[t] = (float)h;
    }
    const float d_remote = (float)(M.t_release_remote - M.t_release);
    for (int64_t t = 0; t < nt; t++) {
        const Task& T = G.tasks[t];
        const int32_t lead = leader_of(G, (int32_t)t);
        if (lead != t) top[t] = top[lead];
        const int me = owner_of_task((int32_t)t);
        for (int32_t s = T.succ_begin; s < T.succ_end; s++) {
            const int32_t nx = G.succ[s];
            top[nx] = std::max(top[nx], top[t] + hop[t] + (owner_of_task(nx) != me ? d_remote : 0.f));
        }
    }
    for (int64_t t = nt - 1; t >= 0; t--) {
        const Task& T = G.tasks[t];
        const int me = owner_of_task((int32_t)t);
        float m = 0.f;
        for (int32_t s = T.succ_begin; s < T.succ_end; s++) m = std::max(m, bot[G.succ[s]] + (owner_of_task(G.succ[s]) != me ? d_remote : 0.f));
        bot[t] = m + hop[t];
    }
    // composition of the longest chain (first segment's start to the last task): walk it from its head
    if (nt > 0) {
        int64_t cur = 0;
        for (int64_t t = 0; t < nt; t++)
            if (top[t] == 0.f && bot[t] > bot[cur]) cur = t;
        while (true) {
            const Task& T = G.tasks[cur];
            const int kind = (T.type == T_GEMM) ? (((T.flags >> TF_NROWS_SHIFT) & 7) == 4 ? 0 : (((T.flags >> TF_NROWS_SHIFT) & 7) == 2 ? 1 : 2))
                                                : ((T.type == T_LU || T.type == T_LLT) ? 3 : (T.type == T_SUB ? 4 : 5));
            R.chain_tasks[kind]++;
            R.chain_math_us[kind] += n_stages(T) * stage_time(T, M);
            R.chain_overhead_us += hop[cur] - n_stages(T) * stage_time(T, M);
            if (T.type == T_GEMM) R.chain_pairs[kind] += T.n_pairs;
            int64_t nx = -1;
            const int me = owner_of_task((int32_t)cur);
            float best = -1.f;
            for (int32_t e = T.succ_begin; e < T.succ_end; e++) {
                const float v = bot[G.succ[e]] + (owner_of_task(G.succ[e]) != me ? d_remote : 0.f);
                if (v > best) { best = v; nx = G.succ[e]; }
            }
            if (nx < 0) break;
            if (owner_of_task((int32_t)nx) != me) { R.chain_overhead_us += d_remote; R.chain_remote_hops++; }
            cur = nx;
        }
    }
    std::vector<char> hi(nt, 0);
    for (int sg = 0; sg < nseg; sg++) {
        float cp = 0.f;
        for (int32_t t = G.seg_begin[sg]; t < G.seg_begin[sg + 1]; t++) cp = std::max(cp, top[t] + bot[t]);
        R.critical_path_us += cp;
        if (M.policy == 2)      // the compiler's classes when it made any (CompileOptions::hi_slack_us), else by slack here
            for (int32_t t = G.seg_begin[sg]; t < G.seg_begin[sg + 1]; t++) {
                const int32_t lead = leader_of(G, t);
                hi[t] = G.n_hi > 0 ? (G.tasks[t].flags & TF_HI) != 0 : cp - (top[lead] + bot[lead]) < M.hi_slack_us;
                R.n_hi += hi[t];
            }
    }

    // Several GPUs (owner-compiled graph): every GPU has its own CTAs and ready queues; releasing a successor on a peer
    // and loading an operand block from a peer (fetch tasks, unmirrored operands) cost NVLink round trips.
    struct Gpu {
        std::vector<int32_t> queue, waiting;
        std::vector<double> pub;
        int32_t head = 0, tail = 0, n_bulk = 0;
        std::deque<int32_t> hiq;
        std::priority_queue<std::pair<float, int32_t>> ready_pq;
        std::vector<int32_t> idle;
    };
    for (int sg = 0; sg < nseg; sg++) {
        const int32_t t0 = G.seg_begin[sg], t1 = G.seg_begin[sg + 1];
        std::vector<Gpu> gpu(NG);
        for (int32_t t = t0; t < t1; t++) gpu[owner_of_task(t)].n_bulk += !hi[t];
        for (Gpu& g : gpu) { g.queue.assign(g.n_bulk, -1); g.pub.assign(g.n_bulk, 1e300); g.waiting.assign(g.n_bulk, -1); }
        const int WT = W * NG;
        std::vector<char> spinning(WT, 0);
        std::vector<int32_t> own(WT, -1);                                // pre-claimed bulk slot
        for (int32_t k = G.seg_init[2 * sg]; k < G.seg_init[2 * sg + 2]; k++) {
            const int32_t t = G.initial[k];
            Gpu& g = gpu[owner_of_task(t)];
            if (M.policy == 1) g.ready_pq.push({bot[t], t});
            else if (hi[t]) g.hiq.push_back(t);
            else { g.queue[g.tail] = t; g.pub[g.tail] = 0.0; g.tail++; }
        }
        std::vector<double> math_free(WT, 0.0);
        std::vector<double> rel(WT * 3, 0.0);      // when ring stage (it % 3) of a CTA is free again
        std::vector<uint32_t> it(WT, 0);
        std::priority_queue<Ev> pq;
        for (int c = 0; c < WT; c++) pq.push({0.0, EV_CLAIM, c, -1});
        double seg_end = 0.0;
        // begin(): the producer of `cta` knows at time tb which task it got
        auto begin = [&](int cta, int32_t task, double tb) {
            const Task& T = G.tasks[task];
            const int nst = n_stages(T);
            const double ts = stage_time(T, M);
            const uint32_t me = (uint32_t)(cta / W);
            double issue = tb + M.t_desc, m = math_free[cta];
            for (int p = 0; p < nst; p++, it[cta]++) {
                double& r = rel[cta * 3 + it[cta] % 3];
                issue = std::max(issue, r);
                double tl_ = M.t_load;
                if (NG > 1) {
                    const Pair& pr = G.pairs[T.pair_begin + p];
                    const bool two = (T.type == T_GEMM || T.type == T_SUB);
                    if (((uint32_t)pr.a >> REF_SHIFT) != me || (two && ((uint32_t)pr.b >> REF_SHIFT) != me)) { tl_ = M.t_load_remote; R.remote_loads++; }
                }
                m = std::max(m, issue + tl_) + ts;
                r = m;                              // the stage is free again once the math warps consumed it
            }
            m += M.t_epilogue;
            math_free[cta] = m;
            R.busy_us += nst * ts + M.t_epilogue;
            pq.push({m, EV_FINISH, cta, task});
            pq.push({issue + 0.05, EV_CLAIM, cta, -1});    // the producer claims again right after its last issue
        };
        auto pop_idle = [&](Gpu& g) -> int {
            while (!g.idle.empty()) {
                const int c = g.idle.back();
                g.idle.pop_back();
                if (spinning[c]) { spinning[c] = 0; return c; }
            }
            return -1;
        };
        while (!pq.empty()) {
            const Ev e = pq.top();
            pq.pop();
            switch (e.kind) {
                case EV_CLAIM: {
                    const int c = e.cta;
                    Gpu& g = gpu[c / W];
                    if (M.policy == 1) {
                        if (!g.ready_pq.empty()) { const int32_t t = g.ready_pq.top().second; g.ready_pq.pop(); begin(c, t, e.t + M.t_poll_hit); }
                        else { spinning[c] = 1; g.idle.push_back(c); }
                        break;
                    }
                    if (M.policy == 2 && !g.hiq.empty()) {
                        const int32_t t = g.hiq.front();
                        g.hiq.pop_front();
                        begin(c, t, e.t + M.t_poll_hit + M.t_cas);
                        break;
                    }
                    if (own[c] < 0) {
                        if (g.head >= g.n_bulk) {   // no bulk work left: policy 0 exits, policy 2 keeps serving the hi queue
                            if (M.policy == 2) { spinning[c] = 1; g.idle.push_back(c); }
                            break;
                        }
                        own[c] = g.head++;
                        g.waiting[own[c]] = c;
                    }
                    if (g.pub[own[c]] <= e.t) { const int32_t slot = own[c]; own[c] = -1; begin(c, g.queue[slot], e.t + M.t_poll_hit); }
                    else { spinning[c] = 1; if (M.policy == 2) g.idle.push_back(c); }
                    break;
                }
                case EV_PUBLISH: {
                    Gpu& g = gpu[owner_of_task(e.task)];
                    if (M.policy == 1) {
                        const int c = pop_idle(g);
                        if (c >= 0) begin(c, e.task, e.t + M.t_poll);
                        else g.ready_pq.push({bot[e.task], e.task});
                    } else if (hi[e.task]) {
                        const int c = pop_idle(g);  // keeps its pre-claimed bulk slot for afterwards
                        if (c >= 0) begin(c, e.task, e.t + M.t_poll + M.t_cas);
                        else g.hiq.push_back(e.task);
                    } else {
                        const int32_t slot = e.cta;
                        g.pub[slot] = e.t;
                        const int c = g.waiting[slot];
                        if (c >= 0 && spinning[c] && own[c] == slot) { spinning[c] = 0; own[c] = -1; begin(c, g.queue[slot], e.t + M.t_poll); }
                    }
                    break;
                }
                case EV_FINISH: {
                    seg_end = std::max(seg_end, e.t);
                    const Task& T = G.tasks[e.task];
                    const int me = e.cta / W;
                    for (int32_t s = T.succ_begin; s < T.succ_end; s++) {
                        const int32_t nx = G.succ[s];
                        if (--dep[nx] != 0) continue;
                        const int o = owner_of_task(nx);
                        Gpu& g = gpu[o];
                        if (o != me) R.remote_releases++;
                        for (int q = 0, gs = task_group_size(G.tasks[nx]); q < gs; q++) {
                            int32_t slot = -1;
                            if (M.policy != 1 && !hi[nx + q]) { slot = g.tail++; g.queue[slot] = nx + q; }   // atomicAdd(tail) at release time
                            pq.push({e.t + (o == me ? M.t_release : M.t_release_remote), EV_PUBLISH, slot, nx + q});
                        }
                    }
                    break;
                }
            }
        }
        R.makespan_us += seg_end + (NG > 1 ? M.t_launch_dist : M.t_launch);
    }
    return R;
}

}  // namespace soglu
