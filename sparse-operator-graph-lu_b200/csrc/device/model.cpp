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

inline double stage_time(const Task& T, const ModelParams& M) {
    switch (T.type) {
        case T_GEMM: {
            const int r16 = (T.flags >> TF_NROWS_SHIFT) & 7;
            return r16 == 4 ? M.t_pair : (r16 == 2 ? M.t_pair_half : M.t_pair_quarter);
        }
        case T_SUB: return M.t_sub;
        case T_LU: return (T.flags & (TF_LINV | TF_UINV)) ? M.t_lu_fused : M.t_lu;
        case T_LLT: return (T.flags & TF_LINV) ? M.t_llt_fused : M.t_lu;
        case T_LOWERINV: case T_UPPERINV: return M.t_inv;
        default: return 1.0;
    }
}
inline int n_stages(const Task& T) { return T.type == T_GEMM ? T.n_pairs : 1; }
// one dependent hop through the task: fetch, first operands, math, write-back, release, pick-up by the successor
inline double hop_time(const Task& T, const ModelParams& M) {
    return M.t_desc + M.t_load + n_stages(T) * stage_time(T, M) + M.t_epilogue + M.t_release + M.t_poll;
}
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
    std::vector<float> top(nt, 0.f), bot(nt, 0.f);
    for (int64_t t = 0; t < nt; t++) {
        const Task& T = G.tasks[t];
        const int32_t lead = leader_of(G, (int32_t)t);
        if (lead != t) top[t] = top[lead];
        const float d = (float)hop_time(T, M);
        for (int32_t s = T.succ_begin; s < T.succ_end; s++) top[G.succ[s]] = std::max(top[G.succ[s]], top[t] + d);
    }
    for (int64_t t = nt - 1; t >= 0; t--) {
        const Task& T = G.tasks[t];
        float m = 0.f;
        for (int32_t s = T.succ_begin; s < T.succ_end; s++) m = std::max(m, bot[G.succ[s]]);
        bot[t] = m + (float)hop_time(T, M);
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

    for (int sg = 0; sg < nseg; sg++) {
        const int32_t t0 = G.seg_begin[sg], t1 = G.seg_begin[sg + 1];
        int32_t n_bulk = 0;
        for (int32_t t = t0; t < t1; t++) n_bulk += !hi[t];
        std::vector<int32_t> queue(n_bulk, -1);          // the (bulk) FIFO ready queue
        std::vector<double> pub(n_bulk, 1e300);          // when the entry becomes visible
        std::vector<int32_t> waiting(n_bulk, -1);        // CTA that pre-claimed the slot
        int32_t head = 0, tail = 0;
        std::deque<int32_t> hiq;                                         // policy 2
        std::priority_queue<std::pair<float, int32_t>> ready_pq;         // policy 1: (remaining path, task)
        std::vector<int32_t> idle;                                       // policies 1, 2: spinning CTAs (lazy deletion)
        std::vector<char> spinning(W, 0);
        std::vector<int32_t> own(W, -1);                                 // pre-claimed bulk slot
        for (int32_t k = G.seg_init[2 * sg]; k < G.seg_init[2 * sg + 2]; k++) {
            const int32_t t = G.initial[k];
            if (M.policy == 1) ready_pq.push({bot[t], t});
            else if (hi[t]) hiq.push_back(t);
            else { queue[tail] = t; pub[tail] = 0.0; tail++; }
        }
        std::vector<double> math_free(W, 0.0);
        std::vector<double> rel(W * 3, 0.0);       // when ring stage (it % 3) of a CTA is free again
        std::vector<uint32_t> it(W, 0);
        std::priority_queue<Ev> pq;
        for (int c = 0; c < W; c++) pq.push({0.0, EV_CLAIM, c, -1});
        double seg_end = 0.0;
        // begin(): the producer of `cta` knows at time tb which task it got
        auto begin = [&](int cta, int32_t task, double tb) {
            const Task& T = G.tasks[task];
            const int nst = n_stages(T);
            const double ts = stage_time(T, M);
            double issue = tb + M.t_desc, m = math_free[cta];
            for (int p = 0; p < nst; p++, it[cta]++) {
                double& r = rel[cta * 3 + it[cta] % 3];
                issue = std::max(issue, r);
                m = std::max(m, issue + M.t_load) + ts;
                r = m;                              // the stage is free again once the math warps consumed it
            }
            m += M.t_epilogue;
            math_free[cta] = m;
            R.busy_us += nst * ts + M.t_epilogue;
            pq.push({m, EV_FINISH, cta, task});
            pq.push({issue + 0.05, EV_CLAIM, cta, -1});    // the producer claims again right after its last issue
        };
        auto pop_idle = [&]() -> int {
            while (!idle.empty()) {
                const int c = idle.back();
                idle.pop_back();
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
                    if (M.policy == 1) {
                        if (!ready_pq.empty()) { const int32_t t = ready_pq.top().second; ready_pq.pop(); begin(c, t, e.t + M.t_poll_hit); }
                        else { spinning[c] = 1; idle.push_back(c); }
                        break;
                    }
                    if (M.policy == 2 && !hiq.empty()) {
                        const int32_t t = hiq.front();
                        hiq.pop_front();
                        begin(c, t, e.t + M.t_poll_hit + M.t_cas);
                        break;
                    }
                    if (own[c] < 0) {
                        if (head >= n_bulk) {       // no bulk work left: policy 0 exits, policy 2 keeps serving the hi queue
                            if (M.policy == 2) { spinning[c] = 1; idle.push_back(c); }
                            break;
                        }
                        own[c] = head++;
                        waiting[own[c]] = c;
                    }
                    if (pub[own[c]] <= e.t) { const int32_t slot = own[c]; own[c] = -1; begin(c, queue[slot], e.t + M.t_poll_hit); }
                    else { spinning[c] = 1; if (M.policy == 2) idle.push_back(c); }
                    break;
                }
                case EV_PUBLISH: {
                    if (M.policy == 1) {
                        const int c = pop_idle();
                        if (c >= 0) begin(c, e.task, e.t + M.t_poll);
                        else ready_pq.push({bot[e.task], e.task});
                    } else if (hi[e.task]) {
                        const int c = pop_idle();   // keeps its pre-claimed bulk slot for afterwards
                        if (c >= 0) begin(c, e.task, e.t + M.t_poll + M.t_cas);
                        else hiq.push_back(e.task);
                    } else {
                        const int32_t slot = e.cta;
                        pub[slot] = e.t;
                        const int c = waiting[slot];
                        if (c >= 0 && spinning[c] && own[c] == slot) { spinning[c] = 0; own[c] = -1; begin(c, queue[slot], e.t + M.t_poll); }
                    }
                    break;
                }
                case EV_FINISH: {
                    seg_end = std::max(seg_end, e.t);
                    const Task& T = G.tasks[e.task];
                    for (int32_t s = T.succ_begin; s < T.succ_end; s++) {
                        const int32_t nx = G.succ[s];
                        if (--dep[nx] != 0) continue;
                        for (int q = 0, g = task_group_size(G.tasks[nx]); q < g; q++) {
                            int32_t slot = -1;
                            if (M.policy != 1 && !hi[nx + q]) { slot = tail++; queue[slot] = nx + q; }   // atomicAdd(tail) at release time
                            pq.push({e.t + M.t_release, EV_PUBLISH, slot, nx + q});
                        }
                    }
                    break;
                }
            }
        }
        R.makespan_us += seg_end + M.t_launch;
    }
    return R;
}

}  // namespace soglu
