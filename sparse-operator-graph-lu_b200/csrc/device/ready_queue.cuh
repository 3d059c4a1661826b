// Claim loop of the shared high-priority ready queue (executor option hi_shared), written against a small queue
// interface so that the SAME code runs in the executor's scheduler lane (PTX loads / atomics on the queue arrays
// in HBM) and in tests/emu/emu_queue.cpp (std::atomic, real host threads as CTAs).
//
// Queue 0 (high priority) is never waited on: a CTA takes the entry at its head only if it is already published,
// with a compare-and-swap on the head.  Queue 1 (bulk) keeps the claim-then-wait discipline of the default
// executor: one atomicAdd claims the next slot, the CTA returns to it until a finishing CTA publishes a task
// there -- and serves queue 0 in the meantime.  The loop ends when the CTA's bulk claim falls beyond the queue and
// the high-priority head has reached its end (heads only grow, every published entry is claimed exactly once).
//
// Q provides: int head_hi(); int ready_hi(int slot); bool cas_head_hi(int expected); int claim_lo();
//             int ready_lo(int slot)   (ready_* return -1 while the slot is not published; acquire semantics)
#pragma once

#if defined(__CUDACC__)
#define RQ_FN __device__ __forceinline__
#else
#define RQ_FN inline
#endif

namespace soglu {

template <class Q, class Issue>
RQ_FN void serve_shared_queues(Q& q, int n_hi, int n_lo, Issue&& issue) {
    int mine = -1;            // pre-claimed bulk slot
    bool lo_left = n_lo > 0;
    while (true) {
        const int h = q.head_hi();
        if (h < n_hi) {
            const int t = q.ready_hi(h);
            if (t >= 0) {
                if (q.cas_head_hi(h)) issue(t);
                continue;
            }
        }
        if (mine < 0 && lo_left) {
            mine = q.claim_lo();
            if (mine >= n_lo) { mine = -1; lo_left = false; }
        }
        if (mine >= 0) {
            const int t = q.ready_lo(mine);
            if (t >= 0) { mine = -1; issue(t); }
        } else if (h >= n_hi) {
            break;            // both queues exhausted
        }
    }
}

}  // namespace soglu
