// Cost model of one task hop (microseconds), used by the task compiler's slack estimates (row split of near-critical
// GEMM tasks) and by the static execution order (earliest / latest start times).  Durations measured with the executor's
// trace option and tools/diag_bench.py on a B200 -- on the executor with the dynamic ready queue; the static-order executor's
// hops are cheaper (signal 1.9 us, detection 0.8 us, first operands 1.6 us: profiles/r02_static_order.md), but split_slack and
// order_alpha were tuned by measurement WITH these constants, so they stay until both are re-measured together.
#pragma once
#include "tasks.h"

namespace soglu {

struct ModelParams {        // microseconds, measured with the executor's trace option on a B200
    double t_pair = 2.38;          // one 64x64x64 product, whole block, 8 math warps of one SM
    double t_pair_half = 1.25;     // 32-row slice
    double t_pair_quarter = 0.72;  // 16-row slice
    double t_lu_fused = 11.7;      // lu + both inverses, blocked kernel (23 k cycles incl. write-out)
    double t_lu = 9.1;
    double t_llt_fused = 10.5;
    double t_inv = 7.0;
    double t_sub = 0.6;
    double t_epilogue = 0.4;       // result write-back + barrier
    double t_release = 1.5;        // store visibility + dependency atomics + publication
    double t_poll = 1.5;           // a spinning scheduler sees the published task
    double t_poll_hit = 0.8;       // the task was already there when the slot was claimed
    double t_desc = 0.6;           // task record fetch
    double t_load = 1.3;           // bulk copy of one operand pair into shared memory
};

// ---- the cost model, shared by the executor model and by the task compiler's slack / chain estimates ---------------
// time of ONE operand stage of a task; rows16 overrides the slice height of a GEMM task (4 = whole block, 2, 1)
inline double model_stage_us(const Task& T, const ModelParams& M, int rows16 = -1) {
    switch (T.type) {
        case T_GEMM: {
            const int r16 = rows16 > 0 ? rows16 : (T.flags >> TF_NROWS_SHIFT) & 7;
            return r16 == 4 ? M.t_pair : (r16 == 2 ? M.t_pair_half : M.t_pair_quarter);
        }
        case T_SUB: return M.t_sub;
        case T_LU: return (T.flags & (TF_LINV | TF_UINV)) ? M.t_lu_fused : M.t_lu;
        case T_LLT: return (T.flags & TF_LINV) ? M.t_llt_fused : M.t_lu;
        case T_LOWERINV: case T_UPPERINV: return M.t_inv;
        default: return 1.0;
    }
}
inline int model_stages(const Task& T) { return T.type == T_GEMM ? T.n_pairs : 1; }
inline double model_in_us(const ModelParams& M) { return M.t_desc + M.t_load; }                    // pick-up to first math
inline double model_out_us(const ModelParams& M) { return M.t_epilogue + M.t_release + M.t_poll; }  // last math to the successor's pick-up
// one dependent hop through the task: fetch, first operands, math, write-back, release, pick-up by the successor
inline double model_hop_us(const Task& T, const ModelParams& M, int rows16 = -1) {
    return model_in_us(M) + model_stages(T) * model_stage_us(T, M, rows16) + model_out_us(M);
}

}  // namespace soglu
