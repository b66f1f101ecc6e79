// scalars.h — layout of the device scalar block ctx->d_red (mirrored in pinned host memory ctx->h_red)
#pragma once
enum {
    S_DELTA = 0,     // delta   = <r, s>            (solverCG.h:92)
    S_DELTA0 = 1,    // delta0                        (solverCG.h:91)
    S_RS = 2,        // <r, s_new> deposited by the c2r epilogue
    S_DKD = 3,       // <d, K d>                      (solverCG.h:105)
    S_BETA = 4,      // fmax(0, (delta - deltamid)/delta0)   (solverCG.h:94)
    S_FREEZE = 5,    // batched solves: != 0 once this lane has converged — k_cg_update leaves its r and u alone
    S_L1 = 8,        // sum |r|      -- the next four are written together by k_cg_update / k_reduce4
    S_L2SQ = 9,      // sum r^2
    S_DELTAMID = 10, // <r, s>                        (solverCG.h:86)
    S_LINF = 11,     // max |r|
    S_GEN = 12,      // 4 generic slots (sum|a|, sum a^2, sum a*b, max|a|)
    S_STRESS = 16,   // 9 slots: sum of element-averaged stress
    S_LS = 25,       // 4 slots: the generic block of a second reduction (<r, d> of the line search, solverCG.h:127), read together with S_GEN
    S_BARRIER2 = 29, // dummy operand of the slab barrier issued on the second stream (chunked pipeline)
    S_BARRIER = 30,  // dummy operand of the stream-ordered slab barrier
    S_STAGE = 31,    // host -> device staging slot
    S_ERRMAX = 32,   // 4 slots: MAX over the slabs of (S_L1, S_L2SQ, -, S_LINF): Solver::compute_error allreduces with MPI_MAX
                     //          for every measure (include/solver.h:430); equals the local values when world_size == 1
    S_GENMAX = 36,   // 4 slots: the same for the generic block S_GEN
    S_COUNT = 40
};
