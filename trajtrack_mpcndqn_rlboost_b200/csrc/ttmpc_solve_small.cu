// ttmpc_solve_small.cu -- second build of the solve kernel with every hot loop rolled
// (see the note at the top of ttmpc_solve.cu): solve_kernel_small / launch_solve_small.
#define TTMPC_SMALL_CODE 1
#include "ttmpc_solve.cu"
