// Optional per-kernel timing (CUDA events on the launch stream) and a launch counter, used by
// bench.py to report the dominant kernel's average duration and `gpu_launches`.
#pragma once
#include <cuda_runtime.h>

namespace pddp {
enum { PROF_MLP_LIN = 0, PROF_MLP_ROLL = 1, PROF_MOMENT_LIN = 2, PROF_ROLL_STEP = 3, PROF_BACKWARD = 4, PROF_COST = 5,
       PROF_LIN_KNOWN = 6, PROF_ROLL_KNOWN = 7, PROF_ACCEPT = 8, PROF_KINDS = 9 };
void prof_begin(int kind, cudaStream_t st);
void prof_end(int kind, cudaStream_t st);
void note_launches(long long n);
}  // namespace pddp
