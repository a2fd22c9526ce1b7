#include "profile.h"
#include "../../include/pddp_b200.h"
#include <vector>

namespace pddp {
static bool g_on = false;
static std::vector<cudaEvent_t> g_ev[PROF_KINDS];
static size_t g_used[PROF_KINDS] = {};
static long long g_launches = 0;

static cudaEvent_t next_event(int kind) {
    if (g_used[kind] == g_ev[kind].size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        g_ev[kind].push_back(e);
    }
    return g_ev[kind][g_used[kind]++];
}
void prof_begin(int kind, cudaStream_t st) { if (g_on) cudaEventRecord(next_event(kind), st); }
void prof_end(int kind, cudaStream_t st) { if (g_on) cudaEventRecord(next_event(kind), st); }
void note_launches(long long n) { g_launches += n; }
}  // namespace pddp

using namespace pddp;

extern "C" void pddp_profile_enable(int on) {
    g_on = on != 0;
    for (int k = 0; k < PROF_KINDS; ++k) g_used[k] = 0;
}

// Synchronises the device, then returns per-kind total milliseconds and launch counts.
extern "C" int pddp_profile_read(double* ms, int64_t* count) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return (int)e;
    for (int k = 0; k < PROF_KINDS; ++k) {
        double total = 0;
        for (size_t i = 0; i + 1 < g_used[k]; i += 2) {
            float t = 0;
            cudaEventElapsedTime(&t, g_ev[k][i], g_ev[k][i + 1]);
            total += t;
        }
        ms[k] = total;
        count[k] = (int64_t)(g_used[k] / 2);
    }
    return 0;
}

extern "C" int64_t pddp_launch_count(void) { return g_launches; }
