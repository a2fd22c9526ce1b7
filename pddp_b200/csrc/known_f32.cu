#define PDDP_KNOWN_T float
#include "known_impl.cuh"
