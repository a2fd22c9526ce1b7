#define PDDP_KNOWN_T double
#include "known_impl.cuh"
