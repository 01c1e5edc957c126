// Shared declarations of libpmp_b200 (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <map>

#include "../../include/pmp_b200.h"

namespace pmp {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define PMP_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (call);                                                         \
        if (_e != cudaSuccess) return pmp::cuda_fail(_e, #call, __FILE__, __LINE__);     \
    } while (0)

#define PMP_CHECK_ARG(cond, msg)                                   \
    do {                                                           \
        if (!(cond)) { pmp::set_error("bad argument: %s", msg); return PMP_ERR_ARG; } \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- profiling classes (CUDA-event timers per kernel family; off by default) -------------
enum ProfClass { PROF_CONV_TC = 0, PROF_CONV_SIMT, PROF_ELEMWISE, PROF_DECODE, PROF_POSTPROC, PROF_ASSEMBLE,
                 PROF_PREP, PROF_TEXT, PROF_COUNT };

struct ProfSlot {
    double ms = 0, flops = 0, bytes = 0;
    long long launches = 0;
};

struct Handle;

// RAII-less helper: brackets one launch with events when profiling is on.
struct ProfScope {
    Handle *h; int cls; cudaStream_t s; double flops, bytes; cudaEvent_t e0 = nullptr, e1 = nullptr;
    ProfScope(Handle *h, int cls, cudaStream_t s, double flops, double bytes);
    ~ProfScope();
};

}  // namespace pmp
