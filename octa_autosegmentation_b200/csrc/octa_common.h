// Shared host-side plumbing of the C-ABI library: error reporting, launch counter, CUDA checks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/octa_b200.h"

namespace octa {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
// while a thread captures a CUDA graph its launches are recorded, not executed: they are counted per graph launch instead
extern thread_local bool t_capturing;
extern thread_local uint64_t t_captured;
inline void count_launch(int n = 1) {
    if (t_capturing) t_captured += (uint64_t)n;
    else g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed);
}

#define OCTA_CUDA_CHECK(expr)                                                                 \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            octa::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return OCTA_E_CUDA;                                                               \
        }                                                                                     \
    } while (0)

#define OCTA_ARG_CHECK(cond, msg)                              \
    do {                                                       \
        if (!(cond)) {                                         \
            octa::set_error("%s: %s", __func__, msg);          \
            return OCTA_E_ARG;                                 \
        }                                                      \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace octa
