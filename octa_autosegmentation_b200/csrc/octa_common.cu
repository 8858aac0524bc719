// Error reporting / bookkeeping entry points of the C ABI (include/octa_b200.h).
#include "octa_common.h"
#include <stdarg.h>

namespace octa {
static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
thread_local bool t_capturing = false;
thread_local uint64_t t_captured = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace octa

extern "C" int octa_abi_version(void) { return OCTA_ABI_VERSION; }
extern "C" const char* octa_last_error(void) { return octa::g_err; }
extern "C" uint64_t octa_launch_count(void) { return octa::g_launches.load(); }
extern "C" int octa_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
