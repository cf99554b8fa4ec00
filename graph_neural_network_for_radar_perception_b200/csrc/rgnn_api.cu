// Error reporting, version, device queries.
#include <stdarg.h>

#include "rgnn_common.cuh"

namespace rgnn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

}  // namespace rgnn

extern "C" int rgnn_version(void) { return 100; }
extern "C" const char* rgnn_last_error(void) { return rgnn::g_err; }
