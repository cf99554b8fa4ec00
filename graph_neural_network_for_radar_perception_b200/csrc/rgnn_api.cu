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
    static int cache[64] = {0};         // per device; a racing first call writes the same value twice
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int& n = cache[dev & 63];
    if (n == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        n = v;
    }
    return n;
}

}  // namespace rgnn

extern "C" int rgnn_version(void) { return 100; }
extern "C" const char* rgnn_last_error(void) { return rgnn::g_err; }
