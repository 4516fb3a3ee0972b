// Library-wide plumbing: version, thread-local error string, cached device attributes.
#include "common.cuh"
#include <string.h>

namespace sgb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int num_sms() {
    static thread_local int cached_dev = -1;
    static thread_local int cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached_dev = dev;
        cached_sms = v;
    }
    return cached_sms;
}

}  // namespace sgb

extern "C" int sgb_version(void) { return SGB_VERSION; }
extern "C" const char* sgb_last_error(void) { return sgb::g_err; }
extern "C" int sgb_num_sms(void) { return sgb::num_sms(); }
