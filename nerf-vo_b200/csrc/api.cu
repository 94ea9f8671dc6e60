// Library-level entry points: error string, version, device info.
#include <stdarg.h>
#include <stdlib.h>
#include "nvo_common.cuh"

static thread_local char g_err[512] = "";

void nvo_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int nvo_sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

int nvo_env_int(const char* name, int fallback) {
    const char* e = getenv(name);
    return e ? atoi(e) : fallback;
}

static long long g_launches = 0;
void nvo_count_launch() { ++g_launches; }

extern "C" const char* nvo_last_error(void) { return g_err; }
extern "C" int64_t nvo_launch_count(void) { return g_launches; }
extern "C" int nvo_version(void) { return 100; }
extern "C" int nvo_batch_size_granularity(void) { return 128; }
