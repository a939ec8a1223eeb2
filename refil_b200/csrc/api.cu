// Error reporting and device queries for the C-ABI (include/refil_b200.h).
#include <stdarg.h>
#include <stdlib.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void refil_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* refil_last_error() { return g_err; }

extern "C" int refil_abi_version() { return 1; }

int refil_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

extern "C" int refil_device_sm_count() { return refil_num_sms(); }

bool refil_pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("REFIL_PDL");
        on = (e && e[0] == '1') ? 1 : 0;      // measured (profiles/r2_tuning.md): off by default
    }
    return on == 1;
}
