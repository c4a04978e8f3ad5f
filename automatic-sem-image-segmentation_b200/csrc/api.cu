// Library-level entry points: version, error string, launch counter, device check.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

namespace semb {

static thread_local char g_err[512] = "";
long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("SEMB_NO_PDL"); return !(e && e[0] && e[0] != '0'); }();
    return on;
}

int check_launch(const char* what) {
    ++g_launch_count;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
        return SEMB_ECUDA;
    }
    return SEMB_OK;
}

}  // namespace semb

extern "C" int semb_version(void) { return SEMB_VERSION; }
extern "C" const char* semb_last_error(void) { return semb::g_err; }
extern "C" int64_t semb_launch_count(void) { return semb::g_launch_count; }

extern "C" int semb_device_ok(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        semb::set_error("no CUDA device");
        cudaGetLastError();
        return 0;
    }
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
        semb::set_error("cudaGetDeviceProperties failed");
        cudaGetLastError();
        return 0;
    }
    if (p.major != 10) {
        semb::set_error("libsemb200 is built for sm_100a only; device is sm_%d%d", p.major, p.minor);
        return 0;
    }
    return 1;
}
