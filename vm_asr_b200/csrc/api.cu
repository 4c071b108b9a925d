// Error plumbing and small utilities of the C ABI (include/vmasr_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace vmasr {

static thread_local char g_err[512] = "";

int fail(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    return fail("%s: %s", what, cudaGetErrorString(e));
}

bool pdl_enabled() {
    static const bool on = [] { const char *e = tuning_env("VMASR_PDL"); return e ? atoi(e) != 0 : true; }();
    return on;
}

int sm_count(int device) {
    static int cached[64];
    static std::once_flag once;
    std::call_once(once, [] { for (int &c : cached) c = 0; });
    if (device < 0 || device >= 64) return 148;
    if (cached[device] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
        cached[device] = n;
    }
    return cached[device];
}

#ifdef VMASR_TUNING
static unsigned long long *g_timeline = nullptr;
unsigned long long *debug_timeline() { return g_timeline; }
#endif

}  // namespace vmasr

#ifdef VMASR_TUNING
// measurement builds only (not in include/vmasr_b200.h): device buffer of 16 x grid timestamps, or NULL to switch it off
extern "C" __attribute__((visibility("default"))) void vmasr_debug_timeline(unsigned long long *buf) { vmasr::g_timeline = buf; }
#endif

extern "C" int vmasr_abi_version(void) { return VMASR_ABI_VERSION; }
extern "C" const char *vmasr_last_error(void) { return vmasr::g_err; }
