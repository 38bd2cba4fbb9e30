// Host-side plumbing of the C ABI: error string, version, device query.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace rnad {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return RNAD_OK;
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return RNAD_ECUDA;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (dev != cached_dev) {
        if (cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace rnad

extern "C" {

const char* rnad_last_error(void) { return rnad::g_error; }

int rnad_version(void) { return 200; }

int rnad_device_sm_count(void) {
    int n = rnad::sm_count();
    if (n < 0) {
        rnad::set_error("no CUDA device");
        return RNAD_ECUDA;
    }
    return n;
}

int rnad_packed_strides(int A, int C, int* ev_stride, int* tr_stride) {
    RNAD_REQUIRE(A >= 1 && A <= RNAD_MAX_ACTIONS, "max_actions %d outside [1,%d]", A, RNAD_MAX_ACTIONS);
    RNAD_REQUIRE(C >= 1 && C <= RNAD_MAX_TRANSITIONS, "max_transitions %d outside [1,%d]", C, RNAD_MAX_TRANSITIONS);
    if (ev_stride) *ev_stride = rnad::ev_stride_of(A);
    if (tr_stride) *tr_stride = rnad::tr_stride_of(C);
    return RNAD_OK;
}

}  // extern "C"
