// Error reporting, version and flat-parameter layout tables of the C ABI (include/simgan_b200.h).
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "sg_common.cuh"

namespace sg {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return SG_OK;
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return SG_ERR_CUDA;
}
}  // namespace sg

extern "C" {
#pragma GCC visibility push(default)

const char* sg_last_error(void) { return sg::g_err; }

int sg_version(void) { return 100; }

long long sg_launch_count(void) { return sg::launches(); }

int sg_device_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int sg_policy_layout(int obs_dim, int hidden, int act_dim, int* offsets) {
    if (obs_dim <= 0 || hidden <= 0 || act_dim <= 0) { sg::set_error("sg_policy_layout: non-positive dims"); return -1; }
    sg::PolicyLayout L = sg::make_policy_layout(obs_dim, hidden, act_dim);
    if (offsets) {
        const int o[SG_POLICY_SEGMENTS] = {L.aw1, L.ab1, L.aw2, L.ab2, L.cw1, L.cb1, L.cw2, L.cb2, L.vw, L.vb, L.mw, L.mb, L.ls};
        for (int i = 0; i < SG_POLICY_SEGMENTS; ++i) offsets[i] = o[i];
    }
    return L.total;
}

int sg_disc_layout(int feat_dim, int hidden, int* offsets) {
    if (feat_dim <= 0 || hidden <= 0) { sg::set_error("sg_disc_layout: non-positive dims"); return -1; }
    sg::DiscLayout L = sg::make_disc_layout(feat_dim, hidden);
    if (offsets) {
        const int o[SG_DISC_SEGMENTS] = {L.w1, L.b1, L.w2, L.b2, L.w3, L.b3};
        for (int i = 0; i < SG_DISC_SEGMENTS; ++i) offsets[i] = o[i];
    }
    return L.total;
}

#pragma GCC visibility pop
}  // extern "C"
