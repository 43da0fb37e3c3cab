// Peer-memory exchange buffers for the fused data-parallel gradient exchange (include/simgan_b200.h: sg_dp_*).
//
// Every rank owns ONE exchange buffer in its own HBM; peers write into it over NVLink (P2P stores through
// pointers obtained with cudaIpcOpenMemHandle) and the owner only ever reads it locally:
//     gather[2][world][2*cap]  u32   slot (b, r): rank r's locally reduced gradient for a step of parity b,
//                                    stored as (value, step id) pairs -- every 8-byte pair is written atomically
// The persistent step kernels push slice c of their locally reduced gradient to every peer inside phase B and
// poll their own buffer until the peers' pairs carry the current step id ("low-latency" protocol: the data IS
// the flag, so there is no fence and no separate flag round trip -- one NVLink one-way latency per step), then add
// the world's slices in RANK ORDER: every rank forms bit-identical totals, with no collective library call and
// no extra kernel launch, and a CTA only waits for the slice it needs.
#pragma once
#include "sg_common.cuh"

namespace sg {

constexpr int kDpMaxWorld = 16;
constexpr int kDpMaxSlices = 160;      // >= #SMs (one slice per CTA) + 1 for the loss sums

struct DpView {                        // passed by value inside the kernel argument structs
    int rank, world;
    int cap;                           // floats per gather slot
    float* gather[kDpMaxWorld];        // base of rank r's exchange buffer as mapped in THIS process
    unsigned int* flags[kDpMaxWorld];
    unsigned int* error_flag;          // local: set when a wait times out
};

__host__ __device__ inline size_t dp_gather_floats(int world, int cap) { return (size_t)2 * world * 2 * cap; }   // (value,id) pairs
__host__ __device__ inline size_t dp_bytes(int world, int cap) {
    return dp_gather_floats(world, cap) * sizeof(float) + (size_t)2 * world * kDpMaxSlices * sizeof(unsigned int) + 256;
}

__device__ __forceinline__ uint4* dp_slot(const DpView& d, int owner, int parity, int writer) {
    // 2*cap u32 per slot = cap/2 ... (value,id) pairs: float4 j of the gradient occupies uint4 [2j, 2j+1]
    return reinterpret_cast<uint4*>(d.gather[owner] + ((size_t)parity * d.world + writer) * 2 * d.cap);
}

// Exchange slice [p0,p1) of the locally reduced gradient `grad` (global, written with st.cg by this CTA) with all
// peers and leave the rank-ordered total in `grad` (and in scr (float*), when non-null, for narrow slices).
// `stepid` strictly increases across steps and calls (and is never 0).  All NT threads of the CTA must call.
// Ends with a CTA barrier.
constexpr long long kDpWaitNs = 60ll * 1000 * 1000 * 1000;      // 60 s

template <int NT>
__device__ __forceinline__ void dp_exchange_slice(const DpView& d, float* __restrict__ grad, int p0, int p1, int slice,
                                                  unsigned int stepid, float* scr) {
    (void)slice;
    const int tid = threadIdx.x;
    const int parity = stepid & 1u;
    const int n4 = (p1 - p0) >> 2;
    __syncthreads();                                   // the local slice (st.cg by other threads) is complete
    // 1. push my slice, tagged with the step id, into every peer's slot for me
    for (int j = tid; j < n4; j += NT) {
        const float4 v = ld_cg4(grad + p0 + 4 * j);
        const uint4 lo = make_uint4(__float_as_uint(v.x), stepid, __float_as_uint(v.y), stepid);
        const uint4 hi = make_uint4(__float_as_uint(v.z), stepid, __float_as_uint(v.w), stepid);
        for (int r = 0; r < d.world; ++r) {
            if (r == d.rank) continue;
            uint4* dst = dp_slot(d, r, parity, d.rank) + (size_t)(p0 >> 1) + 2 * j;
            asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
            asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst + 1), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
        }
    }
    // 2. poll my own buffer until every peer's pairs carry this step id; 3. rank-ordered total
    for (int j = tid; j < n4; j += NT) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < d.world; ++r) {
            float4 v;
            if (r == d.rank) {
                v = ld_cg4(grad + p0 + 4 * j);
            } else {
                const uint4* src = dp_slot(d, d.rank, parity, r) + (size_t)(p0 >> 1) + 2 * j;
                uint4 lo, hi;
                // a peer may start its launch late (its host draws / uploads at its own pace): wait by the wall clock, not by a
                // spin count, and only give up (error flag -> the trace is poisoned, the host raises) after kDpWaitNs
                long long spins = 0, t_start = 0;
                while (true) {
                    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "l"(src) : "memory");
                    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "l"(src + 1) : "memory");
                    if (lo.y == stepid && lo.w == stepid && hi.y == stepid && hi.w == stepid) break;
                    if ((++spins & 1023) == 0) {
                        if (*(volatile unsigned int*)d.error_flag != 0u) break;
                        long long now;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                        if (t_start == 0) t_start = now;
                        else if (now - t_start > kDpWaitNs) { atomicExch(d.error_flag, 1u); break; }
                    }
                }
                v = make_float4(__uint_as_float(lo.x), __uint_as_float(lo.z), __uint_as_float(hi.x), __uint_as_float(hi.z));
            }
            t = f4_add(t, v);
        }
        __stcg(reinterpret_cast<float4*>(grad + p0 + 4 * j), t);
        if (scr) *reinterpret_cast<float4*>(scr + 4 * j) = t;
    }
    __syncthreads();
}

}  // namespace sg
