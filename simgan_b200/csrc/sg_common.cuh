// Shared device primitives for the simgan_b200 kernels (sm_100a).
//
// The per-minibatch math of the reference is a chain of small dense contractions over a tile of R
// minibatch rows owned by one CTA (R = 8): X(R,K)·W(N,K)^T, dY(R,N)·W(N,K) and dY^T·X.  At the
// reference's sizes (hidden 64..256, 1024 rows per minibatch spread over 128 CTAs) these are
// latency-bound register-tile FMA problems, not tensor-core tiles; see DESIGN.md "Kernels".
//
// Layout conventions inside a CTA tile:
//   forward activations   row-major   A[r*ld + k]      (ld multiple of 4, zero padded)
//   backward deltas       transposed  D[n*R + r]       (all R rows of one unit contiguous)
// Weights are nn.Linear (out,in) row-major in global memory and are read with ld.global.cg: they are
// rewritten by the Adam phase between optimizer steps of the same persistent kernel, so the
// non-coherent (.nc / L1) path must not be used.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/simgan_b200.h"

namespace sg {

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
void count_launches(int n);   // kernel launches issued by this library (sg_launch_count)

#define SG_CUDA(call)                                      \
    do {                                                   \
        int _rc = ::sg::check_cuda((call), #call);         \
        if (_rc) return _rc;                               \
    } while (0)

#define SG_REQUIRE(cond, ...)                              \
    do {                                                   \
        if (!(cond)) {                                     \
            ::sg::set_error(__VA_ARGS__);                  \
            return SG_ERR_INVALID;                         \
        }                                                  \
    } while (0)

// Largest dynamic shared-memory size already granted to one kernel, per device (function attributes are per device;
// a process normally drives one GPU, but nothing here may assume it).
struct SmemGrant {
    size_t granted[64] = {};
};
template <class K>
inline int grant_smem(SmemGrant& g, K kernel, size_t smem) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (smem > 48 * 1024 && smem > g.granted[dev]) {
        SG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        g.granted[dev] = smem;
    }
    return SG_OK;
}

constexpr int kStepThreads = 256;   // threads per CTA of the step kernels
constexpr int kHalf = 128;          // actor / critic halves
constexpr int kRows = 8;            // minibatch rows per CTA tile

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// Where a kernel reads the network weights from.  LdGlobal: flat parameter vector in global memory, read
// through L2 (ld.global.cg) because the Adam phase of the same persistent kernel rewrites it between
// steps.  LdShared: a CTA-private image of the flat vector in shared memory ("resident" kernels).
struct LdGlobal {
    static __device__ __forceinline__ float ld(const float* p) { return __ldcg(p); }
    static __device__ __forceinline__ float4 ld4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
};
struct LdShared {
    static __device__ __forceinline__ float ld(const float* p) { return *p; }
    static __device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// OUT(r,n) = sum_k X[r*ldx+k] * W[n*K+k]          (epilogue adds bias / activation and stores)
//   `nth` threads (local id t, nth % 32 == 0) cooperate; EVERY thread of the participating warps
//   must call (warp shuffles).  8 lanes split K (stride 8*V), each lane carries 4 outputs x R rows;
//   a butterfly reduce-scatter leaves every lane with one output and R/2 rows.
//   V = 4 needs K % 4 == 0 and 16-byte aligned W rows; V = 1 handles any K.
// ------------------------------------------------------------------------------------------------
template <int R, int V, class WL = LdGlobal, class Epi>
__device__ __forceinline__ void gemm_xwT(const float* __restrict__ W, const float* __restrict__ X, int ldx,
                                         int N, int K, int t, int nth, Epi epi) {
    static_assert(R % 2 == 0, "R must be even");
    constexpr int KS = 8, NT = 4;
    const int kq = t & (KS - 1);
    const int grp = t >> 3;
    const int ngrp = nth >> 3;
    for (int base = 0; base < N; base += ngrp * NT) {
        const int n0 = base + grp * NT;
        float acc[NT][R];
#pragma unroll
        for (int i = 0; i < NT; ++i)
#pragma unroll
            for (int r = 0; r < R; ++r) acc[i][r] = 0.f;
        if (n0 < N) {
            for (int k = kq * V; k < K; k += KS * V) {
                float w[NT][V];
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    if (n0 + i < N) {
                        if (V == 4) {
                            float4 q = WL::ld4(W + (size_t)(n0 + i) * K + k);
                            w[i][0] = q.x; w[i][1 % V] = q.y; w[i][2 % V] = q.z; w[i][3 % V] = q.w;
                        } else {
                            w[i][0] = WL::ld(W + (size_t)(n0 + i) * K + k);
                        }
                    } else {
#pragma unroll
                        for (int v = 0; v < V; ++v) w[i][v] = 0.f;
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float x[V];
                    if (V == 4) {
                        float4 q = *reinterpret_cast<const float4*>(X + r * ldx + k);
                        x[0] = q.x; x[1 % V] = q.y; x[2 % V] = q.z; x[3 % V] = q.w;
                    } else {
                        x[0] = X[r * ldx + k];
                    }
#pragma unroll
                    for (int i = 0; i < NT; ++i)
#pragma unroll
                        for (int v = 0; v < V; ++v) acc[i][r] = fmaf(w[i][v], x[v], acc[i][r]);
                }
            }
        }
        // reduce-scatter over the 8 K-split lanes
        const bool up4 = (kq & 4) != 0, up2 = (kq & 2) != 0, up1 = (kq & 1) != 0;
        float a2[2][R];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float send = up4 ? acc[j][r] : acc[j + 2][r];
                float recv = __shfl_xor_sync(0xffffffffu, send, 4);
                a2[j][r] = (up4 ? acc[j + 2][r] : acc[j][r]) + recv;
            }
        float a1[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float send = up2 ? a2[0][r] : a2[1][r];
            float recv = __shfl_xor_sync(0xffffffffu, send, 2);
            a1[r] = (up2 ? a2[1][r] : a2[0][r]) + recv;
        }
        float a0[R / 2];
#pragma unroll
        for (int r = 0; r < R / 2; ++r) {
            float send = up1 ? a1[r] : a1[r + R / 2];
            float recv = __shfl_xor_sync(0xffffffffu, send, 1);
            a0[r] = (up1 ? a1[r + R / 2] : a1[r]) + recv;
        }
        const int n = n0 + (up4 ? 2 : 0) + (up2 ? 1 : 0);
        if (n < N) {
#pragma unroll
            for (int r = 0; r < R / 2; ++r) epi(r + (up1 ? R / 2 : 0), n, a0[r]);
        }
    }
}

template <int R>
__device__ __forceinline__ void load_rows_t(const float* __restrict__ Yt, int n, float (&y)[R]) {
    if (R % 4 == 0) {
#pragma unroll
        for (int q = 0; q < R / 4; ++q) {
            float4 v = *reinterpret_cast<const float4*>(Yt + (size_t)n * R + 4 * q);
            y[4 * q] = v.x; y[4 * q + 1] = v.y; y[4 * q + 2] = v.z; y[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) y[r] = Yt[(size_t)n * R + r];
    }
}

// ------------------------------------------------------------------------------------------------
// OUT(r,k) = sum_n Yt[n*R+r] * W[n*K+k]           (back-propagation through a Linear layer)
//   Threads are laid out (k-chunk, n-slice); partial sums go through `scratch`
//   (>= nth*R*V floats of shared memory) and are reduced after a CTA-wide barrier.
//   ALL threads of the CTA must call (contains __syncthreads); `sync_after` adds the trailing one.
// ------------------------------------------------------------------------------------------------
template <int R, int V, class WL = LdGlobal, class Epi>
__device__ __forceinline__ void gemm_yW(const float* __restrict__ W, const float* __restrict__ Yt, int N, int K,
                                        float* __restrict__ scratch, int t, int nth, Epi epi) {
    const int KC = (K + V - 1) / V;              // V==4 requires K%4==0
    const int KCp = KC < nth ? KC : nth;         // chunks handled per sweep
    const int NS = KC < nth ? nth / KC : 1;      // n-slices
    const int Nper = (N + NS - 1) / NS;
    for (int kc0 = 0; kc0 < KC; kc0 += KCp) {
        const int kc = kc0 + (t % KCp);
        const int ns = t / KCp;
        float acc[R][V];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int v = 0; v < V; ++v) acc[r][v] = 0.f;
        const bool live = (ns < NS) && (kc < KC);
        if (live) {
            const int n_end = min(N, (ns + 1) * Nper);
            for (int n = ns * Nper; n < n_end; ++n) {
                float w[V];
                if (V == 4) {
                    float4 q = WL::ld4(W + (size_t)n * K + kc * 4);
                    w[0] = q.x; w[1 % V] = q.y; w[2 % V] = q.z; w[3 % V] = q.w;
                } else {
                    w[0] = WL::ld(W + (size_t)n * K + kc);
                }
                float y[R];
                load_rows_t<R>(Yt, n, y);
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int v = 0; v < V; ++v) acc[r][v] = fmaf(y[r], w[v], acc[r][v]);
            }
            // scratch[(ns*R + r) * (KCp*V) + (kc-kc0)*V + v]
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int v = 0; v < V; ++v) scratch[(size_t)(ns * R + r) * (KCp * V) + (kc - kc0) * V + v] = acc[r][v];
        }
        __syncthreads();
        const int kw = min(KCp * V, K - kc0 * V);  // valid columns of this sweep
        if (kw <= nth) {
            // (row, column) mapping fixed per thread: one division per sweep instead of one per element
            const int rstep = nth / kw;
            const int kk = t % kw;
            for (int r = t / kw; r < R; r += rstep) {
                float s = 0.f;
                for (int q = 0; q < NS; ++q) s += scratch[(size_t)(q * R + r) * (KCp * V) + kk];
                epi(r, kc0 * V + kk, s);
            }
        } else {
            for (int r = 0; r < R; ++r)
                for (int kk = t; kk < kw; kk += nth) {
                    float s = 0.f;
                    for (int q = 0; q < NS; ++q) s += scratch[(size_t)(q * R + r) * (KCp * V) + kk];
                    epi(r, kc0 * V + kk, s);
                }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// G[n*K+k] (+)= sum_r Yt[n*R+r] * X[r*ldx+k]       (per-CTA partial weight gradient -> global)
// ------------------------------------------------------------------------------------------------
// PF > 0 (kernels whose CTAs run several tiles per step and therefore accumulate): rows are processed in groups of PF and
// the PF old values are requested from L2 BEFORE the group's FMAs, so that their round trips overlap each other and
// the arithmetic instead of one exposed L2 latency per row (which was half of the tile time at cfg 5).
template <int R, int V, int PF = 0>
__device__ __forceinline__ void outer_store(float* __restrict__ G, const float* __restrict__ Yt,
                                            const float* __restrict__ X, int ldx, int N, int K, int t, int nth,
                                            bool acc) {
    const int KC = (K + V - 1) / V;
    const int KCp = KC < nth ? KC : nth;
    const int NS = KC < nth ? nth / KC : 1;
    const int Nper = (N + NS - 1) / NS;
    for (int kc0 = 0; kc0 < KC; kc0 += KCp) {
        const int kc = kc0 + (t % KCp);
        const int ns = t / KCp;
        if (ns >= NS || kc >= KC) continue;
        float xr[R][V];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (V == 4) {
                float4 q = *reinterpret_cast<const float4*>(X + r * ldx + kc * 4);
                xr[r][0] = q.x; xr[r][1 % V] = q.y; xr[r][2 % V] = q.z; xr[r][3 % V] = q.w;
            } else {
                xr[r][0] = X[r * ldx + kc];
            }
        }
        const int n_end = min(N, (ns + 1) * Nper);
        if (PF) {
            constexpr int NB = PF > 0 ? PF : 1;
            for (int n0 = ns * Nper; n0 < n_end; n0 += NB) {
                float old[NB][V];
                if (acc) {
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        if (n0 + j < n_end) {
                            if (V == 4) {
                                const float4 q = __ldcg(reinterpret_cast<const float4*>(G + (size_t)(n0 + j) * K + kc * 4));
                                old[j][0] = q.x; old[j][1 % V] = q.y; old[j][2 % V] = q.z; old[j][3 % V] = q.w;
                            } else {
                                old[j][0] = __ldcg(G + (size_t)(n0 + j) * K + kc);
                            }
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    const int n = n0 + j;
                    if (n >= n_end) break;
                    float y[R];
                    load_rows_t<R>(Yt, n, y);
                    float o[V];
#pragma unroll
                    for (int v = 0; v < V; ++v) o[v] = 0.f;
#pragma unroll
                    for (int r = 0; r < R; ++r)
#pragma unroll
                        for (int v = 0; v < V; ++v) o[v] = fmaf(y[r], xr[r][v], o[v]);
                    if (acc) {
#pragma unroll
                        for (int v = 0; v < V; ++v) o[v] += old[j][v];
                    }
                    if (V == 4) __stcg(reinterpret_cast<float4*>(G + (size_t)n * K + kc * 4), make_float4(o[0], o[1 % V], o[2 % V], o[3 % V]));
                    else __stcg(G + (size_t)n * K + kc, o[0]);
                }
            }
            continue;
        }
        for (int n = ns * Nper; n < n_end; ++n) {
            float y[R];
            load_rows_t<R>(Yt, n, y);
            float o[V];
#pragma unroll
            for (int v = 0; v < V; ++v) o[v] = 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int v = 0; v < V; ++v) o[v] = fmaf(y[r], xr[r][v], o[v]);
            if (V == 4) {
                float4* gp = reinterpret_cast<float4*>(G + (size_t)n * K + kc * 4);
                float4 q = make_float4(o[0], o[1 % V], o[2 % V], o[3 % V]);
                if (acc) { const float4 old = __ldcg(gp); q.x += old.x; q.y += old.y; q.z += old.z; q.w += old.w; }
                __stcg(gp, q);
            } else {
                float* gp = G + (size_t)n * K + kc;
                __stcg(gp, acc ? o[0] + __ldcg(gp) : o[0]);
            }
        }
    }
}

// g[n] (+)= sum_{r<rows} Yt[n*R+r]   (bias gradient)
template <int R>
__device__ __forceinline__ void rowsum_store(float* __restrict__ g, const float* __restrict__ Yt, int N, int t, int nth,
                                             bool acc, int rows = R) {
    for (int n = t; n < N; n += nth) {
        float y[R];
        load_rows_t<R>(Yt, n, y);
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) s += (r < rows) ? y[r] : 0.f;
        __stcg(g + n, acc ? s + __ldcg(g + n) : s);
    }
}

// ------------------------------------------------------------------------------------------------
// Grid-wide barrier for cooperative (co-resident) launches: monotonically increasing arrival counter.
// `*counter` must be zero at kernel start.  `gen` is the per-thread-0 generation count.
// Arrive = red.release.gpu after the CTA barrier (cumulative: publishes every thread's prior writes),
// wait = ld.acquire.gpu spin by thread 0 followed by the CTA barrier.  Data exchanged across the
// barrier is written with st.cg and read with ld.cg (L2), so no L1 invalidation is needed.
// A spin cap turns a lost barrier into an error flag instead of a hung GPU.
// ------------------------------------------------------------------------------------------------
constexpr long long kGridBarrierWaitNs = 90ll * 1000 * 1000 * 1000;      // 90 s (> the peer-exchange cap of sg_dp.cuh)
struct GridBarrier {
    unsigned int* counter;
    unsigned int* error_flag;
    unsigned int nblocks;
    unsigned int gen;
    __device__ __forceinline__ void sync() {
        __syncthreads();
        if (threadIdx.x == 0) {
            gen += 1;
            const unsigned int target = gen * nblocks;
            asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
            unsigned int seen;
            long long spins = 0, t_start = 0;
            // once any barrier has timed out every later one falls through: the kernel drains quickly and
            // the host sees the poisoned trace.  The cap is wall-clock (kGridBarrierWaitNs): with data parallelism a CTA
            // may legitimately sit in the peer exchange for as long as the peer's host takes to launch its kernel
            bool dead = false;
            while (true) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
                if (seen >= target) break;
                if ((++spins & 1023) == 0) {
                    if (*(volatile unsigned int*)error_flag != 0u) dead = true;
                    long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    if (t_start == 0) t_start = now;
                    else if (now - t_start > kGridBarrierWaitNs) { atomicExch(error_flag, 1u); dead = true; }
                    if (dead) break;
                }
            }
        }
        __syncthreads();
    }
};

// ------------------------------------------------------------------------------------------------
// TMA bulk copies (cp.async.bulk, 1-D): global -> shared without register staging, completion signalled on an
// mbarrier by transaction bytes.  Used to refresh the CTA-resident weight images after every Adam step.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned int bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
    unsigned int done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// order this thread's earlier generic-proxy accesses (incl. what an acquire made visible) before later async-proxy ones
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Per-phase clock64 totals of one thread of one CTA of a persistent kernel (diagnostics: which part of an
// optimizer step the time goes to).  slot[i] accumulates the cycles between lap(i) and the previous lap.
struct PhaseClock {
    long long* slots;
    bool on;
    long long t0;
    __device__ __forceinline__ void start() { if (on) t0 = clock64(); }
    __device__ __forceinline__ void lap(int i) {
        if (on) { const long long t = clock64(); slots[i] += t - t0; t0 = t; }
    }
};

// Deterministic CTA-wide sum of one double per thread (256 threads): xor-butterfly inside each warp,
// then the 8 warp sums added in warp order by every thread.  `red` = 8 doubles of shared memory.
template <int NT = 256>
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();                       // protects `red` against the previous use
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) tot += red[w];
    return tot;
}
__device__ __forceinline__ double block_sum_256(double v, double* red) { return block_sum<256>(v, red); }

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// Sum of `nslots` partial vectors over the float range [p0,p1) (both multiples of 4):
// out[p] = sum_c part[c*stride + p] in a fixed order.  All 256 threads of the CTA must call.
// `scr4` = 256 float4 of shared memory.  Ends with a CTA barrier (the slice of `out` is visible to the
// whole CTA through L2 afterwards).  Narrow slices (<= 128 float4): thread tid < n4 finalises float4 tid of
// the slice and gets it back in `mine` (returns true), so callers can fuse work on the reduced values.
// KEEP: `scr4` has NT + 128 float4 and the reduced narrow slice is also left in scr4[NT .. NT+n4).
template <int NT = 256, bool KEEP = false>
__device__ __forceinline__ bool reduce_partials_slice(const float* __restrict__ part, size_t stride, int nslots, int p0,
                                                      int p1, float* __restrict__ out, float4* scr4, int tid,
                                                      float4& mine) {
    const int n4 = (p1 - p0) >> 2;
    mine = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n4 <= 0) { __syncthreads(); return true; }
    if (n4 <= 128) {
        // threads per float4 column; capped so that the serial combine below stays short for a very short slice
        // (the last slice of the vector: 3 float4 -> 85 partial sums per column made that CTA the straggler of the phase)
        const int NQ = NT / n4 < 16 ? NT / n4 : 16;
        const int j = tid % n4, q = tid / n4;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < NQ) {
            const float* src = part + p0 + 4 * j;
            // one fully predicated batch of 16 independent L2 loads per sweep (slots beyond nslots read as zero):
            // a sequential tail would cost one L2 round trip per leftover slot
            for (int c = q; c < nslots; c += 16 * NQ) {
                float4 v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int cc = c + u * NQ;
                    v[u] = cc < nslots ? ld_cg4(src + (size_t)cc * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 16; ++u) s = f4_add(s, v[u]);
            }
            scr4[q * n4 + j] = s;
        }
        __syncthreads();
        if (tid < n4) {
            float4 t = scr4[tid];
            for (int qq = 1; qq < NQ; ++qq) t = f4_add(t, scr4[qq * n4 + tid]);
            __stcg(reinterpret_cast<float4*>(out + p0 + 4 * tid), t);
            if (KEEP) scr4[NT + tid] = t;
            mine = t;
        }
        __syncthreads();
        return true;
    } else {
        // wide slices (large networks): every thread owns whole columns
        for (int j = tid; j < n4; j += NT) {
            const float* src = part + p0 + 4 * j;
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c = 0; c < nslots; c += 8) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    v[u] = c + u < nslots ? ld_cg4(src + (size_t)(c + u) * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 8; ++u) s = f4_add(s, v[u]);
            }
            __stcg(reinterpret_cast<float4*>(out + p0 + 4 * j), s);
        }
        __syncthreads();
        return false;
    }
}

// CTA-private shared-memory image of a flat parameter vector <- global (through L2); ends with a CTA barrier
__device__ __forceinline__ void load_param_image(float* Ws, const float* __restrict__ params, int P, int tid) {
    constexpr int U = 8;
    for (int p = 4 * tid; p < P; p += 4 * kStepThreads * U) {
        float4 q[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = p + 4 * kStepThreads * u;
            q[u] = i < P ? ld_cg4(params + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = p + 4 * kStepThreads * u;
            if (i < P) *reinterpret_cast<float4*>(Ws + i) = q[u];
        }
    }
    __syncthreads();
}

// flat policy layout (segment starts in floats), mirrored by sg_policy_layout()
struct PolicyLayout {
    int aw1, ab1, aw2, ab2, cw1, cb1, cw2, cb2, vw, vb, mw, mb, ls, total;
};
__host__ __device__ inline PolicyLayout make_policy_layout(int O, int H, int A) {
    PolicyLayout L;
    int o = 0;
    L.aw1 = o; o += round_up(H * O, 4);
    L.ab1 = o; o += round_up(H, 4);
    L.aw2 = o; o += round_up(H * H, 4);
    L.ab2 = o; o += round_up(H, 4);
    L.cw1 = o; o += round_up(H * O, 4);
    L.cb1 = o; o += round_up(H, 4);
    L.cw2 = o; o += round_up(H * H, 4);
    L.cb2 = o; o += round_up(H, 4);
    L.vw = o; o += round_up(H, 4);
    L.vb = o; o += 4;
    L.mw = o; o += round_up(A * H, 4);
    L.mb = o; o += round_up(A, 4);
    L.ls = o; o += round_up(A, 4);
    L.total = o;
    return L;
}
struct DiscLayout {
    int w1, b1, w2, b2, w3, b3, total;
};
__host__ __device__ inline DiscLayout make_disc_layout(int F, int H) {
    DiscLayout L;
    int o = 0;
    L.w1 = o; o += round_up(H * F, 4);
    L.b1 = o; o += round_up(H, 4);
    L.w2 = o; o += round_up(H * H, 4);
    L.b2 = o; o += round_up(H, 4);
    L.w3 = o; o += round_up(H, 4);
    L.b3 = o; o += 4;
    L.total = o;
    return L;
}

// Adam, in the op order of torch.optim.Adam's single-tensor path (torch/optim/adam.py,
// called from A2C/algo/ppo.py:145 and A2C/algo/gail.py:188).
__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, float one_minus_b1, float b2,
                                            float one_minus_b2, float step_size, float bc2_sqrt, float eps) {
    m = m + one_minus_b1 * (g - m);                         // exp_avg.lerp_(grad, 1-beta1)
    v = __fmul_rn(v, b2);                                   // exp_avg_sq.mul_(beta2)
    v = v + __fmul_rn(__fmul_rn(one_minus_b2, g), g);       //   .addcmul_(grad, grad, value=1-beta2)
    const float denom = __fdiv_rn(__fsqrt_rn(v), bc2_sqrt) + eps;
    p = p + __fdiv_rn(__fmul_rn(-step_size, m), denom);     // param.addcdiv_(exp_avg, denom, value=-step_size)
}

}  // namespace sg
