// "Column-owner" micro-kernels for the resident step kernels (weights in shared memory).
//
// A thread owns ONE output column of a small dense layer for RT of the tile's rows and walks the whole
// contraction itself: no cross-lane reduction, no shared-memory partial sums, no barrier inside a layer.
//   activations   row-major [row][ld]   (ld % 4 == 0, zero padded)   -> one broadcast LDS.128 per row / k-quad
//   weights       either the natural nn.Linear (out,in) image, or the "NQ" image of a square-ish matrix:
//                 element (n,k) at ((n>>2)*(K+1) + k)*4 + (n&3)   (K+1: one float4 of padding per n-quad row)
//                   forward  (thread = n, walks k):  scalar LDS, lanes n..n+31 hit 32 different banks
//                   backward (thread = k, walks n):  one LDS.128 per n-quad, lanes k..k+31 consecutive float4
// so ONE image serves both directions of the layer.
#pragma once
#include "sg_common.cuh"

namespace sg {

__host__ __device__ inline int nq_image_floats(int N, int K) { return ((N + 3) / 4) * (K + 1) * 4; }
__device__ __forceinline__ int nq_index(int n, int k, int K) { return ((n >> 2) * (K + 1) + k) * 4 + (n & 3); }

// acc[i] += sum_k X[(r0+i)*ldx + k] * Wrow[k]      (Wrow = natural row n of an (N,K) matrix)
template <int RT>
__device__ __forceinline__ void col_dot_nat(const float* __restrict__ Wrow, int K, const float* __restrict__ X, int ldx,
                                            int r0, float (&acc)[RT]) {
    for (int k0 = 0; k0 < K; k0 += 4) {
        float w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = (k0 + j < K) ? Wrow[k0 + j] : 0.f;
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            const float4 x = *reinterpret_cast<const float4*>(X + (r0 + i) * ldx + k0);
            acc[i] = fmaf(x.x, w[0], acc[i]); acc[i] = fmaf(x.y, w[1], acc[i]);
            acc[i] = fmaf(x.z, w[2], acc[i]); acc[i] = fmaf(x.w, w[3], acc[i]);
        }
    }
}

// acc[i] += sum_k X[(r0+i)*ldx + k] * W(n,k), W in the NQ image (K % 4 == 0)
template <int RT>
__device__ __forceinline__ void col_dot_nq_fwd(const float* __restrict__ Wq, int K, int n, const float* __restrict__ X,
                                               int ldx, int r0, float (&acc)[RT]) {
    const float* wp = Wq + (size_t)(n >> 2) * (K + 1) * 4 + (n & 3);
#pragma unroll 2
    for (int k0 = 0; k0 < K; k0 += 4) {
        const float w0 = wp[(k0) * 4], w1 = wp[(k0 + 1) * 4], w2 = wp[(k0 + 2) * 4], w3 = wp[(k0 + 3) * 4];
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            const float4 x = *reinterpret_cast<const float4*>(X + (r0 + i) * ldx + k0);
            acc[i] = fmaf(x.x, w0, acc[i]); acc[i] = fmaf(x.y, w1, acc[i]);
            acc[i] = fmaf(x.z, w2, acc[i]); acc[i] = fmaf(x.w, w3, acc[i]);
        }
    }
}

// acc[i] += sum_n DY[(r0+i)*ldy + n] * W(n,k), W in the NQ image of an (N,K) matrix (N % 4 == 0)
template <int RT>
__device__ __forceinline__ void col_dot_nq_bwd(const float* __restrict__ Wq, int N, int K, int k,
                                               const float* __restrict__ DY, int ldy, int r0, float (&acc)[RT]) {
    const float* wp = Wq + (size_t)k * 4;
    const int stride = (K + 1) * 4;
#pragma unroll 2
    for (int n0 = 0; n0 < N; n0 += 4) {
        const float4 w = *reinterpret_cast<const float4*>(wp + (size_t)(n0 >> 2) * stride);
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            const float4 y = *reinterpret_cast<const float4*>(DY + (r0 + i) * ldy + n0);
            acc[i] = fmaf(y.x, w.x, acc[i]); acc[i] = fmaf(y.y, w.y, acc[i]);
            acc[i] = fmaf(y.z, w.z, acc[i]); acc[i] = fmaf(y.w, w.w, acc[i]);
        }
    }
}

// G[n*K + k] (+)= sum_r Yt[n*R + r] * X[r*ldx + k]   for an (N,K) weight gradient, K % 4 == 0.
// Register tile: a work item is 4 rows n x QG k-quads (QG <= 4 chosen so that one sweep covers the matrix);
// the 4xR block of Yt is loaded once and every X quad is used for 4 outputs rows, i.e. ~3 FMAs per float that
// crosses the shared-memory port instead of 1.  float4 results go straight to L2 (st.cg).
template <int R>
__device__ __forceinline__ void outer_cols(float* __restrict__ G, const float* __restrict__ Yt, const float* __restrict__ X,
                                           int ldx, int N, int K, int t, int nth, bool acc) {
    static_assert(R % 4 == 0, "rows come in float4 groups");
    const int KQ = K >> 2;
    const int NB = (N + 3) >> 2;
    int qg = 1;
    while (qg < 4 && NB * ((KQ + qg - 1) / qg) > nth) ++qg;
    const int ngrp = (KQ + qg - 1) / qg;
    for (int idx = t; idx < NB * ngrp; idx += nth) {
        const int nb = idx / ngrp, g = idx - nb * ngrp;
        const int n0 = nb * 4, q0 = g * qg;
        float y[4][R];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (n0 + i < N) {
                load_rows_t<R>(Yt, n0 + i, y[i]);
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) y[i][r] = 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int q = q0 + j;
            if (j >= qg || q >= KQ) break;
            float4 o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 x = *reinterpret_cast<const float4*>(X + r * ldx + 4 * q);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    o[i].x = fmaf(y[i][r], x.x, o[i].x); o[i].y = fmaf(y[i][r], x.y, o[i].y);
                    o[i].z = fmaf(y[i][r], x.z, o[i].z); o[i].w = fmaf(y[i][r], x.w, o[i].w);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (n0 + i < N) {
                    float4* gp = reinterpret_cast<float4*>(G + (size_t)(n0 + i) * K + 4 * q);
                    float4 v = o[i];
                    if (acc) { const float4 old = __ldcg(gp); v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w; }
                    __stcg(gp, v);
                }
            }
        }
    }
}

// Same product, one row n x an interleaved set of k-quads per work item (lighter on registers; what the PPO
// column tile uses).  G[n*K + k] (+)= sum_r Yt[n*R + r] * X[r*ldx + k], K % 4 == 0.
// Work item = (row n, interleaved k-quad lane s of NSEG); float4 results go straight to L2 (st.cg).
template <int R, int PF = 0>
__device__ __forceinline__ void outer_cols_seg(float* __restrict__ G, const float* __restrict__ Yt, const float* __restrict__ X,
                                           int ldx, int N, int K, int t, int nth, bool acc) {
    const int KQ = K >> 2;
    int nseg = 1, lg = 0;
    while (N * nseg < nth && 2 * nseg <= KQ) { nseg *= 2; ++lg; }
    for (int idx = t; idx < N * nseg; idx += nth) {
        const int n = idx >> lg, s = idx & (nseg - 1);
        float y[R];
        load_rows_t<R>(Yt, n, y);
        if (PF) {       // groups of PF quads: old values requested before the group's FMAs when accumulating (see outer_store)
            constexpr int NB = PF > 0 ? PF : 1;
            for (int kq0 = s; kq0 < KQ; kq0 += NB * nseg) {
                float4 old[NB];
                if (acc) {
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        const int kq = kq0 + j * nseg;
                        if (kq < KQ) old[j] = __ldcg(reinterpret_cast<const float4*>(G + (size_t)n * K + 4 * kq));
                    }
                }
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    const int kq = kq0 + j * nseg;
                    if (kq >= KQ) break;
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const float4 x = *reinterpret_cast<const float4*>(X + r * ldx + 4 * kq);
                        o.x = fmaf(y[r], x.x, o.x); o.y = fmaf(y[r], x.y, o.y); o.z = fmaf(y[r], x.z, o.z); o.w = fmaf(y[r], x.w, o.w);
                    }
                    if (acc) { o.x += old[j].x; o.y += old[j].y; o.z += old[j].z; o.w += old[j].w; }
                    __stcg(reinterpret_cast<float4*>(G + (size_t)n * K + 4 * kq), o);
                }
            }
            continue;
        }
        for (int kq = s; kq < KQ; kq += nseg) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 x = *reinterpret_cast<const float4*>(X + r * ldx + 4 * kq);
                o.x = fmaf(y[r], x.x, o.x); o.y = fmaf(y[r], x.y, o.y); o.z = fmaf(y[r], x.z, o.z); o.w = fmaf(y[r], x.w, o.w);
            }
            float4* gp = reinterpret_cast<float4*>(G + (size_t)n * K + 4 * kq);
            if (acc) { const float4 old = __ldcg(gp); o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
            __stcg(gp, o);
        }
    }
}

}  // namespace sg
