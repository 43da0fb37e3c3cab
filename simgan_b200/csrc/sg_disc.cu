// GAIL discriminator on the device (A2C/algo/gail.py).
//
//   sg_disc_update          update_gail_dyn (gail.py:154-193): BCE-with-logits on expert (label 1) and
//                           policy (label 0) rows + the WGAN-GP style gradient penalty on mixup rows
//                           (gail.py:67-89) with a HAND-DERIVED double backward (no autograd in a kernel),
//                           + Adam.  One persistent cooperative kernel runs all minibatches of an epoch:
//                           phase A (tile: forward / double backward -> per-CTA partial gradient), grid
//                           barrier, phase B (CTA c sums slice c of the partials in a fixed order and
//                           applies Adam to that slice -- the discriminator has no gradient clipping, so
//                           no global norm is needed), grid barrier.  Variants: register-resident
//                           (disc_reg_kernel, sg_disc_reg.cuh: hidden widths 48/64/100/128; W2 lives in
//                           registers, refreshed through a TMA-staged shared copy after every Adam step;
//                           one row triple per tile while the minibatch has at most one per SM, else two),
//                           resident (natural shared-memory image), persistent (weights through L2) and
//                           phased (one launch per phase; NCCL data-parallel path).
//   sg_disc_predict_reward  predict_reward_combined (gail.py:201-210) for one (N,F) block
//   sg_disc_relabel         the whole T-step relabel loop of main_gail_dyn_ppo.py:275-297, including the
//                           float64 RunningMeanStd merge (baselines/common/running_mean_std.py:33-56),
//                           without a host round trip per step.
//
// Gradient-penalty backward.  Forward on x^: z1=W1x^+b1, h1=tanh z1, z2=W2h1+b2,
// h2=tanh z2, d=w3.h2+b3.  Input gradient: u2=w3*(1-h2^2), v1=W2^T u2, u1=v1*(1-h1^2), g=W1^T u1,
// n=|g|, penalty=lambda*mean((n-1)^2).  With gbar=(2 lambda/B)(n-1) g/n:
//   dW1 += u1 gbar^T ;  ub1 = W1 gbar ;  vb1 = ub1*(1-h1^2) ;  hb1 = -2 ub1*v1*h1
//   dW2 += u2 vb1^T  ;  ub2 = W2 vb1  ;  dw3 += ub2*(1-h2^2) ;  hb2 = -2 ub2*w3*h2 ;  zb2 = hb2*(1-h2^2)
//   dW2 += zb2 h1^T  ;  db2 += zb2    ;  hb1 += W2^T zb2     ;  zb1 = hb1*(1-h1^2)
//   dW1 += zb1 x^^T  ;  db1 += zb1
#include "sg_common.cuh"
#include <string.h>

#include "sg_disc_reg.cuh"
#include "sg_dp.cuh"

namespace sg {

DpView dp_view(const void* ctx);

constexpr int kTB = 2;            // (expert, policy, mixup) triples per CTA tile
constexpr int kDR = 4 * kTB;      // row slots of a tile: [expert | policy | mixup | penalty-pair]

struct DiscArgs {
    int F, H, P, B, nsteps, row_begin, row_end, ntiles, nslots, SL;
    float gp_lambda, one_minus_b1, b2, one_minus_b2, eps;
    DiscLayout L;
    float *params, *m, *v;
    const float *expert, *policy, *alpha;
    const int32_t *eidx, *pidx;
    const float *step_size, *bc2_sqrt;
    float* trace;
    float *gpart, *grad, *losspart;
    unsigned int* bar;
    long long* prof;
    int first_adam_step;
    int dp_on;            // fused peer-memory gradient exchange (sg_dp.cuh)
    DpView dp;
};

struct DiscSmem {
    float *X, *H1, *H2, *D, *DD, *LOSS, *Y2t, *L2t, *L1t, *U1x, *Z2x, *V1, *HB1, *C3, *SCR;
    int ldf, ldh;
    __host__ __device__ static int floats(int F, int H) {
        const int ldf = round_up(F, 4), ldh = round_up(H, 4);
        return kDR * ldf + 2 * kDR * ldh + 4 * kDR + 3 * kDR * ldh + 2 * kTB * ldh + 3 * kTB * ldh + kStepThreads * kDR * 4;
    }
    __device__ void carve(float* sm, int F, int H) {
        ldf = round_up(F, 4); ldh = round_up(H, 4);
        X = sm; sm += kDR * ldf;
        H1 = sm; sm += kDR * ldh;
        H2 = sm; sm += kDR * ldh;
        D = sm; sm += kDR;
        DD = sm; sm += kDR;
        LOSS = sm; sm += 2 * kDR;
        Y2t = sm; sm += kDR * ldh;    // [H][kDR]
        L2t = sm; sm += kDR * ldh;
        L1t = sm; sm += kDR * ldh;
        U1x = sm; sm += kTB * ldh;    // [H][kTB]
        Z2x = sm; sm += kTB * ldh;
        V1 = sm; sm += kTB * ldh;     // [kTB][ldh]
        HB1 = sm; sm += kTB * ldh;
        C3 = sm; sm += kTB * ldh;
        SCR = sm;
    }
};

__device__ __forceinline__ float softplusf(float z) { return fmaxf(z, 0.f) + log1pf(expf(-fabsf(z))); }
__device__ __forceinline__ float sigmoidf(float z) { return __fdiv_rn(1.f, 1.f + expf(-z)); }

// forward trunk for the rows of X (row-major [R][ldf]) -> H1, H2 (row-major) and logits D
template <int R, class WL>
__device__ __forceinline__ void disc_tile_forward(const float* __restrict__ params, const DiscLayout& L, int F, int H,
                                                  const float* X, int ldf, float* H1, float* H2, int ldh, float* D, int tid) {
    const float* W1 = params + L.w1; const float* B1 = params + L.b1;
    const float* W2 = params + L.w2; const float* B2 = params + L.b2;
    const float* W3 = params + L.w3; const float* B3 = params + L.b3;
    auto e1 = [&](int r, int n, float s) { H1[r * ldh + n] = tanhf(s + WL::ld(B1 + n)); };
    if ((F & 3) == 0) gemm_xwT<R, 4, WL>(W1, X, ldf, H, F, tid, kStepThreads, e1);
    else gemm_xwT<R, 1, WL>(W1, X, ldf, H, F, tid, kStepThreads, e1);
    __syncthreads();
    auto e2 = [&](int r, int n, float s) { H2[r * ldh + n] = tanhf(s + WL::ld(B2 + n)); };
    if ((H & 3) == 0) gemm_xwT<R, 4, WL>(W2, H1, ldh, H, H, tid, kStepThreads, e2);
    else gemm_xwT<R, 1, WL>(W2, H1, ldh, H, H, tid, kStepThreads, e2);
    __syncthreads();
    auto e3 = [&](int r, int n, float s) { D[r] = s + WL::ld(B3); };
    if ((H & 3) == 0) gemm_xwT<R, 4, WL>(W3, H2, ldh, 1, H, tid, kStepThreads, e3);
    else gemm_xwT<R, 1, WL>(W3, H2, ldh, 1, H, tid, kStepThreads, e3);
    __syncthreads();
}

template <class WL>
__device__ void disc_tile(const DiscArgs& a, const float* __restrict__ W, int step, int tile, float* __restrict__ gout,
                          float* __restrict__ lossout, DiscSmem& sm, bool acc) {
    constexpr int R = kDR, TB = kTB;
    const int tid = threadIdx.x, nth = kStepThreads;
    const int F = a.F, H = a.H, ldf = sm.ldf, ldh = sm.ldh;
    const bool vecH = (H & 3) == 0, vecF = (F & 3) == 0;
    const int row0 = a.row_begin + tile * TB;
    const int32_t* eidx = a.eidx + (size_t)step * a.B;
    const int32_t* pidx = a.pidx + (size_t)step * a.B;
    const float* alpha = a.alpha + (size_t)step * a.B;
    const float invB = 1.f / (float)a.B;
    const float* W1 = W + a.L.w1; const float* W2 = W + a.L.w2; const float* W3 = W + a.L.w3;

    // rows: [0,TB) expert, [TB,2TB) policy, [2TB,3TB) mixup = alpha*e + (1-alpha)*p (gail.py:72-75), rest zero
    for (int e = tid; e < R * ldf; e += nth) {
        const int r = e / ldf, k = e - r * ldf;
        const int kind = r / TB, j = r - kind * TB;
        const int row = row0 + j;
        float x = 0.f;
        if (kind < 3 && row < a.row_end && k < F) {
            const float xe = a.expert[(size_t)eidx[row] * F + k];
            const float xp = a.policy[(size_t)pidx[row] * F + k];
            const float al = alpha[row];
            x = kind == 0 ? xe : (kind == 1 ? xp : __fadd_rn(__fmul_rn(al, xe), __fmul_rn(__fsub_rn(1.f, al), xp)));
        }
        sm.X[e] = x;
    }
    __syncthreads();
    disc_tile_forward<R, WL>(W, a.L, F, H, sm.X, ldf, sm.H1, sm.H2, ldh, sm.D, tid);

    // BCE-with-logits seeds (gail.py:171-176): expert target 1, policy target 0, both batch means
    if (tid < R) {
        const int kind = tid / TB, j = tid - kind * TB;
        const bool ok = (row0 + j) < a.row_end;
        float dd = 0.f, le = 0.f, lp = 0.f;
        if (ok && kind == 0) { const float d = sm.D[tid]; dd = (sigmoidf(d) - 1.f) * invB; le = softplusf(-d); }
        if (ok && kind == 1) { const float d = sm.D[tid]; dd = sigmoidf(d) * invB; lp = softplusf(d); }
        sm.DD[tid] = dd; sm.LOSS[tid] = le; sm.LOSS[R + tid] = lp;
    }
    __syncthreads();
    // Y2t = [dz2_e | dz2_p | u2 | 0] ;  L2t slots e,p and the u2 slot are final here
    for (int e = tid; e < H * R; e += nth) {
        const int n = e / R, r = e - n * R;
        const int kind = r / TB;
        const float h2 = sm.H2[r * ldh + n];
        const float w3 = WL::ld(W3 + n);
        const float base = w3 * (1.f - h2 * h2);
        float y = 0.f;
        if (kind < 2) y = sm.DD[r] * base;
        else if (kind == 2) y = base;                // u2
        sm.Y2t[e] = y;
        if (kind < 2) sm.L2t[e] = y;
        else if (kind == 2) sm.L2t[n * R + r + TB] = base;   // u2 pairs with vb1 in slot 3TB+j
    }
    __syncthreads();
    // pass A: (Y2t . W2): e/p rows -> dz1 ; mixup rows -> v1, u1
    auto epiA = [&](int r, int k, float s) {
        const int kind = r / TB, j = r - kind * TB;
        const float h1 = sm.H1[r * ldh + k];
        const float t1 = s * (1.f - h1 * h1);
        if (kind < 2) sm.L1t[k * R + r] = t1;                      // dz1
        else if (kind == 2) { sm.V1[j * ldh + k] = s; sm.U1x[k * TB + j] = t1; sm.L1t[k * R + r + TB] = t1; }  // u1 -> slot 3TB+j
    };
    if (vecH) gemm_yW<R, 4, WL>(W2, sm.Y2t, H, H, sm.SCR, tid, nth, epiA);
    else gemm_yW<R, 1, WL>(W2, sm.Y2t, H, H, sm.SCR, tid, nth, epiA);
    // pass B: g = u1 . W1  (input gradient of the mixup rows) -> X rows 3TB+j (raw g for now)
    float* G = sm.X + 3 * TB * ldf;
    auto epiB = [&](int r, int k, float s) { G[r * ldf + k] = s; };
    if (vecF) gemm_yW<TB, 4, WL>(W1, sm.U1x, H, F, sm.SCR, tid, nth, epiB);
    else gemm_yW<TB, 1, WL>(W1, sm.U1x, H, F, sm.SCR, tid, nth, epiB);
    // ||g||, penalty and gbar = (2 lambda / B)(n-1) g / n
    if (tid < 32 * TB) {
        const int j = tid >> 5, lane = tid & 31;
        float ssq = 0.f;
        for (int k = lane; k < F; k += 32) { const float g = G[j * ldf + k]; ssq += g * g; }
        ssq = warp_sum(ssq);
        const float nrm = sqrtf(ssq);
        const bool ok = (row0 + j) < a.row_end;
        const float coef = (ok && nrm > 0.f) ? (2.f * a.gp_lambda * invB) * (nrm - 1.f) / nrm : 0.f;
        for (int k = lane; k < ldf; k += 32) G[j * ldf + k] = (k < F) ? coef * G[j * ldf + k] : 0.f;
        if (lane == 0) sm.LOSS[2 * TB + j] = ok ? (nrm - 1.f) * (nrm - 1.f) : 0.f;   // LOSS[2TB..3TB) : penalty terms
    }
    __syncthreads();
    // pass C: ub1 = gbar . W1^T -> vb1 (H1 rows 3TB+j), hb1
    auto epiC = [&](int r, int n, float s) {
        const float h1 = sm.H1[(2 * TB + r) * ldh + n];
        sm.H1[(3 * TB + r) * ldh + n] = s * (1.f - h1 * h1);                 // vb1
        sm.HB1[r * ldh + n] = -2.f * s * sm.V1[r * ldh + n] * h1;            // hb1
    };
    if (vecF) gemm_xwT<TB, 4, WL>(W1, G, ldf, H, F, tid, nth, epiC);
    else gemm_xwT<TB, 1, WL>(W1, G, ldf, H, F, tid, nth, epiC);
    __syncthreads();
    // pass D: ub2 = vb1 . W2^T -> dw3 term, zb2
    const float* VB1 = sm.H1 + 3 * TB * ldh;
    auto epiD = [&](int r, int n, float s) {
        const float h2 = sm.H2[(2 * TB + r) * ldh + n];
        const float om = 1.f - h2 * h2;
        sm.C3[r * ldh + n] = s * om;
        const float zb2 = -2.f * s * WL::ld(W3 + n) * h2 * om;
        sm.Z2x[n * TB + r] = zb2;
        sm.L2t[n * R + 2 * TB + r] = zb2;
    };
    if (vecH) gemm_xwT<TB, 4, WL>(W2, VB1, ldh, H, H, tid, nth, epiD);
    else gemm_xwT<TB, 1, WL>(W2, VB1, ldh, H, H, tid, nth, epiD);
    __syncthreads();
    // pass E: hb1 += zb2 . W2 ; zb1 = hb1*(1-h1^2)
    auto epiE = [&](int r, int k, float s) {
        const float h1 = sm.H1[(2 * TB + r) * ldh + k];
        sm.L1t[k * R + 2 * TB + r] = (sm.HB1[r * ldh + k] + s) * (1.f - h1 * h1);
    };
    if (vecH) gemm_yW<TB, 4, WL>(W2, sm.Z2x, H, H, sm.SCR, tid, nth, epiE);
    else gemm_yW<TB, 1, WL>(W2, sm.Z2x, H, H, sm.SCR, tid, nth, epiE);

    // parameter gradients of this tile
    if (vecH) outer_store<R, 4>(gout + a.L.w2, sm.L2t, sm.H1, ldh, H, H, tid, nth, acc);
    else outer_store<R, 1>(gout + a.L.w2, sm.L2t, sm.H1, ldh, H, H, tid, nth, acc);
    rowsum_store<R>(gout + a.L.b2, sm.L2t, H, tid, nth, acc, 3 * TB);
    if (vecF) outer_store<R, 4>(gout + a.L.w1, sm.L1t, sm.X, ldf, H, F, tid, nth, acc);
    else outer_store<R, 1>(gout + a.L.w1, sm.L1t, sm.X, ldf, H, F, tid, nth, acc);
    rowsum_store<R>(gout + a.L.b1, sm.L1t, H, tid, nth, acc, 3 * TB);
    for (int n = tid; n < H; n += nth) {
        float s = 0.f;
        for (int r = 0; r < 2 * TB; ++r) s += sm.DD[r] * sm.H2[r * ldh + n];
        for (int j = 0; j < TB; ++j) s += sm.C3[j * ldh + n];
        __stcg(gout + a.L.w3 + n, acc ? s + __ldcg(gout + a.L.w3 + n) : s);
    }
    if (tid == 0) {
        float db3 = 0.f, le = 0.f, lp = 0.f, gp = 0.f;
        for (int r = 0; r < 2 * TB; ++r) db3 += sm.DD[r];
        for (int r = 0; r < TB; ++r) { le += sm.LOSS[r]; lp += sm.LOSS[R + TB + r]; gp += sm.LOSS[2 * TB + r]; }
        if (acc) { db3 += __ldcg(gout + a.L.b3); le += lossout[0]; lp += lossout[1]; gp += lossout[2]; }
        __stcg(gout + a.L.b3, db3);
        lossout[0] = le; lossout[1] = lp; lossout[2] = gp;
    }
    __syncthreads();
}

// ---- phase B: slice `cta` of the flat gradient (+ loss sums by CTA 0) --------------------------------
// Returns true when thread tid < n4 holds float4 tid of the reduced slice in `mine` (narrow slices).
// `loss_smem` (3 floats, optional): CTA 0 also leaves the loss sums there for its own trace write (no L2 round trip).
template <int NT = kStepThreads, bool KEEP = false>
__device__ bool disc_reduce_slice(const DiscArgs& a, int cta, float4* scr4, float4& mine, float* loss_smem = nullptr) {
    const int tid = threadIdx.x;
    const int p0 = min(a.P, cta * a.SL), p1 = min(a.P, p0 + a.SL);
    // CTA 0 also sums the loss partials: its LAST warp issues those loads first so that they ride along with the
    // slice loads instead of adding a serial L2 round trip to the CTA every other CTA waits for.
    constexpr int LW = NT / 32 - 1, MAXC = 5;              // up to 160 slots (>= #SMs)
    const bool lossw = cta == 0 && (tid >> 5) == LW;
    float l0[MAXC], l1[MAXC], l2[MAXC];
    if (lossw) {
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c = (tid & 31) + 32 * i;
            const bool ok = c < a.nslots;
            l0[i] = ok ? ld_cg(a.losspart + c * 4) : 0.f;
            l1[i] = ok ? ld_cg(a.losspart + c * 4 + 1) : 0.f;
            l2[i] = ok ? ld_cg(a.losspart + c * 4 + 2) : 0.f;
        }
    }
    const bool narrow = reduce_partials_slice<NT, KEEP>(a.gpart, (size_t)a.P, a.nslots, p0, p1, a.grad, scr4, tid, mine);
    if (lossw) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < MAXC; ++i) { s0 += l0[i]; s1 += l1[i]; s2 += l2[i]; }
        s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
        if ((tid & 31) == 0) {
            __stcg(a.grad + a.P, s0); __stcg(a.grad + a.P + 1, s1); __stcg(a.grad + a.P + 2, s2);
            if (loss_smem) { loss_smem[0] = s0; loss_smem[1] = s1; loss_smem[2] = s2; }
        }
    }
    return narrow;
}

// ---- Adam on slice `cta` (the discriminator has no gradient clipping: A2C/algo/gail.py:186-188) --------
// `have`: thread tid < n4 holds float4 tid of the slice's gradient in `mine` (fused path).
__device__ void disc_adam_slice(const DiscArgs& a, int step, int cta, bool have, float4 mine) {
    const int tid = threadIdx.x;
    const float ss = a.step_size[step], bc2 = a.bc2_sqrt[step];
    const int p0 = min(a.P, cta * a.SL), p1 = min(a.P, p0 + a.SL);
    if (have) {
        const int pq = p0 + 4 * tid;
        if (pq < p1) {
            const float4 w4 = ld_cg4(a.params + pq), m4 = ld_cg4(a.m + pq), v4 = ld_cg4(a.v + pq);
            float w[4] = {w4.x, w4.y, w4.z, w4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
            const float g[4] = {mine.x, mine.y, mine.z, mine.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) adam_update(w[i], m[i], v[i], g[i], a.one_minus_b1, a.b2, a.one_minus_b2, ss, bc2, a.eps);
            __stcg(reinterpret_cast<float4*>(a.params + pq), make_float4(w[0], w[1], w[2], w[3]));
            __stcg(reinterpret_cast<float4*>(a.m + pq), make_float4(m[0], m[1], m[2], m[3]));
            __stcg(reinterpret_cast<float4*>(a.v + pq), make_float4(v[0], v[1], v[2], v[3]));
        }
    } else {
        for (int p = p0 + tid; p < p1; p += kStepThreads) {
            const float g = ld_cg(a.grad + p);
            float pv = __ldcg(a.params + p), mv = __ldcg(a.m + p), vv = __ldcg(a.v + p);
            adam_update(pv, mv, vv, g, a.one_minus_b1, a.b2, a.one_minus_b2, ss, bc2, a.eps);
            __stcg(a.params + p, pv); __stcg(a.m + p, mv); __stcg(a.v + p, vv);
        }
    }
    if (cta == 0) {
        __syncthreads();                       // loss sums were stored by the last warp (fused path)
        if (tid == 0) {
            const float invB = 1.f / (float)a.B;
            const float le = ld_cg(a.grad + a.P) * invB, lp = ld_cg(a.grad + a.P + 1) * invB;
            const float gp = a.gp_lambda * ld_cg(a.grad + a.P + 2) * invB;
            float* tr = a.trace + (size_t)step * 3;
            tr[0] = (le + lp) + gp; tr[1] = le; tr[2] = lp;
        }
    }
}

// Phase B of the register-resident kernel: slice reduction + Adam with ONE parameter per thread.  The
// parameter / moment loads are issued before the partial sums are fetched (one L2 round trip covers both) and
// the reduced gradient is handed over through shared memory.  Same arithmetic per element as disc_adam_slice.
template <int NT>
__device__ void disc_reduce_adam_fused(const DiscArgs& a, int step, int cta, float4* scr4) {
    const int tid = threadIdx.x;
    const int p0 = min(a.P, cta * a.SL), p1 = min(a.P, p0 + a.SL);
    const int n4 = (p1 - p0) >> 2;
    const bool narrow = n4 <= 128 && (p1 - p0) <= NT;       // one parameter per thread, gradient via shared memory
    const int pe = p0 + tid;
    const bool own = narrow && pe < p1;
    float pv = 0.f, mv = 0.f, vv = 0.f;
    if (own) { pv = __ldcg(a.params + pe); mv = __ldcg(a.m + pe); vv = __ldcg(a.v + pe); }
    float4 mine;
    __shared__ float loss_smem[4];
    disc_reduce_slice<NT, true>(a, cta, scr4, mine, loss_smem);
    if (a.dp_on)
        dp_exchange_slice<NT>(a.dp, a.grad, p0, p1, cta, (unsigned int)(a.first_adam_step + step),
                              n4 <= 128 ? reinterpret_cast<float*>(scr4 + NT) : nullptr);
    const float ss = a.step_size[step], bc2 = a.bc2_sqrt[step];
    if (narrow) {
        if (own) {
            const float g = reinterpret_cast<const float*>(scr4 + NT)[tid];
            adam_update(pv, mv, vv, g, a.one_minus_b1, a.b2, a.one_minus_b2, ss, bc2, a.eps);
            __stcg(a.params + pe, pv); __stcg(a.m + pe, mv); __stcg(a.v + pe, vv);
        }
    } else {
        for (int p = p0 + tid; p < p1; p += NT) {
            const float g = ld_cg(a.grad + p);
            float w = __ldcg(a.params + p), m = __ldcg(a.m + p), v = __ldcg(a.v + p);
            adam_update(w, m, v, g, a.one_minus_b1, a.b2, a.one_minus_b2, ss, bc2, a.eps);
            __stcg(a.params + p, w); __stcg(a.m + p, m); __stcg(a.v + p, v);
        }
    }
    if (cta == 0) {
        __syncthreads();                       // loss sums were stored by the last warp
        if (tid == 0) {
            const float invB = 1.f / (float)a.B;
            const float le = loss_smem[0] * invB, lp = loss_smem[1] * invB;
            const float gp = a.gp_lambda * loss_smem[2] * invB;
            float* tr = a.trace + (size_t)step * 3;
            tr[0] = (le + lp) + gp; tr[1] = le; tr[2] = lp;
        }
    }
}

template <class WL>
__device__ __forceinline__ void disc_phaseA(const DiscArgs& a, const float* W, int step, int cta, int ncta, float* smem) {
    DiscSmem sm;
    sm.carve(smem, a.F, a.H);
    bool acc = false;
    for (int tile = cta; tile < a.ntiles; tile += ncta) {
        disc_tile<WL>(a, W, step, tile, a.gpart + (size_t)cta * a.P, a.losspart + cta * 4, sm, acc);
        acc = true;
    }
}

__device__ __forceinline__ void disc_poison_on_timeout(const DiscArgs& a) {
    // a timed-out grid barrier poisons the trace so the host raises instead of trusting the result
    if (blockIdx.x == 0 && threadIdx.x == 0 && *(volatile unsigned int*)(a.bar + 1) != 0u) a.trace[0] = __int_as_float(0x7fc00000);
}

// RESIDENT: every CTA refreshes a private shared-memory image of the parameters at the start of each step
// and the tile phase reads its weights from there; otherwise weights come from global memory through L2.
template <bool RESIDENT>
__global__ void __launch_bounds__(kStepThreads, 1) disc_persistent_kernel(DiscArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;
    float* tile = RESIDENT ? smem + a.P : smem;
    GridBarrier gb{a.bar, a.bar + 1, gridDim.x, 0};
    PhaseClock pc{a.prof + 8 * blockIdx.x, threadIdx.x == 0};
    pc.start();
    for (int step = 0; step < a.nsteps; ++step) {
        if (RESIDENT) {
            load_param_image(Ws, a.params, a.P, threadIdx.x);
            pc.lap(0);
            disc_phaseA<LdShared>(a, Ws, step, blockIdx.x, gridDim.x, tile);
        } else {
            disc_phaseA<LdGlobal>(a, a.params, step, blockIdx.x, gridDim.x, tile);
        }
        pc.lap(1);
        gb.sync();
        pc.lap(2);
        float4 mine;
        bool have = disc_reduce_slice(a, blockIdx.x, reinterpret_cast<float4*>(tile), mine);
        if (a.dp_on) {
            const int p0 = min(a.P, (int)blockIdx.x * a.SL), p1 = min(a.P, p0 + a.SL);
            dp_exchange_slice<kStepThreads>(a.dp, a.grad, p0, p1, blockIdx.x, (unsigned int)(a.first_adam_step + step), nullptr);
            have = false;                  // re-read the exchanged totals from the flat gradient
        }
        disc_adam_slice(a, step, blockIdx.x, have, mine);
        pc.lap(3);
        gb.sync();
        pc.lap(4);
    }
    disc_poison_on_timeout(a);
}
// Register-resident variant (H = 4*HQ known at compile time, H <= 128): W2 lives in registers, the rest of the
// parameters in a small shared image; both are refreshed from global memory after every Adam step.
// TB = row triples per tile: 2 (disc_tile_reg), or 1 (disc_tile_reg1) when that still leaves at most one tile per SM.
template <int HQ, int TB, bool MULTI>
__global__ void __launch_bounds__(kStepThreads, 1) disc_reg_kernel(DiscArgs a) {
    extern __shared__ __align__(16) float smem[];
    const DiscRegImage I = make_disc_reg_image(a.F, a.H);
    float* img = smem;
    float* tile = smem + I.total;
    float* stage = tile + DiscRegSmem::floats(a.F, a.H);      // H*H floats: coalesced landing zone of W2
    DiscRegSmem sm;
    sm.carve(tile, a.F, a.H);
    DiscRegW2<HQ> w;
    __shared__ __align__(8) unsigned long long fill_bar;
    if (threadIdx.x == 0) mbar_init(&fill_bar, 1);
    __syncthreads();
    GridBarrier gb{a.bar, a.bar + 1, gridDim.x, 0};
    PhaseClock pc{a.prof + 8 * blockIdx.x, threadIdx.x == 0};
    pc.start();
    DiscPrefetch pf;
    pf.on = a.ntiles <= (int)gridDim.x && 4 * TB * round_up(a.F, 4) <= kStepThreads;
    pf.valid = false;
    pf.ei = pf.pi = -1;
    pf.xe = pf.xp = pf.al = 0.f;
    for (int step = 0; step < a.nsteps; ++step) {
        disc_reg_fill<HQ>(w, img, stage, a.params, a.L, I, threadIdx.x, &fill_bar, (unsigned int)(step & 1));
        pc.lap(0);
        bool acc = false;
        for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
            if (TB == 1) disc_tile_reg1<HQ>(a, w, img, I, step, t, a.gpart + (size_t)blockIdx.x * a.P, a.losspart + blockIdx.x * 4, sm, acc, pf);
            else disc_tile_reg<HQ, MULTI>(a, w, img, I, step, t, a.gpart + (size_t)blockIdx.x * a.P, a.losspart + blockIdx.x * 4, sm, acc, pf);
            acc = true;
        }
        pc.lap(1);
        gb.sync();
        pc.lap(2);
        disc_reduce_adam_fused<kStepThreads>(a, step, blockIdx.x, reinterpret_cast<float4*>(tile));
        pc.lap(3);
        gb.sync();
        pc.lap(4);
    }
    disc_poison_on_timeout(a);
}

__global__ void __launch_bounds__(kStepThreads, 1) disc_phaseA_kernel(DiscArgs a, int step) {
    extern __shared__ __align__(16) float smem[];
    disc_phaseA<LdGlobal>(a, a.params, step, blockIdx.x, gridDim.x, smem);
}
__global__ void __launch_bounds__(kStepThreads) disc_phaseB_kernel(DiscArgs a) {
    __shared__ float4 scr4[kStepThreads];
    float4 mine;
    disc_reduce_slice(a, blockIdx.x, scr4, mine);
}
__global__ void __launch_bounds__(kStepThreads) disc_phaseC_kernel(DiscArgs a, int step) {
    disc_adam_slice(a, step, blockIdx.x, false, make_float4(0.f, 0.f, 0.f, 0.f));
}

// ---- reward prediction ---------------------------------------------------------------------------
// raw reward of predict_reward_combined (gail.py:203-205) for n_rows rows of (.,F); optionally the
// one-step return update (gail.py:206-209) when `returns` is given (single-block use).
__global__ void __launch_bounds__(kStepThreads) disc_reward_kernel(const float* __restrict__ params, DiscLayout L, int F,
                                                                   int H, const float* __restrict__ d_in, int n_rows,
                                                                   float offset, float* __restrict__ reward,
                                                                   float* __restrict__ returns, const float* __restrict__ masks,
                                                                   float gamma, int has_returns) {
    constexpr int R = kRows;
    extern __shared__ __align__(16) float smem[];
    const int ldf = round_up(F, 4), ldh = round_up(H, 4);
    float* Ws = smem;                                   // parameter image: read once per CTA, reused by every tile
    float* X = Ws + L.total; float* H1 = X + R * ldf; float* H2 = H1 + R * ldh; float* D = H2 + R * ldh;
    const int tid = threadIdx.x;
    load_param_image(Ws, params, L.total, tid);
    for (int tile = blockIdx.x; tile * R < n_rows; tile += gridDim.x) {
        const int row0 = tile * R;
        for (int e = tid; e < R * ldf; e += kStepThreads) {
            const int r = e / ldf, k = e - r * ldf;
            X[e] = (row0 + r < n_rows && k < F) ? d_in[(size_t)(row0 + r) * F + k] : 0.f;
        }
        __syncthreads();
        disc_tile_forward<R, LdShared>(Ws, L, F, H, X, ldf, H1, H2, ldh, D, tid);
        if (tid < R && row0 + tid < n_rows) {
            const int row = row0 + tid;
            const float s = sigmoidf(D[tid]);
            // (s + 1e-7).log() - (1 - s + 1e-7).log() + offset
            const float rew = __fadd_rn(__fsub_rn(logf(__fadd_rn(s, 1e-7f)), logf(__fadd_rn(__fsub_rn(1.f, s), 1e-7f))), offset);
            reward[row] = rew;
            if (returns) returns[row] = has_returns ? __fadd_rn(__fmul_rn(__fmul_rn(returns[row], gamma), masks[row]), rew) : rew;
        }
        __syncthreads();
    }
}

// returns_t = returns_{t-1}*gamma*masks[t] + raw_t per env column; keeps every step's returns for the
// running statistics (gail.py:206-209 called with masks[step], main_gail_dyn_ppo.py:276-280).
// One CTA per 32 columns; all threads stream chunks of steps through double-buffered shared memory with
// cp.async while the first `cols` threads walk the recurrence (see returns_scan_staged_kernel).
constexpr int kRTC = 64;
constexpr int kRlStages = 4;
constexpr size_t kRlScanSmemBytes = (size_t)(2 * kRlStages + 1) * kRTC * 32 * sizeof(float);
__device__ __forceinline__ void rl_cp_async4(float* dst_smem, const float* src) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void rl_cp_async16(float* dst_smem, const float* src) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__global__ void __launch_bounds__(256) relabel_scan_kernel(const float* __restrict__ raw, const float* __restrict__ masks,
                                                           float* __restrict__ ret_all, float* __restrict__ disc_returns,
                                                           int T, int N, float gamma, int has_returns) {
    extern __shared__ __align__(16) float rl_scan_smem[];
    typedef float (*Stage)[kRTC][32];
    Stage sR = reinterpret_cast<Stage>(rl_scan_smem);
    Stage sM = sR + kRlStages;
    float (*sO)[32] = reinterpret_cast<float (*)[32]>(sM + kRlStages);
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * 32;
    const int cols = min(32, N - n0);
    const int nchunks = (T + kRTC - 1) / kRTC;
    // 16-byte copies when every row segment is float4-aligned (see returns_scan_staged_kernel)
    const bool vec = (N & 3) == 0 && (reinterpret_cast<unsigned long long>(raw) & 15ull) == 0 &&
                     (reinterpret_cast<unsigned long long>(masks) & 15ull) == 0;
    auto issue = [&](int k) {       // always commits a (possibly empty) group, so that group k <-> chunk k
        const int b = k % kRlStages;
        if (k < nchunks && vec) {
            const int c4 = 4 * (tid & 7);
            if (c4 < cols) {
                for (int tt = tid >> 3; tt < kRTC; tt += 32) {
                    const int t = k * kRTC + tt;
                    if (t < T) {
                        const size_t i0 = (size_t)t * N + n0 + c4;
                        rl_cp_async16(&sR[b][tt][c4], raw + i0);
                        rl_cp_async16(&sM[b][tt][c4], masks + i0);
                    }
                }
            }
        } else if (k < nchunks) {
            for (int e = tid; e < kRTC * 32; e += 256) {
                const int tt = e >> 5, c = e & 31;
                const int t = k * kRTC + tt;
                if (c < cols && t < T) {
                    rl_cp_async4(&sR[b][tt][c], raw + (size_t)t * N + n0 + c);
                    rl_cp_async4(&sM[b][tt][c], masks + (size_t)t * N + n0 + c);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    float ret = 0.f;
    if (tid < cols && has_returns) ret = disc_returns[n0 + tid];
    for (int k = 0; k < kRlStages - 1; ++k) issue(k);
    for (int k = 0; k < nchunks; ++k) {
        issue(k + kRlStages - 1);         // refills the buffer chunk k-1 was read from (all threads passed its barrier)
        asm volatile("cp.async.wait_group %0;" ::"n"(kRlStages - 1) : "memory");
        __syncthreads();
        const int b = k % kRlStages;
        if (tid < cols) {
            const int len = min(kRTC, T - k * kRTC);
            int tt = 0;
            if (k == 0 && !has_returns) { ret = sR[b][0][tid]; sO[0][tid] = ret; tt = 1; }     // the first ever call clones (gail.py:205-206)
            // groups of 8 steps: all shared-memory loads of a group are issued before its dependent chain starts
            for (; tt + 8 <= len; tt += 8) {
                float m[8], r[8], o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { m[j] = sM[b][tt + j][tid]; r[j] = sR[b][tt + j][tid]; }
#pragma unroll
                for (int j = 0; j < 8; ++j) { ret = __fadd_rn(__fmul_rn(__fmul_rn(ret, gamma), m[j]), r[j]); o[j] = ret; }
#pragma unroll
                for (int j = 0; j < 8; ++j) sO[tt + j][tid] = o[j];
            }
            for (; tt < len; ++tt) {
                ret = __fadd_rn(__fmul_rn(__fmul_rn(ret, gamma), sM[b][tt][tid]), sR[b][tt][tid]);
                sO[tt][tid] = ret;
            }
        }
        __syncthreads();
        for (int e = tid; e < kRTC * 32; e += 256) {
            const int tt = e >> 5, c = e & 31;
            const int t = k * kRTC + tt;
            if (c < cols && t < T) ret_all[(size_t)t * N + n0 + c] = sO[tt][c];
        }
    }
    if (tid < cols) disc_returns[n0 + tid] = ret;
}

// numpy's float32 pairwise summation (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum), so that
// np.mean / np.var of the N per-step returns are reproduced bit for bit.  SQ: sum (a[i]-c)^2 instead.
template <bool SQ>
__device__ float np_pairwise_sum(const float* a, int n, float c) {
    auto f = [&](int i) { if (SQ) { const float d = __fsub_rn(a[i], c); return __fmul_rn(d, d); } return a[i]; };
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, f(i));
        return res;
    } else if (n <= 128) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = f(j);
        int i;
        for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], f(i + j));
        }
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, f(i));
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return __fadd_rn(np_pairwise_sum<SQ>(a, n2, c), np_pairwise_sum<SQ>(a + n2, n - n2, c));
    }
}

// per-step batch moments in numpy's float32 arithmetic: np.mean(x), np.var(x) (running_mean_std.py:34-35)
__global__ void __launch_bounds__(128) relabel_moments_kernel(const float* __restrict__ ret_all, int T, int N,
                                                              float* __restrict__ bmean, float* __restrict__ bvar,
                                                              float* __restrict__ mean_returns) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const float* x = ret_all + (size_t)t * N;
    const float mean = __fdiv_rn(np_pairwise_sum<false>(x, N, 0.f), (float)N);
    const float var = __fdiv_rn(np_pairwise_sum<true>(x, N, mean), (float)N);
    bmean[t] = mean; bvar[t] = var;
    mean_returns[t] = mean;
}

// the sequential float64 Chan merge (running_mean_std.py:45-56); scale[t] = sqrt(var_t + 1e-7).
//
// The merge is an inherently serial chain over the T steps (one thread), and in its plain form every step carries
// three IEEE divisions by tot_t = count_t + N and one square root: ~500 cycles per step, all dependent.  Everything
// that does not depend on the running mean / variance is therefore taken OFF the chain, per chunk of steps:
//   1. thread 0 walks the count chain alone (tot_t depends on nothing else);
//   2. all threads form r_t = RN(1/tot_t) (__drcp_rn) and the batch second moments in parallel;
//   3. thread 0 walks the mean / variance chain, dividing with div_rcp(): q = a*r followed by two fused residual
//      corrections q += (a - q*tot)*r.  With r the correctly rounded reciprocal and the first correction making q
//      faithful, the second one yields exactly RN(a/tot) (Markstein's theorem; needs tot's significand not all ones
//      and no under/overflow in the residuals);
//   4. all threads check that every quotient of the chunk stayed inside that domain (an ineligible tot was given a
//      NaN reciprocal in 2.); if one did not, thread 0 redoes the chunk with plain divisions;
//   5. all threads take the square roots.
// Measured: the chain is bound by the number of DEPENDENT fp64 operations per step (8 on both the mean and the
// variance recurrence now, ~12 with plain divisions), ~45-50 cycles each.
// tests/test_gpu_parity.py::test_relabel_normalize_bitexact pins the result bit for bit against NumPy.
constexpr int kRmsChunk = 512;

__device__ __forceinline__ double div_rcp(double a, double b, double r) {
    double q = __dmul_rn(a, r);
    q = __fma_rn(__fma_rn(-q, b, a), r, q);
    return __fma_rn(__fma_rn(-q, b, a), r, q);
}
// quotient magnitudes for which no residual of div_rcp can have under- or overflowed (NaN fails the test too)
__device__ __forceinline__ bool div_rcp_in_domain(double q) {
    const double m = fabs(q);
    return q == 0.0 || (m > 1e-270 && m < 1e270);
}

// one Chan merge step with plain IEEE divisions, NumPy's op order (no FMA contraction)
__device__ __forceinline__ void rms_step_plain(double bmean, double m_b, double bc, double& mean, double& var, double& count) {
    const double delta = __dsub_rn(bmean, mean);
    const double tot = __dadd_rn(count, bc);
    const double new_mean = __dadd_rn(mean, __ddiv_rn(__dmul_rn(delta, bc), tot));
    const double m_a = __dmul_rn(var, count);
    const double corr = __ddiv_rn(__dmul_rn(__dmul_rn(__dmul_rn(delta, delta), count), bc), tot);
    const double m2 = __dadd_rn(__dadd_rn(m_a, m_b), corr);
    mean = new_mean; var = __ddiv_rn(m2, tot); count = tot;
}

__global__ void __launch_bounds__(256) relabel_rms_kernel(const float* __restrict__ bmean, const float* __restrict__ bvar, int T, int N,
                                                          double* __restrict__ rms, double* __restrict__ scale) {
    __shared__ float sMean[kRmsChunk];
    __shared__ double sMb[kRmsChunk], sTot[kRmsChunk], sRcp[kRmsChunk], sVarOut[kRmsChunk];
    __shared__ double sQ1[kRmsChunk], sQ2[kRmsChunk];      // the other two quotients of a step, for the domain check
    __shared__ int sBad;
    const int tid = threadIdx.x;
    double mean = rms[0], var = rms[1], count = rms[2];
    const double bc = (double)N;
    for (int t0 = 0; t0 < T; t0 += kRmsChunk) {
        const int len = min(kRmsChunk, T - t0);
        for (int e = tid; e < len; e += 256) {
            sMean[e] = bmean[t0 + e];
            sMb[e] = (double)__fmul_rn(bvar[t0 + e], (float)N);      // float32 * int stays float32 in numpy
        }
        if (tid == 0) {
            sBad = 0;
            double c = count;
            for (int i = 0; i < len; ++i) { c = __dadd_rn(c, bc); sTot[i] = c; }
        }
        __syncthreads();
        for (int e = tid; e < len; e += 256) {
            const double tot = sTot[e];
            const unsigned long long bits = (unsigned long long)__double_as_longlong(tot);
            const bool eligible = tot > 1e-270 && tot < 1e270 && (bits & 0xFFFFFFFFFFFFFull) != 0xFFFFFFFFFFFFFull;
            sRcp[e] = eligible ? __drcp_rn(tot) : __longlong_as_double(0x7ff8000000000000ll);     // NaN poisons the step
        }
        __syncthreads();
        const double mean0 = mean, var0 = var, count0 = count;
        if (tid == 0) {
            for (int i = 0; i < len; ++i) {
                // explicit round-to-nearest ops in NumPy's order; only the divisions are restructured (see above)
                const double tot = sTot[i], r = sRcp[i];
                const double delta = __dsub_rn((double)sMean[i], mean);
                const double q1 = div_rcp(__dmul_rn(delta, bc), tot, r);
                const double m_a = __dmul_rn(var, count);
                const double q2 = div_rcp(__dmul_rn(__dmul_rn(__dmul_rn(delta, delta), count), bc), tot, r);
                const double m2 = __dadd_rn(__dadd_rn(m_a, sMb[i]), q2);
                mean = __dadd_rn(mean, q1); var = div_rcp(m2, tot, r); count = tot;
                sQ1[i] = q1; sQ2[i] = q2; sVarOut[i] = var;
            }
        }
        __syncthreads();
        // domain check of every quotient of the chunk, off the chain; a failure anywhere redoes the chunk plainly
        bool bad = false;
        for (int e = tid; e < len; e += 256)
            bad = bad || !div_rcp_in_domain(sQ1[e]) || !div_rcp_in_domain(sQ2[e]) || !div_rcp_in_domain(sVarOut[e]);
        if (bad) sBad = 1;
        __syncthreads();
        if (sBad && tid == 0) {
            mean = mean0; var = var0; count = count0;
            for (int i = 0; i < len; ++i) {
                rms_step_plain((double)sMean[i], sMb[i], bc, mean, var, count);
                sVarOut[i] = var;
            }
        }
        __syncthreads();
        for (int e = tid; e < len; e += 256) scale[t0 + e] = sqrt(__dadd_rn(sVarOut[e], 1e-7));
        __syncthreads();
    }
    if (tid == 0) { rms[0] = mean; rms[1] = var; rms[2] = count; }
}

// Self-check of div_rcp against __ddiv_rn on pseudo-random operands (tests only): returns the number of mismatches.
__global__ void div_rcp_selfcheck_kernel(unsigned long long seed, int per_thread, double b_lo, double b_hi, unsigned long long* mismatches) {
    unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    auto next = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    unsigned long long bad_count = 0;
    for (int i = 0; i < per_thread; ++i) {
        // b: log-uniform in [b_lo, b_hi] with random significand; a: random significand, exponent in [-60, 60]
        const double ub = (double)(next() >> 11) * (1.0 / 9007199254740992.0);
        const double b = b_lo * exp(ub * log(b_hi / b_lo));
        const unsigned long long ma = next() & 0xFFFFFFFFFFFFFull;
        const long long ea = 1023 + (long long)(next() % 121) - 60;
        double a = __longlong_as_double((long long)(((unsigned long long)ea << 52) | ma));
        if (next() & 1) a = -a;
        const double q = div_rcp(a, b, __drcp_rn(b));
        if (div_rcp_in_domain(q) && q != __ddiv_rn(a, b)) ++bad_count;
    }
    if (bad_count) atomicAdd(mismatches, bad_count);
}

// rewards[t] = float32(clip(float64(raw)/sqrt(var_t+1e-7), -10, 10))   (main_gail_dyn_ppo.py:288-292)
__global__ void __launch_bounds__(256) relabel_apply_kernel(const float* __restrict__ raw, const double* __restrict__ scale,
                                                            float* __restrict__ rewards, int T, int N) {
    const long long total = (long long)T * N;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(e / N);
        double v = (double)raw[e] / scale[t];
        v = fmin(fmax(v, -10.0), 10.0);
        rewards[e] = (float)v;
    }
}

constexpr size_t kDiscMaxDynSmem = 227 * 1024 - 1024;
static size_t disc_tile_smem_floats(const sg_disc_config* c) {
    size_t f = (size_t)DiscSmem::floats(c->feat_dim, c->hidden);
    return f < 4 * kStepThreads ? 4 * kStepThreads : f;      // phase B needs 256 float4 of scratch
}
static bool disc_reg_ok(const sg_disc_config* c) {
    const int h = c->hidden;
    return h == 48 || h == 64 || h == 100 || h == 128;      // instantiated widths of disc_reg_kernel
}
static size_t disc_reg_smem_bytes(const sg_disc_config* c) {
    size_t tile = (size_t)DiscRegSmem::floats(c->feat_dim, c->hidden);
    // phase B scratch (NT + 128 float4) aliases the tile arrays + the W2 staging area behind them
    const size_t stage = (size_t)c->hidden * c->hidden;
    if (tile + stage < 4 * (kStepThreads + 128)) tile = 4 * (kStepThreads + 128);
    return ((size_t)make_disc_reg_image(c->feat_dim, c->hidden).total + tile + stage) * sizeof(float);
}
static size_t disc_resident_smem_bytes(const sg_disc_config* c) {
    if (disc_reg_ok(c)) return disc_reg_smem_bytes(c);
    DiscLayout L = make_disc_layout(c->feat_dim, c->hidden);
    return ((size_t)L.total + disc_tile_smem_floats(c)) * sizeof(float);
}
// row triples per tile: the register-resident kernel runs single-triple tiles while that leaves at most one per SM
static int disc_tb(const sg_disc_config* c) {
    int sms = sg_device_sm_count();
    if (sms <= 0) sms = 148;
    int mode = c->mode;
    if (mode == 0) mode = disc_resident_smem_bytes(c) <= kDiscMaxDynSmem ? 3 : 2;
    return (mode == 3 && disc_reg_ok(c) && c->row_end - c->row_begin <= sms) ? 1 : kTB;
}
static int disc_tiles(const sg_disc_config* c) { const int tb = disc_tb(c); return (c->row_end - c->row_begin + tb - 1) / tb; }
static int disc_grid(const sg_disc_config* c, int* sms_out) {
    int sms = sg_device_sm_count();
    if (sms <= 0) sms = 148;
    if (sms_out) *sms_out = sms;
    int tiles = disc_tiles(c);
    int g = tiles < sms ? tiles : sms;
    // never fewer than 64 CTAs: CTAs without a tile skip phase A but still own a slice of the reduce / Adam phases,
    // which keeps those phases on the narrow one-parameter-per-thread path for small (or sharded) minibatches
    const int gmin = sms < 64 ? sms : 64;
    if (g < gmin) g = gmin;
    return g < 1 ? 1 : g;
}
static int disc_validate(const sg_disc_config* c) {
    SG_REQUIRE(c, "sg_disc: null config");
    SG_REQUIRE(c->feat_dim > 0 && c->hidden > 0 && c->batch_size > 0 && c->n_steps > 0, "sg_disc: non-positive sizes");
    SG_REQUIRE(c->row_begin >= 0 && c->row_begin < c->row_end && c->row_end <= c->batch_size,
               "sg_disc: shard [%d,%d) outside minibatch of %d rows", c->row_begin, c->row_end, c->batch_size);
    SG_REQUIRE(c->first_adam_step >= 1, "sg_disc: first_adam_step is 1-based");
    SG_REQUIRE(c->mode >= 0 && c->mode <= 3, "sg_disc: mode must be 0 (auto), 1 (phased), 2 (persistent) or 3 (resident)");
    const size_t smem = disc_tile_smem_floats(c) * sizeof(float);
    SG_REQUIRE(smem <= kDiscMaxDynSmem, "sg_disc: tile needs %zu bytes of shared memory", smem);
    SG_REQUIRE(c->mode != 3 || disc_resident_smem_bytes(c) <= kDiscMaxDynSmem,
               "sg_disc: resident mode needs %zu bytes of shared memory", disc_resident_smem_bytes(c));
    return SG_OK;
}
struct DiscWs { size_t gpart, grad, losspart, bar, prof, total; };
static DiscWs disc_ws(const sg_disc_config* c, int grid) {
    DiscLayout L = make_disc_layout(c->feat_dim, c->hidden);
    DiscWs w;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 255) / 256 * 256; return at; };
    w.gpart = take((size_t)grid * L.total * sizeof(float));
    w.grad = take((size_t)(L.total + 4) * sizeof(float));
    w.losspart = take((size_t)grid * 4 * sizeof(float));
    w.bar = take(2 * sizeof(unsigned int));
    w.prof = take((size_t)grid * 8 * sizeof(long long));
    w.total = o;
    return w;
}

struct RelabelWs { size_t raw, ret_all, bmean, bvar, scale, total; };
static RelabelWs relabel_ws(int T, int N) {
    RelabelWs w;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 255) / 256 * 256; return at; };
    w.raw = take((size_t)T * N * sizeof(float));
    w.ret_all = take((size_t)T * N * sizeof(float));
    w.bmean = take((size_t)T * sizeof(float));
    w.bvar = take((size_t)T * sizeof(float));
    w.scale = take((size_t)T * sizeof(double));
    w.total = o;
    return w;
}

static int launch_reward(const float* params, int F, int H, const float* d_in, int n_rows, float offset, float* reward,
                         float* returns, const float* masks, float gamma, int has_returns, cudaStream_t s) {
    DiscLayout L = make_disc_layout(F, H);
    const size_t smem = (size_t)(L.total + kRows * round_up(F, 4) + 2 * kRows * round_up(H, 4) + kRows) * sizeof(float);
    SG_REQUIRE(smem <= 220 * 1024, "disc reward: tile + parameter image need %zu bytes of shared memory", smem);
    static SmemGrant grant;
    if (int rc = grant_smem(grant, disc_reward_kernel, smem)) return rc;
    int tiles = (n_rows + kRows - 1) / kRows;
    int sms = sg_device_sm_count();
    if (sms <= 0) sms = 148;
    const int per_sm = smem > 100 * 1024 ? 1 : 2;       // resident CTAs by shared memory
    int grid = tiles < per_sm * sms ? tiles : per_sm * sms;
    disc_reward_kernel<<<grid, kStepThreads, smem, s>>>(params, L, F, H, d_in, n_rows, offset, reward, returns, masks, gamma, has_returns);
    count_launches(1);
    SG_CUDA(cudaGetLastError());
    return SG_OK;
}

}  // namespace sg

using namespace sg;

extern "C" {
#pragma GCC visibility push(default)

int64_t sg_disc_workspace_bytes(const sg_disc_config* cfg) {
    if (disc_validate(cfg)) return -1;
    return (int64_t)disc_ws(cfg, disc_grid(cfg, nullptr)).total;
}

int64_t sg_disc_phase_cycles_offset(const sg_disc_config* cfg) {
    if (disc_validate(cfg)) return -1;
    return (int64_t)disc_ws(cfg, disc_grid(cfg, nullptr)).prof;
}

int sg_disc_update(const sg_disc_config* cfg, float* params, float* adam_m, float* adam_v, const float* expert,
                   const float* policy_feat, const int32_t* expert_idx, const int32_t* policy_idx, const float* alpha,
                   const float* step_size, const float* bc2_sqrt, float* trace, void* workspace,
                   sg_allreduce_fn allreduce_cb, void* allreduce_user, void* stream) {
    int rc = disc_validate(cfg);
    if (rc) return rc;
    SG_REQUIRE(params && adam_m && adam_v && expert && policy_feat && expert_idx && policy_idx && alpha && step_size &&
                   bc2_sqrt && trace && workspace, "sg_disc_update: null pointer");
    SG_REQUIRE(!(allreduce_cb && cfg->mode != 1), "sg_disc_update: the allreduce callback needs mode 1");
    SG_REQUIRE(!(cfg->dp_ctx && (cfg->mode == 1 || allreduce_cb)), "sg_disc_update: dp_ctx needs a persistent mode and no callback");
    cudaStream_t s = (cudaStream_t)stream;
    int sms = 0;
    const int grid = disc_grid(cfg, &sms);
    const DiscWs w = disc_ws(cfg, grid);
    char* ws = (char*)workspace;
    DiscArgs a;
    a.F = cfg->feat_dim; a.H = cfg->hidden;
    a.L = make_disc_layout(a.F, a.H);
    a.P = a.L.total; a.B = cfg->batch_size; a.nsteps = cfg->n_steps;
    a.row_begin = cfg->row_begin; a.row_end = cfg->row_end;
    a.ntiles = disc_tiles(cfg);
    a.nslots = grid < a.ntiles ? grid : a.ntiles;
    a.SL = round_up((a.P + grid - 1) / grid, 4);
    a.gp_lambda = (float)cfg->gp_lambda;
    a.one_minus_b1 = (float)(1.0 - cfg->beta1); a.b2 = (float)cfg->beta2; a.one_minus_b2 = (float)(1.0 - cfg->beta2);
    a.eps = (float)cfg->adam_eps;
    a.params = params; a.m = adam_m; a.v = adam_v;
    a.expert = expert; a.policy = policy_feat; a.alpha = alpha; a.eidx = expert_idx; a.pidx = policy_idx;
    a.step_size = step_size; a.bc2_sqrt = bc2_sqrt; a.trace = trace;
    a.gpart = (float*)(ws + w.gpart); a.grad = (float*)(ws + w.grad); a.losspart = (float*)(ws + w.losspart);
    a.bar = (unsigned int*)(ws + w.bar); a.prof = (long long*)(ws + w.prof);
    a.first_adam_step = cfg->first_adam_step;
    a.dp_on = cfg->dp_ctx != nullptr;
    if (a.dp_on) {
        a.dp = dp_view(cfg->dp_ctx);
        SG_REQUIRE(a.dp.cap >= a.P + 4 && grid < kDpMaxSlices, "sg_disc_update: dp context too small (%d floats)", a.dp.cap);
    } else {
        memset(&a.dp, 0, sizeof(a.dp));
    }
    const size_t smem_tile = disc_tile_smem_floats(cfg) * sizeof(float);
    const size_t smem_res = disc_resident_smem_bytes(cfg);
    int mode = cfg->mode;
    if (mode == 0) mode = smem_res <= kDiscMaxDynSmem ? 3 : 2;
    SG_CUDA(cudaMemsetAsync(ws, 0, w.total, s));
    if (mode == 3 || mode == 2) {
        const void* fn = (const void*)disc_persistent_kernel<false>;
        int threads = kStepThreads;
        if (mode == 3) {
            fn = (const void*)disc_persistent_kernel<true>;
            if (disc_reg_ok(cfg)) {
                const bool one = disc_tb(cfg) == 1, multi = a.ntiles > grid;
#define SG_DISC_REG_FN(HQ_) (one ? (const void*)disc_reg_kernel<HQ_, 1, false>                                      \
                                 : (multi ? (const void*)disc_reg_kernel<HQ_, 2, true> : (const void*)disc_reg_kernel<HQ_, 2, false>))
                switch (cfg->hidden) {
                    case 48: fn = SG_DISC_REG_FN(12); break;
                    case 64: fn = SG_DISC_REG_FN(16); break;
                    case 100: fn = SG_DISC_REG_FN(25); break;
                    default: fn = SG_DISC_REG_FN(32); break;
                }
#undef SG_DISC_REG_FN
            }
        }
        const size_t smem = mode == 3 ? smem_res : smem_tile;
        SG_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        SG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
        SG_REQUIRE(per_sm >= 1 && grid <= per_sm * sms, "sg_disc_update: cooperative grid of %d CTAs does not fit", grid);
        void* kargs[] = {(void*)&a};
        SG_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(threads), kargs, smem, s));
        count_launches(1);
    } else {
        SG_CUDA(cudaFuncSetAttribute(disc_phaseA_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tile));
        for (int step = 0; step < a.nsteps; ++step) {
            disc_phaseA_kernel<<<grid, kStepThreads, smem_tile, s>>>(a, step);
            disc_phaseB_kernel<<<grid, kStepThreads, 0, s>>>(a);
            if (allreduce_cb) {
                int cb = allreduce_cb(a.grad, a.P + 3, allreduce_user);
                SG_REQUIRE(cb == 0, "sg_disc_update: allreduce callback failed with %d at step %d", cb, step);
            }
            disc_phaseC_kernel<<<grid, kStepThreads, 0, s>>>(a, step);
            count_launches(3);
        }
        SG_CUDA(cudaGetLastError());
    }
    return SG_OK;
}

int sg_disc_predict_reward(const float* params, int feat_dim, int hidden, const float* d_in, int n_rows, double gamma,
                           const float* masks, double offset, int has_returns, float* reward, float* returns,
                           void* stream) {
    SG_REQUIRE(params && d_in && reward && n_rows > 0 && feat_dim > 0 && hidden > 0, "sg_disc_predict_reward: bad arguments");
    SG_REQUIRE(!returns || !has_returns || masks, "sg_disc_predict_reward: masks required to update returns");
    return launch_reward(params, feat_dim, hidden, d_in, n_rows, (float)offset, reward, returns, masks, (float)gamma,
                         has_returns, (cudaStream_t)stream);
}

int64_t sg_relabel_workspace_bytes(int T, int N) {
    if (T <= 0 || N <= 0) return -1;
    return (int64_t)relabel_ws(T, N).total;
}

static int relabel_from_raw(const float* raw, const float* masks, float* rewards, int T, int N, double gamma,
                            float* disc_returns, int has_returns, double* rms_state, float* mean_returns, char* ws,
                            const RelabelWs& w, cudaStream_t s) {
    float* ret_all = (float*)(ws + w.ret_all);
    float* bmean = (float*)(ws + w.bmean);
    float* bvar = (float*)(ws + w.bvar);
    double* scale = (double*)(ws + w.scale);
    static SmemGrant scan_grant;
    if (int rc = grant_smem(scan_grant, relabel_scan_kernel, kRlScanSmemBytes)) return rc;
    relabel_scan_kernel<<<(N + 31) / 32, 256, kRlScanSmemBytes, s>>>(raw, masks, ret_all, disc_returns, T, N, (float)gamma, has_returns);
    relabel_moments_kernel<<<(T + 127) / 128, 128, 0, s>>>(ret_all, T, N, bmean, bvar, mean_returns);
    relabel_rms_kernel<<<1, 256, 0, s>>>(bmean, bvar, T, N, rms_state, scale);
    long long total = (long long)T * N;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    relabel_apply_kernel<<<blocks, 256, 0, s>>>(raw, scale, rewards, T, N);
    count_launches(4);
    SG_CUDA(cudaGetLastError());
    return SG_OK;
}

int sg_relabel_normalize(const float* raw_reward, const float* masks, float* rewards, int T, int N, double gamma,
                         float* disc_returns, int has_returns, double* rms_state, float* mean_returns, void* workspace,
                         void* stream) {
    SG_REQUIRE(raw_reward && masks && rewards && disc_returns && rms_state && mean_returns && workspace,
               "sg_relabel_normalize: null pointer");
    SG_REQUIRE(T > 0 && N > 0, "sg_relabel_normalize: non-positive sizes");
    const RelabelWs w = relabel_ws(T, N);
    return relabel_from_raw(raw_reward, masks, rewards, T, N, gamma, disc_returns, has_returns, rms_state, mean_returns,
                            (char*)workspace, w, (cudaStream_t)stream);
}

int sg_selftest_division(uint64_t seed, int blocks, int per_thread, double b_lo, double b_hi, uint64_t* mismatches,
                         void* stream) {
    SG_REQUIRE(mismatches && blocks > 0 && per_thread > 0 && b_lo > 0 && b_hi > b_lo, "sg_selftest_division: bad arguments");
    div_rcp_selfcheck_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((unsigned long long)seed, per_thread, b_lo, b_hi,
                                                                      (unsigned long long*)mismatches);
    count_launches(1);
    SG_CUDA(cudaGetLastError());
    return SG_OK;
}

int sg_disc_relabel(const float* params, int feat_dim, int hidden, const float* obs_feat, const float* masks,
                    float* rewards, int T, int N, double gamma, double offset, float* disc_returns, int has_returns,
                    double* rms_state, float* mean_returns, void* workspace, void* stream) {
    SG_REQUIRE(params && obs_feat && masks && rewards && disc_returns && rms_state && mean_returns && workspace,
               "sg_disc_relabel: null pointer");
    SG_REQUIRE(T > 0 && N > 0 && feat_dim > 0 && hidden > 0, "sg_disc_relabel: non-positive sizes");
    cudaStream_t s = (cudaStream_t)stream;
    const RelabelWs w = relabel_ws(T, N);
    char* ws = (char*)workspace;
    float* raw = (float*)(ws + w.raw);
    // reward for step t reads obs_feat[t+1] (main_gail_dyn_ppo.py:278): rows N.. of the (T+1,N,F) buffer
    int rc = launch_reward(params, feat_dim, hidden, obs_feat + (size_t)N * feat_dim, T * N, (float)offset, raw, nullptr,
                           nullptr, 0.f, 0, s);
    if (rc) return rc;
    return relabel_from_raw(raw, masks, rewards, T, N, gamma, disc_returns, has_returns, rms_state, mean_returns, ws, w, s);
}

#pragma GCC visibility pop
}  // extern "C"
