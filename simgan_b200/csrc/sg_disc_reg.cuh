// Register-resident discriminator tile (hidden width H = 4*HQ <= 128, known at compile time).
//
// The 2nd trunk layer W2 (H x H) is walked four times per optimizer step -- forward, dZ1/v1, ub2 and hb1
// of the gradient-penalty double backward (A2C/algo/gail.py:67-89, 165-188) -- always against a handful of
// rows, so the cost of those passes is the cost of fetching W2.  Here every thread keeps its slice of W2 in
// REGISTERS for the whole step, once in row form and once in column form:
//     thread t:  unit u = t >> 1, half kh = t & 1, quads [kh*Q0, kh*Q0+Q0) of the contraction (Q0 = ceil(HQ/2))
//       wrow[i] = W2[u][k]   (k in my quads)   -> "F" passes  out[r][u] = sum_k x[r][k] W2[u][k]
//       wcol[i] = W2[m][u]   (m in my quads)   -> "B" passes  out[r][u] = sum_m y[r][m] W2[m][u]
// A pass is then RT broadcast LDS.128 of activations per quad against 4*RT FMAs, one xor-shuffle joins the two
// halves, and no weight traffic at all.  W1 (H x F), w3 and the biases are read from a small natural image
// in shared memory.  Same math, same slot bookkeeping as disc_tile (sg_disc.cu); summation order differs.
#pragma once
#include "sg_common.cuh"
#include "sg_colgemm.cuh"

namespace sg {

struct DiscRegSmem {
    float *X, *H1, *H2, *DD, *LOSS, *Y2, *L2t, *L1t, *U1, *Z2, *C3, *PB;
    int ldf, ldh;
    __host__ __device__ static int pb_floats(int F) { return 2 * round_up(F, 4) * 8; }     // pass-B partials
    __host__ __device__ static int floats(int F, int H) {
        const int ldf = round_up(F, 4), ldh = round_up(H, 4);
        return 8 * ldf + 8 * ldh + 6 * ldh + 8 + 24 + 6 * ldh + 8 * ldh + 8 * ldh + 2 * ldh + 2 * ldh + 2 * ldh + pb_floats(F);
    }
    __device__ void carve(float* sm, int F, int H) {
        ldf = round_up(F, 4); ldh = round_up(H, 4);
        X = sm; sm += 8 * ldf;        // rows 0..5 inputs (e0 e1 p0 p1 m0 m1), rows 6,7 gbar
        H1 = sm; sm += 8 * ldh;       // rows 0..5 h1, rows 6,7 vb1
        H2 = sm; sm += 6 * ldh;
        DD = sm; sm += 8;
        LOSS = sm; sm += 24;
        Y2 = sm; sm += 6 * ldh;       // rows 0..3 dz2 (e,p), rows 4,5 u2      (operand of pass A)
        L2t = sm; sm += 8 * ldh;      // [H][8]: slots 0..3 dz2, 4,5 zb2, 6,7 u2
        L1t = sm; sm += 8 * ldh;      // [H][8]: slots 0..3 dz1, 4,5 zb1, 6,7 u1
        U1 = sm; sm += 2 * ldh;       // [2][ldh]  operand of pass B
        Z2 = sm; sm += 2 * ldh;       // [2][ldh]  operand of pass E
        C3 = sm; sm += 2 * ldh;       // dw3 terms of the penalty
        PB = sm;
    }
};

__device__ __forceinline__ float dsoftplusf(float z) { return fmaxf(z, 0.f) + log1pf(expf(-fabsf(z))); }
__device__ __forceinline__ float dsigmoidf(float z) { return __fdiv_rn(1.f, 1.f + expf(-z)); }

// flat image of everything except W2: [W1 | b1 | b2 | w3 | b3] with the offsets of DiscLayout minus the W2 block
struct DiscRegImage {
    int w1, b1, b2, w3, b3, total;
};
__host__ __device__ inline DiscRegImage make_disc_reg_image(int F, int H) {
    DiscRegImage I;
    int o = 0;
    I.w1 = o; o += round_up(H * F, 4);
    I.b1 = o; o += round_up(H, 4);
    I.b2 = o; o += round_up(H, 4);
    I.w3 = o; o += round_up(H, 4);
    I.b3 = o; o += 4;
    I.total = o;
    return I;
}

template <int HQ>
struct DiscRegW2 {
    static constexpr int Q0 = (HQ + 1) / 2;
    float wrow[Q0 * 4];
    float wcol[Q0 * 4];
};

// (re)load this thread's register slices of W2 and the small shared image.  W2 travels global -> shared as ONE TMA
// bulk copy (cp.async.bulk) into `stage` (H*H floats, natural layout) and is then picked into registers from
// shared memory: the row form needs 16-byte pieces of 128 different rows and the column form scalar columns, which
// as direct global loads cost ~20 memory transactions per warp-load.
template <int HQ>
__device__ __forceinline__ void disc_reg_fill(DiscRegW2<HQ>& w, float* __restrict__ img, float* __restrict__ stage,
                                              const float* __restrict__ params, const DiscLayout& L, const DiscRegImage& I,
                                              int tid, unsigned long long* bar, unsigned int parity) {
    constexpr int H = 4 * HQ, Q0 = DiscRegW2<HQ>::Q0;
    const int u = tid >> 1, kh = tid & 1;
    const float* W2 = params + L.w2;
    const bool live = u < H;
    // three TMA bulk copies (W2 -> stage; [W1|b1] and [b2|w3|b3] -> img) on one mbarrier; nothing else ever writes
    // these shared-memory regions, so there is no generic/async proxy hazard on them
    const int n1 = L.w2;                       // floats before the W2 block (multiple of 4)
    const int n2 = L.total - L.b2;             // floats after it
    if (tid == 0) {
        fence_proxy_async();                   // the parameters were written by other CTAs' generic stores (grid barrier acquired)
        mbar_expect_tx(bar, (unsigned int)((H * H + n1 + n2) * sizeof(float)));
        tma_bulk_g2s(stage, W2, (unsigned int)(H * H * sizeof(float)), bar);
        tma_bulk_g2s(img, params, (unsigned int)(n1 * sizeof(float)), bar);
        tma_bulk_g2s(img + n1, params + L.b2, (unsigned int)(n2 * sizeof(float)), bar);
    }
    mbar_wait(bar, parity);
#pragma unroll
    for (int i = 0; i < Q0; ++i) {
        const int q = kh * Q0 + i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live && q < HQ) v = *reinterpret_cast<const float4*>(stage + (size_t)u * H + 4 * q);
        w.wrow[4 * i] = v.x; w.wrow[4 * i + 1] = v.y; w.wrow[4 * i + 2] = v.z; w.wrow[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < Q0 * 4; ++i) {
        const int m = kh * Q0 * 4 + i;
        w.wcol[i] = (live && m < H) ? stage[(size_t)m * H + u] : 0.f;
    }
    __syncthreads();                           // `stage` may alias buffers the tile phase writes
}

// out[r] = sum over my quads of A[r][k] * wreg[k]; both lanes of a unit get the full sum
template <int HQ, int NR>
__device__ __forceinline__ void reg_pass(const float (&wreg)[DiscRegW2<HQ>::Q0 * 4], const float* __restrict__ A, int ld,
                                         const int (&rows)[NR], int kh, float (&out)[NR]) {
    constexpr int Q0 = DiscRegW2<HQ>::Q0;
#pragma unroll
    for (int j = 0; j < NR; ++j) out[j] = 0.f;
#pragma unroll
    for (int i = 0; i < Q0; ++i) {
        int q = kh * Q0 + i;
        q = q < HQ ? q : HQ - 1;               // phantom quad of the upper half: weights are zero, address stays valid
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            const float4 x = *reinterpret_cast<const float4*>(A + rows[j] * ld + 4 * q);
            out[j] = fmaf(x.x, wreg[4 * i], out[j]); out[j] = fmaf(x.y, wreg[4 * i + 1], out[j]);
            out[j] = fmaf(x.z, wreg[4 * i + 2], out[j]); out[j] = fmaf(x.w, wreg[4 * i + 3], out[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < NR; ++j) out[j] += __shfl_xor_sync(0xffffffffu, out[j], 1);
}

// Tile inputs of the NEXT optimizer step (see PpoPrefetch in sg_ppo.cu): expert / policy feature elements and the mixup
// alpha of the row my gather element belongs to, fetched while this step's reduce / Adam phase waits on its barriers.
struct DiscPrefetch {
    bool on, valid;
    int ei, pi;           // expert / policy row indices (stage 1), -1 = padding element
    float xe, xp, al;
};
template <int TB, class Args>
__device__ __forceinline__ void disc_prefetch_idx(const Args& a, DiscPrefetch& pf, int step, int tile, int ldf) {
    const int tid = threadIdx.x;
    pf.ei = pf.pi = -1;
    pf.al = 0.f;
    if (tid < 4 * TB * ldf) {
        const int r = tid / ldf;
        const int kind = r / TB, j = r - kind * TB;
        const int row = a.row_begin + tile * TB + j;
        if (kind < 3 && row < a.row_end) {
            pf.ei = a.eidx[(size_t)step * a.B + row];
            pf.pi = a.pidx[(size_t)step * a.B + row];
            pf.al = a.alpha[(size_t)step * a.B + row];
        }
    }
}
template <class Args>
__device__ __forceinline__ void disc_prefetch_rows(const Args& a, DiscPrefetch& pf, int ldf) {
    const int k = threadIdx.x % ldf;
    const bool ok = pf.ei >= 0 && k < a.F;
    pf.xe = ok ? a.expert[(size_t)pf.ei * a.F + k] : 0.f;
    pf.xp = ok ? a.policy[(size_t)pf.pi * a.F + k] : 0.f;
    pf.valid = true;
}

// One tile = 2 (expert, policy, mixup) row triples.  `img` = small natural image (DiscRegImage), `w` = W2 slices.
// MULTI: the CTA runs several tiles per step and accumulates its partial gradient across them; single-tile kernels
// compile without any accumulate code.
template <int HQ, bool MULTI, class Args>
__device__ void disc_tile_reg(const Args& a, const DiscRegW2<HQ>& w, const float* __restrict__ img, const DiscRegImage& I,
                              int step, int tile, float* __restrict__ gout, float* __restrict__ lossout,
                              DiscRegSmem& sm, bool acc_in, DiscPrefetch& pf) {
    const bool acc = MULTI && acc_in;
    constexpr int H = 4 * HQ, TB = 2, R = 8;
    const int tid = threadIdx.x, nth = kStepThreads;
    const int F = a.F, ldf = sm.ldf, ldh = sm.ldh;
    const int u = tid >> 1, kh = tid & 1;
    const bool live = u < H;
    const int row0 = a.row_begin + tile * TB;
    const int32_t* eidx = a.eidx + (size_t)step * a.B;
    const int32_t* pidx = a.pidx + (size_t)step * a.B;
    const float* alpha = a.alpha + (size_t)step * a.B;
    const float invB = 1.f / (float)a.B;
    const float* W1 = img + I.w1; const float* B1 = img + I.b1; const float* B2 = img + I.b2;
    const float* W3 = img + I.w3; const float* B3 = img + I.b3;

    // ---- S0: rows [e0 e1 | p0 p1 | m0 m1 | 0 0], mixup = alpha*e + (1-alpha)*p (gail.py:72-75) -----------------
    if (pf.valid) {
        if (tid < R * ldf) {                       // fetched during the previous step's barriers
            const int kind = (tid / ldf) / TB;
            const bool ok = pf.ei >= 0 && (tid % ldf) < F;
            const float mixed = __fadd_rn(__fmul_rn(pf.al, pf.xe), __fmul_rn(__fsub_rn(1.f, pf.al), pf.xp));
            sm.X[tid] = !ok ? 0.f : (kind == 0 ? pf.xe : (kind == 1 ? pf.xp : mixed));
        }
    } else {
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
    }
    __syncthreads();
    pf.valid = false;
    const bool pf_next = pf.on && step + 1 < a.nsteps;
    if (pf_next) disc_prefetch_idx<TB>(a, pf, step + 1, tile, ldf);      // stage 1: next step's row indices and alpha
    // rows owned by this lane in 6-row stages: kh, kh+2, kh+4  (e_kh, p_kh, m_kh)
    const int myrows[3] = {kh, kh + 2, kh + 4};
    // ---- S1: layer 1 forward, W1 natural in shared memory -------------------------------------------------------
    if (live) {
        float s[3] = {0.f, 0.f, 0.f};
        const float* wr = W1 + (size_t)u * F;
        for (int k0 = 0; k0 < F; k0 += 4) {
            float wv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) wv[j] = (k0 + j < F) ? wr[k0 + j] : 0.f;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float4 x = *reinterpret_cast<const float4*>(sm.X + myrows[i] * ldf + k0);
                s[i] = fmaf(x.x, wv[0], s[i]); s[i] = fmaf(x.y, wv[1], s[i]); s[i] = fmaf(x.z, wv[2], s[i]); s[i] = fmaf(x.w, wv[3], s[i]);
            }
        }
        const float b = B1[u];
#pragma unroll
        for (int i = 0; i < 3; ++i) sm.H1[myrows[i] * ldh + u] = tanhf(s[i] + b);
    }
    __syncthreads();
    // ---- S2: layer 2 forward ("F" pass on 6 rows) -----------------------------------------------------------------
    {
        const int rows6[6] = {0, 1, 2, 3, 4, 5};
        float o[6];
        reg_pass<HQ, 6>(w.wrow, sm.H1, ldh, rows6, kh, o);
        if (live) {
            const float b = B2[u];
#pragma unroll
            for (int i = 0; i < 3; ++i) sm.H2[myrows[i] * ldh + u] = tanhf((kh ? o[2 * i + 1] : o[2 * i]) + b);
        }
    }
    __syncthreads();
    // ---- S3: logits + BCE-with-logits seeds (gail.py:171-176): warp r handles row r ------------------------------------
    {
        const int wid = tid >> 5, lane = tid & 31;
        if (wid < 6) {
            float d = 0.f;
            for (int n = lane; n < H; n += 32) d = fmaf(W3[n], sm.H2[wid * ldh + n], d);
            d = warp_sum(d) + B3[0];
            if (lane == 0) {
                const int kind = wid / TB, j = wid - kind * TB;
                const bool ok = (row0 + j) < a.row_end;
                float dd = 0.f, le = 0.f, lp = 0.f;
                if (ok && kind == 0) { dd = (dsigmoidf(d) - 1.f) * invB; le = dsoftplusf(-d); }
                if (ok && kind == 1) { dd = dsigmoidf(d) * invB; lp = dsoftplusf(d); }
                sm.DD[wid] = dd; sm.LOSS[wid] = le; sm.LOSS[8 + wid] = lp;
            }
        }
    }
    __syncthreads();
    // ---- S4: dz2 (e,p) and u2 (mixup) ---------------------------------------------------------------------------------
    if (live) {
        const float w3 = W3[u];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int r = myrows[i];
            const float h2 = sm.H2[r * ldh + u];
            const float base = w3 * (1.f - h2 * h2);
            if (i < 2) {
                const float y = sm.DD[r] * base;
                sm.Y2[r * ldh + u] = y; sm.L2t[u * R + r] = y;
            } else {
                sm.Y2[r * ldh + u] = base; sm.L2t[u * R + 6 + kh] = base;       // u2 pairs with vb1 in slot 6+j
            }
        }
    }
    __syncthreads();
    // ---- S5: pass A ("B" pass on 6 rows): dz1 for e,p rows; v1, u1 for the mixup rows ------------------------------------
    float v1 = 0.f, h1m = 0.f;
    {
        const int rows6[6] = {0, 1, 2, 3, 4, 5};
        float o[6];
        reg_pass<HQ, 6>(w.wcol, sm.Y2, ldh, rows6, kh, o);
        if (live) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int r = myrows[i];
                const float h1 = sm.H1[r * ldh + u];
                sm.L1t[u * R + r] = (kh ? o[2 * i + 1] : o[2 * i]) * (1.f - h1 * h1);
            }
            v1 = kh ? o[5] : o[4];
            h1m = sm.H1[(4 + kh) * ldh + u];
            const float u1 = v1 * (1.f - h1m * h1m);
            sm.U1[kh * ldh + u] = u1;
            sm.L1t[u * R + 6 + kh] = u1;                                           // u1 pairs with gbar in slot 6+j
        }
    }
    __syncthreads();
    // ---- S6: pass B: g[j][f] = sum_n u1[j][n] W1[n][f]  (input gradient of the mixup rows) ---------------------------------
    {
        const int nout = 2 * ldf;                          // (j, f) outputs, f padded to ldf
        int nch = nth / nout; nch = nch < 1 ? 1 : (nch > 8 ? 8 : nch);
        const int clen = (H + nch - 1) / nch;
        for (int e = tid; e < nout * nch; e += nth) {
            const int c = e / nout, o = e - c * nout;
            const int j = o / ldf, f = o - j * ldf;
            float s = 0.f;
            if (f < F) {
                const int n1 = min(H, (c + 1) * clen);
                for (int n = c * clen; n < n1; ++n) s = fmaf(sm.U1[j * ldh + n], W1[(size_t)n * F + f], s);
            }
            sm.PB[c * nout + o] = s;
        }
        __syncthreads();
        if (tid < 32 * TB) {
            const int j = tid >> 5, lane = tid & 31;
            float ssq = 0.f;
            for (int f = lane; f < ldf; f += 32) {
                float g = 0.f;
                for (int c = 0; c < nch; ++c) g += sm.PB[c * nout + j * ldf + f];
                sm.X[(6 + j) * ldf + f] = g;               // raw g for now (zero in the padding lanes)
                ssq += g * g;
            }
            ssq = warp_sum(ssq);
            const float nrm = sqrtf(ssq);
            const bool ok = (row0 + j) < a.row_end;
            // penalty lambda*mean((|g|-1)^2) -> gbar = (2 lambda / B)(n-1) g / n
            const float coef = (ok && nrm > 0.f) ? (2.f * a.gp_lambda * invB) * (nrm - 1.f) / nrm : 0.f;
            for (int f = lane; f < ldf; f += 32) sm.X[(6 + j) * ldf + f] *= coef;
            if (lane == 0) sm.LOSS[16 + j] = ok ? (nrm - 1.f) * (nrm - 1.f) : 0.f;
        }
    }
    __syncthreads();
    // ---- S7: pass C: ub1 = gbar . W1^T -> vb1, hb1 (lane kh owns mixup row j = kh) --------------------------------------------
    float hb1 = 0.f;
    if (live) {
        float s = 0.f;
        const float* wr = W1 + (size_t)u * F;
        const float* gb = sm.X + (6 + kh) * ldf;
        for (int k0 = 0; k0 < F; k0 += 4) {
            const float4 x = *reinterpret_cast<const float4*>(gb + k0);
            s = fmaf(x.x, wr[k0], s);
            if (k0 + 1 < F) s = fmaf(x.y, wr[k0 + 1], s);
            if (k0 + 2 < F) s = fmaf(x.z, wr[k0 + 2], s);
            if (k0 + 3 < F) s = fmaf(x.w, wr[k0 + 3], s);
        }
        sm.H1[(6 + kh) * ldh + u] = s * (1.f - h1m * h1m);                          // vb1
        hb1 = -2.f * s * v1 * h1m;
    }
    __syncthreads();
    // ---- S8: pass D ("F" pass on the 2 vb1 rows): ub2 -> dw3 term, zb2 --------------------------------------------------------
    {
        const int rows2[2] = {6, 7};
        float o[2];
        reg_pass<HQ, 2>(w.wrow, sm.H1, ldh, rows2, kh, o);
        if (live) {
            const float ub2 = kh ? o[1] : o[0];
            const float h2 = sm.H2[(4 + kh) * ldh + u];
            const float om = 1.f - h2 * h2;
            sm.C3[kh * ldh + u] = ub2 * om;
            const float zb2 = -2.f * ub2 * W3[u] * h2 * om;
            sm.Z2[kh * ldh + u] = zb2;
            sm.L2t[u * R + 4 + kh] = zb2;
        }
    }
    __syncthreads();
    // ---- S9: pass E ("B" pass on the 2 zb2 rows): hb1 += zb2 . W2 ; zb1 = hb1*(1-h1^2) -------------------------------------------
    {
        const int rows2[2] = {0, 1};
        float o[2];
        reg_pass<HQ, 2>(w.wcol, sm.Z2, ldh, rows2, kh, o);
        if (live) sm.L1t[u * R + 4 + kh] = (hb1 + (kh ? o[1] : o[0])) * (1.f - h1m * h1m);
    }
    __syncthreads();
    // ---- S10: parameter gradients of this tile ---------------------------------------------------------------------------------------
    outer_cols<R>(gout + a.L.w2, sm.L2t, sm.H1, ldh, H, H, tid, nth, acc);
    rowsum_store<R>(gout + a.L.b2, sm.L2t, H, tid, nth, acc, 3 * TB);
    if ((F & 3) == 0) outer_cols<R>(gout + a.L.w1, sm.L1t, sm.X, ldf, H, F, tid, nth, acc);
    else outer_store<R, 1>(gout + a.L.w1, sm.L1t, sm.X, ldf, H, F, tid, nth, acc);
    rowsum_store<R>(gout + a.L.b1, sm.L1t, H, tid, nth, acc, 3 * TB);
    for (int n = tid; n < H; n += nth) {
        float s = 0.f;
        for (int r = 0; r < 2 * TB; ++r) s += sm.DD[r] * sm.H2[r * ldh + n];
        for (int j = 0; j < TB; ++j) s += sm.C3[j * ldh + n];
        __stcg(gout + a.L.w3 + n, acc ? s + __ldcg(gout + a.L.w3 + n) : s);
    }
    if (tid == nth - 1) {
        float db3 = 0.f, le = 0.f, lp = 0.f, gp = 0.f;
        for (int r = 0; r < 2 * TB; ++r) db3 += sm.DD[r];
        for (int r = 0; r < TB; ++r) { le += sm.LOSS[r]; lp += sm.LOSS[8 + TB + r]; gp += sm.LOSS[16 + r]; }
        if (acc) { db3 += __ldcg(gout + a.L.b3); le += lossout[0]; lp += lossout[1]; gp += lossout[2]; }
        __stcg(gout + a.L.b3, db3);
        lossout[0] = le; lossout[1] = lp; lossout[2] = gp;
    }
    if (pf_next) disc_prefetch_rows(a, pf, ldf);                        // stage 2: the feature rows themselves
    __syncthreads();
}

// One tile = ONE (expert, policy, mixup) row triple: used when the minibatch (or this rank's shard of it) has at most
// one triple per SM, so that twice as many CTAs each do half the work of disc_tile_reg.  Row slots (R = 4):
//   X  [e | p | m | gbar]     H1 [h1_e | h1_p | h1_m | vb1]     H2 [h2_e | h2_p | h2_m]     Y2 [dz2_e | dz2_p | u2]
//   L2t[u][4] = {dz2_e, dz2_p, zb2, u2}  pairs with the H1 rows     L1t[u][4] = {dz1_e, dz1_p, zb1, u1}  pairs with the X rows
// Both lanes of a unit hold every pass result (the xor-shuffle in reg_pass); lane 0 finishes the expert and the mixup
// row, lane 1 the policy row.
template <int HQ, class Args>
__device__ void disc_tile_reg1(const Args& a, const DiscRegW2<HQ>& w, const float* __restrict__ img, const DiscRegImage& I,
                               int step, int tile, float* __restrict__ gout, float* __restrict__ lossout,
                               DiscRegSmem& sm, bool /*acc: never -- this tile is only used with one tile per CTA*/, DiscPrefetch& pf) {
    constexpr bool acc = false;           // compile-time: no read-modify-write code (and no speculative L2 loads) in the stores
    constexpr int H = 4 * HQ, R = 4;
    const int tid = threadIdx.x, nth = kStepThreads;
    const int F = a.F, ldf = sm.ldf, ldh = sm.ldh;
    const int u = tid >> 1, kh = tid & 1;
    const bool live = u < H;
    const int row = a.row_begin + tile;
    const bool rowok = row < a.row_end;
    const float invB = 1.f / (float)a.B;
    const float* W1 = img + I.w1; const float* B1 = img + I.b1; const float* B2 = img + I.b2;
    const float* W3 = img + I.w3; const float* B3 = img + I.b3;

    // ---- S0: rows [e | p | m | 0], mixup = alpha*e + (1-alpha)*p (gail.py:72-75) ---------------------------------------
    if (pf.valid) {
        if (tid < R * ldf) {                       // fetched during the previous step's barriers
            const int kind = tid / ldf;
            const bool ok = pf.ei >= 0 && (tid % ldf) < F;
            const float mixed = __fadd_rn(__fmul_rn(pf.al, pf.xe), __fmul_rn(__fsub_rn(1.f, pf.al), pf.xp));
            sm.X[tid] = !ok ? 0.f : (kind == 0 ? pf.xe : (kind == 1 ? pf.xp : mixed));
        }
    } else {
        const int ei = rowok ? a.eidx[(size_t)step * a.B + row] : 0, pi = rowok ? a.pidx[(size_t)step * a.B + row] : 0;
        const float al = rowok ? a.alpha[(size_t)step * a.B + row] : 0.f;
        for (int e = tid; e < R * ldf; e += nth) {
            const int kind = e / ldf, k = e - kind * ldf;
            float x = 0.f;
            if (kind < 3 && rowok && k < F) {
                const float xe = a.expert[(size_t)ei * F + k];
                const float xp = a.policy[(size_t)pi * F + k];
                x = kind == 0 ? xe : (kind == 1 ? xp : __fadd_rn(__fmul_rn(al, xe), __fmul_rn(__fsub_rn(1.f, al), xp)));
            }
            sm.X[e] = x;
        }
    }
    __syncthreads();
    pf.valid = false;
    const bool pf_next = pf.on && step + 1 < a.nsteps;
    if (pf_next) disc_prefetch_idx<1>(a, pf, step + 1, tile, ldf);       // stage 1: next step's row indices and alpha
    const int rows3[3] = {0, 1, 2};
    // ---- S1: layer 1 forward, W1 natural in shared memory: lane 0 -> rows e, m; lane 1 -> row p ---------------------------
    if (live) {
        const int ra = kh;                        // 0: expert, 1: policy
        float s0 = 0.f, s1 = 0.f;
        const float* wr = W1 + (size_t)u * F;
        for (int k0 = 0; k0 < F; k0 += 4) {
            float wv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) wv[j] = (k0 + j < F) ? wr[k0 + j] : 0.f;
            const float4 x = *reinterpret_cast<const float4*>(sm.X + ra * ldf + k0);
            s0 = fmaf(x.x, wv[0], s0); s0 = fmaf(x.y, wv[1], s0); s0 = fmaf(x.z, wv[2], s0); s0 = fmaf(x.w, wv[3], s0);
            if (!kh) {
                const float4 y = *reinterpret_cast<const float4*>(sm.X + 2 * ldf + k0);
                s1 = fmaf(y.x, wv[0], s1); s1 = fmaf(y.y, wv[1], s1); s1 = fmaf(y.z, wv[2], s1); s1 = fmaf(y.w, wv[3], s1);
            }
        }
        const float b = B1[u];
        sm.H1[ra * ldh + u] = tanhf(s0 + b);
        if (!kh) sm.H1[2 * ldh + u] = tanhf(s1 + b);
    }
    __syncthreads();
    // ---- S2: layer 2 forward ("F" pass on 3 rows) ---------------------------------------------------------------------------
    {
        float o[3];
        reg_pass<HQ, 3>(w.wrow, sm.H1, ldh, rows3, kh, o);
        if (live) {
            const float b = B2[u];
            sm.H2[kh * ldh + u] = tanhf((kh ? o[1] : o[0]) + b);
            if (!kh) sm.H2[2 * ldh + u] = tanhf(o[2] + b);
        }
    }
    __syncthreads();
    // ---- S3: logits + BCE-with-logits seeds (gail.py:171-176): warp r handles row r ----------------------------------------
    {
        const int wid = tid >> 5, lane = tid & 31;
        if (wid < 2) {
            float d = 0.f;
            for (int n = lane; n < H; n += 32) d = fmaf(W3[n], sm.H2[wid * ldh + n], d);
            d = warp_sum(d) + B3[0];
            if (lane == 0) {
                float dd = 0.f, l = 0.f;
                if (rowok && wid == 0) { dd = (dsigmoidf(d) - 1.f) * invB; l = dsoftplusf(-d); }
                if (rowok && wid == 1) { dd = dsigmoidf(d) * invB; l = dsoftplusf(d); }
                sm.DD[wid] = dd; sm.LOSS[wid] = l;           // LOSS[0] expert term, LOSS[1] policy term
            }
        }
    }
    __syncthreads();
    // ---- S4: dz2 (e,p) and u2 (mixup) -------------------------------------------------------------------------------------------
    if (live) {
        const float w3 = W3[u];
        {
            const float h2 = sm.H2[kh * ldh + u];
            const float y = sm.DD[kh] * (w3 * (1.f - h2 * h2));
            sm.Y2[kh * ldh + u] = y; sm.L2t[u * R + kh] = y;
        }
        if (!kh) {
            const float h2 = sm.H2[2 * ldh + u];
            const float base = w3 * (1.f - h2 * h2);
            sm.Y2[2 * ldh + u] = base; sm.L2t[u * R + 3] = base;          // u2 pairs with vb1 in slot 3
        }
    }
    __syncthreads();
    // ---- S5: pass A ("B" pass on 3 rows): dz1 for the e,p rows; v1, u1 for the mixup row -------------------------------------
    float v1 = 0.f, h1m = 0.f;
    {
        float o[3];
        reg_pass<HQ, 3>(w.wcol, sm.Y2, ldh, rows3, kh, o);
        if (live) {
            const float h1 = sm.H1[kh * ldh + u];
            sm.L1t[u * R + kh] = (kh ? o[1] : o[0]) * (1.f - h1 * h1);
            v1 = o[2];
            h1m = sm.H1[2 * ldh + u];
            if (!kh) {
                const float u1 = v1 * (1.f - h1m * h1m);
                sm.U1[u] = u1;
                sm.L1t[u * R + 3] = u1;                                    // u1 pairs with gbar in slot 3
            }
        }
    }
    __syncthreads();
    // ---- S6: pass B: g[f] = sum_n u1[n] W1[n][f]  (input gradient of the mixup row) -------------------------------------------
    {
        const int nout = ldf;
        int nch = nth / nout; nch = nch < 1 ? 1 : (nch > 8 ? 8 : nch);
        const int clen = (H + nch - 1) / nch;
        for (int e = tid; e < nout * nch; e += nth) {
            const int c = e / nout, f = e - c * nout;
            float s = 0.f;
            if (f < F) {
                const int n1 = min(H, (c + 1) * clen);
                for (int n = c * clen; n < n1; ++n) s = fmaf(sm.U1[n], W1[(size_t)n * F + f], s);
            }
            sm.PB[c * nout + f] = s;
        }
        __syncthreads();
        if (tid < 32) {
            const int lane = tid;
            float ssq = 0.f;
            for (int f = lane; f < ldf; f += 32) {
                float g = 0.f;
                for (int c = 0; c < nch; ++c) g += sm.PB[c * nout + f];
                sm.X[3 * ldf + f] = g;                     // raw g for now (zero in the padding lanes)
                ssq += g * g;
            }
            ssq = warp_sum(ssq);
            const float nrm = sqrtf(ssq);
            // penalty lambda*mean((|g|-1)^2) -> gbar = (2 lambda / B)(n-1) g / n
            const float coef = (rowok && nrm > 0.f) ? (2.f * a.gp_lambda * invB) * (nrm - 1.f) / nrm : 0.f;
            for (int f = lane; f < ldf; f += 32) sm.X[3 * ldf + f] *= coef;
            if (lane == 0) sm.LOSS[2] = rowok ? (nrm - 1.f) * (nrm - 1.f) : 0.f;
        }
    }
    __syncthreads();
    // ---- S7: pass C: ub1 = gbar . W1^T -> vb1, hb1 (lane 0 of the unit) ---------------------------------------------------------
    float hb1 = 0.f;
    if (live && !kh) {
        float s = 0.f;
        const float* wr = W1 + (size_t)u * F;
        const float* gb = sm.X + 3 * ldf;
        for (int k0 = 0; k0 < F; k0 += 4) {
            const float4 x = *reinterpret_cast<const float4*>(gb + k0);
            s = fmaf(x.x, wr[k0], s);
            if (k0 + 1 < F) s = fmaf(x.y, wr[k0 + 1], s);
            if (k0 + 2 < F) s = fmaf(x.z, wr[k0 + 2], s);
            if (k0 + 3 < F) s = fmaf(x.w, wr[k0 + 3], s);
        }
        sm.H1[3 * ldh + u] = s * (1.f - h1m * h1m);                                 // vb1
        hb1 = -2.f * s * v1 * h1m;
    }
    __syncthreads();
    // ---- S8: pass D ("F" pass on the vb1 row): ub2 -> dw3 term, zb2 ------------------------------------------------------------------
    {
        const int rows1[1] = {3};
        float o[1];
        reg_pass<HQ, 1>(w.wrow, sm.H1, ldh, rows1, kh, o);
        if (live && !kh) {
            const float ub2 = o[0];
            const float h2 = sm.H2[2 * ldh + u];
            const float om = 1.f - h2 * h2;
            sm.C3[u] = ub2 * om;
            const float zb2 = -2.f * ub2 * W3[u] * h2 * om;
            sm.Z2[u] = zb2;
            sm.L2t[u * R + 2] = zb2;
        }
    }
    __syncthreads();
    // ---- S9: pass E ("B" pass on the zb2 row): hb1 += zb2 . W2 ; zb1 = hb1*(1-h1^2) -----------------------------------------------------
    {
        const int rows1[1] = {0};
        float o[1];
        reg_pass<HQ, 1>(w.wcol, sm.Z2, ldh, rows1, kh, o);
        if (live && !kh) sm.L1t[u * R + 2] = (hb1 + o[0]) * (1.f - h1m * h1m);
    }
    __syncthreads();
    // ---- S10: parameter gradients of this tile -------------------------------------------------------------------------------------------
    outer_cols<R>(gout + a.L.w2, sm.L2t, sm.H1, ldh, H, H, tid, nth, acc);
    rowsum_store<R>(gout + a.L.b2, sm.L2t, H, tid, nth, acc, 3);
    if ((F & 3) == 0) outer_cols<R>(gout + a.L.w1, sm.L1t, sm.X, ldf, H, F, tid, nth, acc);
    else outer_store<R, 1>(gout + a.L.w1, sm.L1t, sm.X, ldf, H, F, tid, nth, acc);
    rowsum_store<R>(gout + a.L.b1, sm.L1t, H, tid, nth, acc, 3);
    for (int n = tid; n < H; n += nth) {
        const float s = sm.DD[0] * sm.H2[n] + sm.DD[1] * sm.H2[ldh + n] + sm.C3[n];
        __stcg(gout + a.L.w3 + n, acc ? s + __ldcg(gout + a.L.w3 + n) : s);
    }
    if (tid == nth - 1) {
        float db3 = sm.DD[0] + sm.DD[1], le = sm.LOSS[0], lp = sm.LOSS[1], gp = sm.LOSS[2];
        if (acc) { db3 += __ldcg(gout + a.L.b3); le += lossout[0]; lp += lossout[1]; gp += lossout[2]; }
        __stcg(gout + a.L.b3, db3);
        lossout[0] = le; lossout[1] = lp; lossout[2] = gp;
    }
    if (pf_next) disc_prefetch_rows(a, pf, ldf);                        // stage 2: the feature rows themselves
    __syncthreads();
}

}  // namespace sg
