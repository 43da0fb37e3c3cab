// SplitPolicy on the device (third_party/a2c_ppo_acktr/model_split.py:39-95, 157-238): the policy class the
// reference's shipped training scripts use (--use-split-pi).  Three 2x(Linear+Tanh) trunks on the same
// observation -- contact actor, actuator actor, critic -- and a diagonal Gaussian whose mean AND log-std are
// linear heads of the actor trunks (state-dependent log-std), contact actions first, actuator actions after.
//
//   sg_split_forward      act / get_value / evaluate_actions (model_split.py:69-95)
//   sg_split_ppo_update   PPO.update (A2C/algo/ppo.py:65-157) for this policy: same three-phase persistent
//                         kernel as sg_ppo.cu (tile phase -> slice reduce + sum of squares -> clip + Adam),
//                         384 threads per CTA = one 128-thread group per trunk, weights read through L2.
//
// Flat parameter layout (sg_split_layout): per net j in {contact, actuator, critic}: W1 (H,O), b1, W2 (H,H), b2;
// then the heads as ONE matrix per net: contact [mean (4f,H); logstd (4f,H)], actuator [mean (3f,H); logstd (3f,H)],
// critic [w3 (1,H)], each followed by its bias vector in the same order.  The 22 nn.Parameters of the Python
// module are views into that vector, so the head pairs are contiguous although nn.Module.parameters() lists the
// two means before the two log-stds.
#include "sg_common.cuh"
#include "sg_policy.cuh"
#include "sg_dp.cuh"
#include <string.h>

namespace sg {

DpView dp_view(const void* ctx);

constexpr int kSplitThreads = 384;
constexpr int kNets = 3;

struct SplitLayout {
    int w1[kNets], b1[kNets], w2[kNets], b2[kNets], wh[kNets], bh[kNets];
    int nh[kNets];        // head outputs: 8f, 6f, 1
    int na[kNets];        // action (mean) outputs: 4f, 3f, 0
    int total;
};

__host__ __device__ inline SplitLayout make_split_layout(int O, int H, int f) {
    SplitLayout L;
    int o = 0;
    for (int j = 0; j < kNets; ++j) {
        L.w1[j] = o; o += round_up(H * O, 4);
        L.b1[j] = o; o += round_up(H, 4);
        L.w2[j] = o; o += round_up(H * H, 4);
        L.b2[j] = o; o += round_up(H, 4);
    }
    L.na[0] = 4 * f; L.na[1] = 3 * f; L.na[2] = 0;
    L.nh[0] = 8 * f; L.nh[1] = 6 * f; L.nh[2] = 1;
    for (int j = 0; j < kNets; ++j) {
        L.wh[j] = o; o += round_up(L.nh[j] * H, 4);
        L.bh[j] = o; o += round_up(L.nh[j], 4);
    }
    L.total = o;
    return L;
}

template <int R>
struct SplitSmem {
    float *X, *ACT, *H1, *H2, *OUT, *ROW, *DHt, *DZ2t, *DZ1t, *SCR;
    int ldo, lda, ldh, ldq;
    __host__ __device__ static int floats(int O, int H, int f) {
        const int ldo = round_up(O, 4), lda = round_up(7 * f, 4), ldh = round_up(H, 4), ldq = round_up(8 * f, 4);
        return R * ldo + R * lda + 2 * kNets * R * ldh + kNets * R * ldq + 8 * R + kNets * ldq * R + 2 * kNets * R * ldh +
               kNets * 128 * R * 4;
    }
    __device__ void carve(float* sm, int O, int H, int f) {
        ldo = round_up(O, 4); lda = round_up(7 * f, 4); ldh = round_up(H, 4); ldq = round_up(8 * f, 4);
        X = sm; sm += R * ldo;
        ACT = sm; sm += R * lda;
        H1 = sm; sm += kNets * R * ldh;
        H2 = sm; sm += kNets * R * ldh;
        OUT = sm; sm += kNets * R * ldq;        // head outputs [net][row][ldq]
        ROW = sm; sm += 8 * R;
        DHt = sm; sm += kNets * ldq * R;        // head seeds  [net][n][R]
        DZ2t = sm; sm += kNets * R * ldh;
        DZ1t = sm; sm += kNets * R * ldh;
        SCR = sm;
    }
};

// X must be loaded (zero padded).  Ends with a CTA barrier; OUT holds the three nets' head outputs.
// W2: base of the three trunks' second-layer matrices, stride hh4 floats per net, read with WL2 -- the flat global
// vector (through L2) or the CTA's shared-memory copy (split_ppo_kernel<R, true>).
template <int R, class WL2 = LdGlobal>
__device__ __forceinline__ void split_tile_forward(const float* __restrict__ W, const SplitLayout& L, int O, int H,
                                                   const SplitSmem<R>& sm, int tid, const float* __restrict__ W2, int hh4) {
    const int net = tid >> 7, t = tid & 127;
    const int ldh = sm.ldh;
    float* h1 = sm.H1 + net * R * ldh;
    float* h2 = sm.H2 + net * R * ldh;
    float* out = sm.OUT + net * R * sm.ldq;
    const float* B1 = W + L.b1[net]; const float* B2 = W + L.b2[net]; const float* BH = W + L.bh[net];
    auto e1 = [&](int r, int n, float s) { h1[r * ldh + n] = tanhf(s + ld_cg(B1 + n)); };
    if ((O & 3) == 0) gemm_xwT<R, 4>(W + L.w1[net], sm.X, sm.ldo, H, O, t, 128, e1);
    else gemm_xwT<R, 1>(W + L.w1[net], sm.X, sm.ldo, H, O, t, 128, e1);
    __syncthreads();
    auto e2 = [&](int r, int n, float s) { h2[r * ldh + n] = tanhf(s + ld_cg(B2 + n)); };
    if ((H & 3) == 0) gemm_xwT<R, 4, WL2>(W2 + net * hh4, h1, ldh, H, H, t, 128, e2);
    else gemm_xwT<R, 1, WL2>(W2 + net * hh4, h1, ldh, H, H, t, 128, e2);
    __syncthreads();
    const int ldq = sm.ldq;
    auto eh = [&](int r, int n, float s) { out[r * ldq + n] = s + ld_cg(BH + n); };
    if ((H & 3) == 0) gemm_xwT<R, 4>(W + L.wh[net], h2, ldh, L.nh[net], H, t, 128, eh);
    else gemm_xwT<R, 1>(W + L.wh[net], h2, ldh, L.nh[net], H, t, 128, eh);
    __syncthreads();
}

// mean / log-std of action k of row r from the head outputs
template <int R>
__device__ __forceinline__ void split_mu_ls(const SplitSmem<R>& sm, const SplitLayout& L, int r, int k, float& mu, float& ls) {
    const int n0 = L.na[0];
    if (k < n0) {
        const float* o = sm.OUT + r * sm.ldq;
        mu = o[k]; ls = o[n0 + k];
    } else {
        const float* o = sm.OUT + (R + r) * sm.ldq;
        mu = o[k - n0]; ls = o[L.na[1] + (k - n0)];
    }
}

// ---- forward kernel: act / get_value / evaluate_actions ---------------------------------------------------
template <int R>
__global__ void __launch_bounds__(kSplitThreads) split_forward_kernel(const float* __restrict__ params, SplitLayout L, int O, int H,
                                                                      int f, const float* __restrict__ obs, int B,
                                                                      const float* __restrict__ noise,
                                                                      const float* __restrict__ actions_in,
                                                                      float* __restrict__ value, float* __restrict__ action,
                                                                      float* __restrict__ logp, float* __restrict__ entropy_rows) {
    extern __shared__ __align__(16) float smem[];
    SplitSmem<R> sm;
    sm.carve(smem, O, H, f);
    const int tid = threadIdx.x, A = 7 * f;
    for (int tile = blockIdx.x; tile * R < B; tile += gridDim.x) {
        const int row0 = tile * R;
        for (int e = tid; e < R * sm.ldo; e += kSplitThreads) {
            const int r = e / sm.ldo, k = e - r * sm.ldo;
            sm.X[e] = (row0 + r < B && k < O) ? obs[(size_t)(row0 + r) * O + k] : 0.f;
        }
        __syncthreads();
        split_tile_forward<R>(params, L, O, H, sm, tid, params + L.w2[0], L.w2[1] - L.w2[0]);
        for (int e = tid; e < R * A; e += kSplitThreads) {
            const int r = e / A, k = e - r * A;
            const int row = row0 + r;
            float act = 0.f;
            if (row < B) {
                float mu, ls;
                split_mu_ls<R>(sm, L, r, k, mu, ls);
                if (actions_in) act = actions_in[(size_t)row * A + k];
                else if (noise) act = __fadd_rn(__fmul_rn(noise[(size_t)row * A + k], expf(ls)), mu);
                else act = mu;
                if (action) action[(size_t)row * A + k] = act;
            }
            sm.ACT[r * sm.lda + k] = act;
        }
        __syncthreads();
        if (tid < R && row0 + tid < B) {
            const int r = tid, row = row0 + tid;
            float lp = 0.f, ent = 0.f;
            for (int k = 0; k < A; ++k) {
                float mu, ls;
                split_mu_ls<R>(sm, L, r, k, mu, ls);
                const float sigma = expf(ls);
                const float d = sm.ACT[r * sm.lda + k] - mu;
                lp += -(d * d) / (2.f * (sigma * sigma)) - logf(sigma) - SG_LOG_SQRT_2PI;
                ent += 0.5f + 0.5f * SG_LOG_2PI + logf(sigma);
            }
            if (value) value[row] = sm.OUT[(2 * R + r) * sm.ldq];
            if (logp) logp[row] = lp;
            if (entropy_rows) entropy_rows[row] = ent;
        }
        __syncthreads();
    }
}

// ---- PPO update ----------------------------------------------------------------------------------------------
struct SplitArgs {
    int O, H, f, A, S, P;
    int nmb, mbs, nsteps, row_begin, row_end, ntiles, nslots, SL, nslices;
    int clipped_vloss;
    float clip, ratio_lo, ratio_hi, c_v, c_e, max_norm;
    float one_minus_b1, b2, one_minus_b2, eps;
    SplitLayout L;
    float *params, *m, *v;
    const float *obs, *actions, *vpred, *ret, *oldlp, *advstats;
    const int32_t* perm;
    const float *step_size, *bc2_sqrt;
    float* trace;
    float *gpart, *grad, *losspart;
    double* ssq;
    unsigned int* bar;
    int first_adam_step;
    int dp_on;            // fused peer-memory gradient exchange (sg_dp.cuh), as in sg_ppo.cu
    DpView dp;
};

// MULTI: the CTA runs several tiles per step and accumulates across them; single-tile kernels compile without any
// accumulate code (with a run-time flag the compiler issues the old-value loads speculatively, exposed L2 latency).
template <int R, class WL2, bool MULTI>
__device__ void split_tile(const SplitArgs& a, int step, int tile, float* __restrict__ gout, float* __restrict__ lossout,
                           SplitSmem<R>& sm, bool acc_in, const float* __restrict__ W2, int hh4) {
    const bool acc = MULTI && acc_in;
    static_assert(R == 8 || R == 4, "loss warp maps R rows x 32/R lanes");
    constexpr int LPR = 32 / R;           // lanes per row in the loss warp
    const int tid = threadIdx.x;
    const int O = a.O, H = a.H, A = a.A;
    const SplitLayout& L = a.L;
    const float* W = a.params;
    const int epoch = step / a.nmb, mb = step - epoch * a.nmb;
    const int32_t* idx = a.perm + (size_t)epoch * a.S + (size_t)mb * a.mbs;
    const int row0 = a.row_begin + tile * R;
    float* rRet = sm.ROW; float* rVp = sm.ROW + R; float* rOlp = sm.ROW + 2 * R; float* rAdv = sm.ROW + 3 * R;
    float* rValid = sm.ROW + 4 * R;
    for (int e = tid; e < R * sm.ldo; e += kSplitThreads) {
        const int r = e / sm.ldo, k = e - r * sm.ldo;
        const int row = row0 + r;
        sm.X[e] = (row < a.row_end && k < O) ? a.obs[(size_t)idx[row] * O + k] : 0.f;
    }
    for (int e = tid; e < R * sm.lda; e += kSplitThreads) {
        const int r = e / sm.lda, k = e - r * sm.lda;
        const int row = row0 + r;
        sm.ACT[e] = (row < a.row_end && k < A) ? a.actions[(size_t)idx[row] * A + k] : 0.f;
    }
    if (tid >= kSplitThreads - R) {
        const int r = tid - (kSplitThreads - R);
        const int row = row0 + r;
        const bool ok = row < a.row_end;
        const int i = ok ? idx[row] : 0;
        const float ret = ok ? a.ret[i] : 0.f, vp = ok ? a.vpred[i] : 0.f;
        rRet[r] = ret; rVp[r] = vp; rOlp[r] = ok ? a.oldlp[i] : 0.f;
        const float mean = a.advstats[0], sd = a.advstats[1];
        rAdv[r] = ok ? __fdiv_rn(__fsub_rn(__fsub_rn(ret, vp), mean), __fadd_rn(sd, 1e-5f)) : 0.f;   // ppo.py:66-68
        rValid[r] = ok ? 1.f : 0.f;
    }
    __syncthreads();
    split_tile_forward<R, WL2>(W, L, O, H, sm, tid, W2, hh4);

    // per-row losses and head seeds: one warp, LPR lanes per row
    if (tid < 32) {
        const int r = tid / LPR, sub = tid % LPR;
        const bool ok = rValid[r] != 0.f;
        const float invB = 1.f / (float)a.mbs;
        float lp = 0.f, ent = 0.f;
        for (int k = sub; k < A; k += LPR) {
            float mu, ls;
            split_mu_ls<R>(sm, L, r, k, mu, ls);
            const float sigma = expf(ls);
            const float d = sm.ACT[r * sm.lda + k] - mu;
            lp += -(d * d) / (2.f * (sigma * sigma)) - logf(sigma) - SG_LOG_SQRT_2PI;
            ent += 0.5f + 0.5f * SG_LOG_2PI + logf(sigma);
        }
#pragma unroll
        for (int o = 1; o < LPR; o <<= 1) {
            lp += __shfl_xor_sync(0xffffffffu, lp, o);
            ent += __shfl_xor_sync(0xffffffffu, ent, o);
        }
        float vl = 0.f, al = 0.f, dv = 0.f, coef = 0.f;
        if (ok) {
            const float ratio = expf(lp - rOlp[r]);
            const float adv = rAdv[r];
            const float s1 = ratio * adv;
            const float s2 = fminf(fmaxf(ratio, a.ratio_lo), a.ratio_hi) * adv;
            al = -fminf(s1, s2);
            const float inr = (ratio >= a.ratio_lo && ratio <= a.ratio_hi) ? 1.f : 0.f;
            const float gsel = s1 < s2 ? 1.f : (s2 < s1 ? inr : 0.5f + 0.5f * inr);     // torch.min / clamp backward
            coef = -invB * gsel * adv * ratio;
            const float v = sm.OUT[(2 * R + r) * sm.ldq], vp = rVp[r], ret = rRet[r];
            if (a.clipped_vloss) {
                const float diff = v - vp;
                const float vc = vp + fminf(fmaxf(diff, -a.clip), a.clip);
                const float e1 = v - ret, e2 = vc - ret;
                const float l1 = e1 * e1, l2 = e2 * e2;
                vl = 0.5f * fmaxf(l1, l2);
                const float in2 = (diff >= -a.clip && diff <= a.clip) ? 1.f : 0.f;
                const float g = l1 > l2 ? e1 : (l2 > l1 ? in2 * e2 : 0.5f * (e1 + in2 * e2));
                dv = a.c_v * invB * g;
            } else {
                const float e1 = ret - v;
                vl = 0.5f * e1 * e1;
                dv = a.c_v * invB * (v - ret);
            }
        } else {
            ent = 0.f;
        }
        // seeds: d loss / d mean, d loss / d logstd (log-prob term + the entropy bonus -c_e * mean_r H_r)
        for (int k = sub; k < A; k += LPR) {
            float mu, ls;
            split_mu_ls<R>(sm, L, r, k, mu, ls);
            const float sigma = expf(ls);
            const float var = sigma * sigma;
            const float d = sm.ACT[r * sm.lda + k] - mu;
            const float dmu = ok ? coef * d / var : 0.f;
            const float dls = ok ? coef * (d * d / var - 1.f) - a.c_e * invB : 0.f;
            const int net = k < L.na[0] ? 0 : 1;
            const int kk = k - (net ? L.na[0] : 0);
            float* dh = sm.DHt + net * sm.ldq * R;
            dh[kk * R + r] = dmu;
            dh[(L.na[net] + kk) * R + r] = dls;
        }
        if (sub == 0) sm.DHt[2 * sm.ldq * R + r] = dv;
        float svl = sub == 0 ? vl : 0.f, sal = sub == 0 ? al : 0.f, sen = sub == 0 ? ent : 0.f;
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
            svl += __shfl_xor_sync(0xffffffffu, svl, o);
            sal += __shfl_xor_sync(0xffffffffu, sal, o);
            sen += __shfl_xor_sync(0xffffffffu, sen, o);
        }
        if (tid == 0) {
            if (acc) { svl += lossout[0]; sal += lossout[1]; sen += lossout[2]; }
            lossout[0] = svl; lossout[1] = sal; lossout[2] = sen;
        }
    }
    __syncthreads();

    // ---- backward, one 128-thread group per net ------------------------------------------------------------------
    const int net = tid >> 7, t = tid & 127;
    const int ldh = sm.ldh;
    const bool vecH = (H & 3) == 0, vecO = (O & 3) == 0;
    const float* h1 = sm.H1 + net * R * ldh;
    const float* h2 = sm.H2 + net * R * ldh;
    float* dz2 = sm.DZ2t + net * R * ldh;
    float* dz1 = sm.DZ1t + net * R * ldh;
    float* scr = sm.SCR + net * 128 * R * 4;
    const float* Dh = sm.DHt + net * sm.ldq * R;
    const int NH = L.nh[net];
    auto epi_h = [&](int r, int k, float s) { const float h = h2[r * ldh + k]; dz2[k * R + r] = s * (1.f - h * h); };
    if (vecH) gemm_yW<R, 4>(W + L.wh[net], Dh, NH, H, scr, t, 128, epi_h);
    else gemm_yW<R, 1>(W + L.wh[net], Dh, NH, H, scr, t, 128, epi_h);
    if (vecH) outer_store<R, 4>(gout + L.wh[net], Dh, h2, ldh, NH, H, t, 128, acc);
    else outer_store<R, 1>(gout + L.wh[net], Dh, h2, ldh, NH, H, t, 128, acc);
    rowsum_store<R>(gout + L.bh[net], Dh, NH, t, 128, acc);
    auto epi_2 = [&](int r, int k, float s) { const float h = h1[r * ldh + k]; dz1[k * R + r] = s * (1.f - h * h); };
    if (vecH) gemm_yW<R, 4, WL2>(W2 + net * hh4, dz2, H, H, scr, t, 128, epi_2);
    else gemm_yW<R, 1, WL2>(W2 + net * hh4, dz2, H, H, scr, t, 128, epi_2);
    if (vecH) outer_store<R, 4>(gout + L.w2[net], dz2, h1, ldh, H, H, t, 128, acc);
    else outer_store<R, 1>(gout + L.w2[net], dz2, h1, ldh, H, H, t, 128, acc);
    rowsum_store<R>(gout + L.b2[net], dz2, H, t, 128, acc);
    if (vecO) outer_store<R, 4>(gout + L.w1[net], dz1, sm.X, sm.ldo, H, O, t, 128, acc);
    else outer_store<R, 1>(gout + L.w1[net], dz1, sm.X, sm.ldo, H, O, t, 128, acc);
    rowsum_store<R>(gout + L.b1[net], dz1, H, t, 128, acc);
    __syncthreads();
}

// W2RES: the three H x H second-layer matrices (the bulk of the weights, walked twice per tile) live in a
// shared-memory copy that every CTA refreshes with three TMA bulk copies after each Adam step; everything else is
// read through L2.  (The whole parameter vector + tile does not fit 227 KB at the shipped hidden size of 100.)
template <int R, bool W2RES, bool MULTI>
__global__ void __launch_bounds__(kSplitThreads, 1) split_ppo_kernel(SplitArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ double red[kSplitThreads / 32];
    __shared__ __align__(8) unsigned long long img_bar;
    constexpr int NT = kSplitThreads;
    const int tid = threadIdx.x, cta = blockIdx.x;
    SplitSmem<R> sm;
    sm.carve(smem, a.O, a.H, a.f);
    const int hh4 = a.L.w2[1] - a.L.w2[0];                              // global stride between the nets' W2 blocks
    const int hhs = round_up(a.H * a.H, 4);                             // stride in the shared-memory copy
    float* W2s = smem + SplitSmem<R>::floats(a.O, a.H, a.f);
    if (W2RES) {
        if (tid == 0) mbar_init(&img_bar, 1);
        __syncthreads();
    }
    GridBarrier gb{a.bar, a.bar + 1, gridDim.x, 0};
    const int p0 = min(a.P, cta * a.SL), p1 = min(a.P, p0 + a.SL);
    for (int step = 0; step < a.nsteps; ++step) {
        // A: tile phase
        bool acc = false;
        const bool work = cta < a.ntiles;
        if (W2RES && work) {
            if (tid == 0) {
                fence_proxy_async();        // parameters were written by other CTAs' generic stores (grid barrier acquired)
                mbar_expect_tx(&img_bar, (unsigned int)(kNets * hhs * sizeof(float)));
                for (int j = 0; j < kNets; ++j)
                    tma_bulk_g2s(W2s + j * hhs, a.params + a.L.w2[j], (unsigned int)(hhs * sizeof(float)), &img_bar);
            }
            mbar_wait(&img_bar, (unsigned int)(step & 1));
        }
        for (int tile = cta; tile < a.ntiles; tile += gridDim.x) {
            if (W2RES) split_tile<R, LdShared, MULTI>(a, step, tile, a.gpart + (size_t)cta * a.P, a.losspart + cta * 4, sm, acc, W2s, hhs);
            else split_tile<R, LdGlobal, MULTI>(a, step, tile, a.gpart + (size_t)cta * a.P, a.losspart + cta * 4, sm, acc, a.params + a.L.w2[0], hh4);
            acc = true;
        }
        gb.sync();
        // B: slice reduce + sum of squares (+ loss sums by CTA 0)
        float4 mine;
        reduce_partials_slice<NT, false>(a.gpart, (size_t)a.P, a.nslots, p0, p1, a.grad, reinterpret_cast<float4*>(smem), tid, mine);
        if (cta == 0 && tid < 32) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
            for (int c = tid; c < a.nslots; c += 32) {
                s0 += ld_cg(a.losspart + c * 4); s1 += ld_cg(a.losspart + c * 4 + 1); s2 += ld_cg(a.losspart + c * 4 + 2);
            }
            s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
            if (tid == 0) { __stcg(a.grad + a.P, s0); __stcg(a.grad + a.P + 1, s1); __stcg(a.grad + a.P + 2, s2); }
        }
        // data parallel: swap my locally reduced slice with the peers' over NVLink, keep the rank-ordered total
        if (a.dp_on) dp_exchange_slice<NT>(a.dp, a.grad, p0, p1, cta, (unsigned int)(a.first_adam_step + step), nullptr);
        {
            double s = 0.0;
            for (int p = p0 + 4 * tid; p < p1; p += 4 * NT) {
                const float4 g = ld_cg4(a.grad + p);
                s += (double)g.x * (double)g.x; s += (double)g.y * (double)g.y; s += (double)g.z * (double)g.z; s += (double)g.w * (double)g.w;
            }
            const double tot = block_sum<NT>(s, red);
            if (tid == 0) __stcg(a.ssq + cta, tot);
        }
        gb.sync();
        // C: global-norm clip + Adam on the slice
        {
            double s = 0.0;
            for (int c = tid; c < a.nslices; c += NT) s += __ldcg(a.ssq + c);
            const double tot = block_sum<NT>(s, red);
            const float norm = (float)sqrt(tot);
            float clip = a.max_norm / (norm + 1e-6f);
            if (clip > 1.f) clip = 1.f;
            if (cta == 0 && tid == 0) {
                const float invB = 1.f / (float)a.mbs;
                float* tr = a.trace + (size_t)step * 4;
                tr[0] = ld_cg(a.grad + a.P) * invB; tr[1] = ld_cg(a.grad + a.P + 1) * invB;
                tr[2] = ld_cg(a.grad + a.P + 2) * invB; tr[3] = norm;
            }
            const float ss = a.step_size[step], bc2 = a.bc2_sqrt[step];
            for (int p = p0 + tid; p < p1; p += NT) {
                const float g = ld_cg(a.grad + p) * clip;
                float pv = __ldcg(a.params + p), mv = __ldcg(a.m + p), vv = __ldcg(a.v + p);
                adam_update(pv, mv, vv, g, a.one_minus_b1, a.b2, a.one_minus_b2, ss, bc2, a.eps);
                __stcg(a.params + p, pv); __stcg(a.m + p, mv); __stcg(a.v + p, vv);
            }
        }
        gb.sync();
    }
    if (cta == 0 && tid == 0 && *(volatile unsigned int*)(a.bar + 1) != 0u) a.trace[0] = __int_as_float(0x7fc00000);
}

static int split_feet(const sg_ppo_config* c) { return c->act_dim / 7; }
// rows per tile: 8, or 4 when that still leaves at most one tile per SM (twice the CTAs on half the rows each)
static int split_rows(const sg_ppo_config* c) {
    int sms = sg_device_sm_count();
    if (sms <= 0) sms = 148;
    const int rows = c->row_end - c->row_begin;
    return (rows + 3) / 4 <= sms ? 4 : kRows;
}
static int split_tiles(const sg_ppo_config* c) { const int r = split_rows(c); return (c->row_end - c->row_begin + r - 1) / r; }
static int split_grid(const sg_ppo_config* c, int* sms_out) {
    int sms = sg_device_sm_count();
    if (sms <= 0) sms = 148;
    if (sms_out) *sms_out = sms;
    int tiles = split_tiles(c);
    int g = tiles < sms ? tiles : sms;
    // CTAs without a tile still own a slice of the reduce / clip / Adam phases; this parameter vector is 3-4x the
    // plain policy's, so more (narrower) slices pay for the slightly slower grid barrier
    const int gmin = sms < 128 ? sms : 128;
    if (g < gmin) g = gmin;
    return g;
}
static size_t split_smem_bytes(int O, int H, int f) {
    size_t fl = (size_t)SplitSmem<kRows>::floats(O, H, f);
    if (fl < 4 * (kSplitThreads + 128)) fl = 4 * (kSplitThreads + 128);
    return fl * sizeof(float);
}
static int split_validate(const sg_ppo_config* c) {
    SG_REQUIRE(c, "sg_split_ppo: null config");
    SG_REQUIRE(c->obs_dim > 0 && c->hidden > 0 && c->act_dim > 0 && c->act_dim % 7 == 0,
               "sg_split_ppo: act_dim must be 7*num_feet (4 contact + 3 actuator outputs per foot, model_split.py:205)");
    SG_REQUIRE(c->T > 0 && c->N > 0 && c->ppo_epoch > 0 && c->num_mini_batch > 0, "sg_split_ppo: non-positive sizes");
    SG_REQUIRE(c->mini_batch_size > 0 && (long long)c->mini_batch_size * c->num_mini_batch <= (long long)c->T * c->N,
               "sg_split_ppo: mini_batch_size*num_mini_batch exceeds T*N");
    SG_REQUIRE(c->row_begin >= 0 && c->row_begin < c->row_end && c->row_end <= c->mini_batch_size,
               "sg_split_ppo: shard [%d,%d) outside minibatch of %d rows", c->row_begin, c->row_end, c->mini_batch_size);
    SG_REQUIRE(c->first_adam_step >= 1, "sg_split_ppo: first_adam_step is 1-based");
    SG_REQUIRE(split_smem_bytes(c->obs_dim, c->hidden, split_feet(c)) <= 226 * 1024, "sg_split_ppo: tile does not fit shared memory");
    return SG_OK;
}
struct SplitWs { size_t gpart, grad, losspart, ssq, bar, total; };
static SplitWs split_ws(const sg_ppo_config* c, int grid) {
    SplitLayout L = make_split_layout(c->obs_dim, c->hidden, split_feet(c));
    SplitWs w;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 255) / 256 * 256; return at; };
    w.gpart = take((size_t)grid * L.total * sizeof(float));
    w.grad = take((size_t)(L.total + 4) * sizeof(float));
    w.losspart = take((size_t)grid * 4 * sizeof(float));
    w.ssq = take((size_t)grid * sizeof(double));
    w.bar = take(2 * sizeof(unsigned int));
    w.total = o;
    return w;
}

}  // namespace sg

using namespace sg;

extern "C" {
#pragma GCC visibility push(default)

int sg_split_layout(int obs_dim, int hidden, int num_feet, int* offsets) {
    if (obs_dim <= 0 || hidden <= 0 || num_feet <= 0) { sg::set_error("sg_split_layout: non-positive dims"); return -1; }
    SplitLayout L = make_split_layout(obs_dim, hidden, num_feet);
    if (offsets) {
        // nn.Module.parameters() order of SplitPolicy (model_split.py:172-185, 220-224)
        int i = 0;
        for (int j = 0; j < 2; ++j) { offsets[i++] = L.w1[j]; offsets[i++] = L.b1[j]; offsets[i++] = L.w2[j]; offsets[i++] = L.b2[j]; }
        offsets[i++] = L.w1[2]; offsets[i++] = L.b1[2]; offsets[i++] = L.w2[2]; offsets[i++] = L.b2[2];
        offsets[i++] = L.wh[2]; offsets[i++] = L.bh[2];                                            // critic_full.4
        offsets[i++] = L.wh[0]; offsets[i++] = L.bh[0];                                            // contact_mean
        offsets[i++] = L.wh[1]; offsets[i++] = L.bh[1];                                            // actuator_mean
        offsets[i++] = L.wh[0] + L.na[0] * hidden; offsets[i++] = L.bh[0] + L.na[0];               // contact_logstd
        offsets[i++] = L.wh[1] + L.na[1] * hidden; offsets[i++] = L.bh[1] + L.na[1];               // actuator_logstd
    }
    return L.total;
}

int sg_split_forward(const float* params, int obs_dim, int hidden, int num_feet, const float* obs, int B, const float* noise,
                     const float* actions_in, float* value, float* action, float* logp, float* entropy_rows, void* stream) {
    SG_REQUIRE(params && obs && B > 0 && obs_dim > 0 && hidden > 0 && num_feet > 0, "sg_split_forward: bad arguments");
    SplitLayout L = make_split_layout(obs_dim, hidden, num_feet);
    const size_t smem = split_smem_bytes(obs_dim, hidden, num_feet);
    SG_REQUIRE(smem <= 226 * 1024, "sg_split_forward: tile needs %zu bytes of shared memory", smem);
    static SmemGrant grant;
    if (int rc = grant_smem(grant, split_forward_kernel<kRows>, smem)) return rc;
    int tiles = (B + kRows - 1) / kRows;
    int grid = tiles < 592 ? tiles : 592;
    split_forward_kernel<kRows><<<grid, kSplitThreads, smem, (cudaStream_t)stream>>>(params, L, obs_dim, hidden, num_feet, obs, B, noise,
                                                                                      actions_in, value, action, logp, entropy_rows);
    count_launches(1);
    SG_CUDA(cudaGetLastError());
    return SG_OK;
}

int64_t sg_split_ppo_workspace_bytes(const sg_ppo_config* cfg) {
    if (split_validate(cfg)) return -1;
    return (int64_t)split_ws(cfg, split_grid(cfg, nullptr)).total;
}

int sg_split_ppo_update(const sg_ppo_config* cfg, float* params, float* adam_m, float* adam_v, const float* obs,
                        const float* actions, const float* value_preds, const float* returns, const float* old_logp,
                        const float* adv_stats, const int32_t* perm, const float* step_size, const float* bc2_sqrt,
                        float* trace, void* workspace, void* stream) {
    int rc = split_validate(cfg);
    if (rc) return rc;
    SG_REQUIRE(params && adam_m && adam_v && obs && actions && value_preds && returns && old_logp && adv_stats && perm &&
                   step_size && bc2_sqrt && trace && workspace, "sg_split_ppo_update: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    int sms = 0;
    const int grid = split_grid(cfg, &sms);
    const SplitWs w = split_ws(cfg, grid);
    char* ws = (char*)workspace;
    SplitArgs a;
    a.O = cfg->obs_dim; a.H = cfg->hidden; a.f = split_feet(cfg); a.A = cfg->act_dim;
    a.S = cfg->T * cfg->N;
    a.L = make_split_layout(a.O, a.H, a.f);
    a.P = a.L.total;
    a.nmb = cfg->num_mini_batch; a.mbs = cfg->mini_batch_size;
    a.nsteps = cfg->ppo_epoch * cfg->num_mini_batch;
    a.row_begin = cfg->row_begin; a.row_end = cfg->row_end;
    a.ntiles = split_tiles(cfg);
    a.nslots = grid < a.ntiles ? grid : a.ntiles;
    a.SL = round_up((a.P + grid - 1) / grid, 4);
    a.nslices = (a.P + a.SL - 1) / a.SL;
    a.clipped_vloss = cfg->use_clipped_value_loss;
    a.clip = (float)cfg->clip_param;
    a.ratio_lo = (float)(1.0 - cfg->clip_param); a.ratio_hi = (float)(1.0 + cfg->clip_param);
    a.c_v = (float)cfg->value_loss_coef; a.c_e = (float)cfg->entropy_coef; a.max_norm = (float)cfg->max_grad_norm;
    a.one_minus_b1 = (float)(1.0 - cfg->beta1); a.b2 = (float)cfg->beta2; a.one_minus_b2 = (float)(1.0 - cfg->beta2);
    a.eps = (float)cfg->adam_eps;
    a.params = params; a.m = adam_m; a.v = adam_v;
    a.obs = obs; a.actions = actions; a.vpred = value_preds; a.ret = returns; a.oldlp = old_logp; a.advstats = adv_stats;
    a.perm = perm; a.step_size = step_size; a.bc2_sqrt = bc2_sqrt; a.trace = trace;
    a.gpart = (float*)(ws + w.gpart); a.grad = (float*)(ws + w.grad); a.losspart = (float*)(ws + w.losspart);
    a.ssq = (double*)(ws + w.ssq); a.bar = (unsigned int*)(ws + w.bar);
    a.first_adam_step = cfg->first_adam_step;
    a.dp_on = cfg->dp_ctx != nullptr;
    if (a.dp_on) {
        a.dp = dp_view(cfg->dp_ctx);
        SG_REQUIRE(a.dp.cap >= a.P + 4 && a.nslices < kDpMaxSlices, "sg_split_ppo_update: dp context too small (%d floats, %d slices)", a.dp.cap, a.nslices);
    } else {
        memset(&a.dp, 0, sizeof(a.dp));
    }
    const size_t smem_tile = split_smem_bytes(a.O, a.H, a.f);
    const bool r4 = split_rows(cfg) == 4;
    const size_t tile_floats = r4 ? (size_t)SplitSmem<4>::floats(a.O, a.H, a.f) : (size_t)SplitSmem<kRows>::floats(a.O, a.H, a.f);
    const size_t smem_res = (tile_floats + (size_t)kNets * round_up(a.H * a.H, 4)) * sizeof(float);
    const bool w2res = smem_res <= 226 * 1024 && (cfg->mode == 0 || cfg->mode == 3);      // mode 2: weights through L2 only
    const size_t smem = w2res ? (smem_res > smem_tile ? smem_res : smem_tile) : smem_tile;
    SG_CUDA(cudaMemsetAsync(ws, 0, w.total, s));
    const bool multi = a.ntiles > grid;
#define SG_SPLIT_FN(R_) (w2res ? (multi ? (const void*)split_ppo_kernel<R_, true, true> : (const void*)split_ppo_kernel<R_, true, false>) \
                               : (multi ? (const void*)split_ppo_kernel<R_, false, true> : (const void*)split_ppo_kernel<R_, false, false>))
    const void* fn = r4 ? SG_SPLIT_FN(4) : SG_SPLIT_FN(kRows);
#undef SG_SPLIT_FN
    SG_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kSplitThreads, smem));
    SG_REQUIRE(per_sm >= 1 && grid <= per_sm * sms, "sg_split_ppo_update: cooperative grid of %d CTAs does not fit", grid);
    void* kargs[] = {(void*)&a};
    SG_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kSplitThreads), kargs, smem, s));
    count_launches(1);
    return SG_OK;
}

#pragma GCC visibility pop
}  // extern "C"
