// Actor-critic tile forward shared by sg_policy_forward (act / get_value / evaluate_actions) and the
// PPO step kernel.  One CTA of 256 threads owns R minibatch rows; threads 0..127 run the actor trunk,
// threads 128..255 the critic trunk (A2C/model.py:255-264), then the Gaussian mean head
// (A2C/distributions.py:109-110) and critic_linear.
#pragma once
#include "sg_common.cuh"

namespace sg {

#define SG_LOG_SQRT_2PI 0.91893853320467274178f   // log(sqrt(2*pi))
#define SG_LOG_2PI 1.83787706640934548356f        // log(2*pi)

template <int R>
struct PolicyTile {
    float *X, *ACT, *H1, *H2, *MU, *VAL;
    int ldo, lda, ldh;
    // returns floats consumed
    __device__ __host__ static int floats(int O, int H, int A) {
        return R * round_up(O, 4) + R * round_up(A, 4) + 4 * R * round_up(H, 4) + R * round_up(A, 4) + R;
    }
    __device__ float* carve(float* sm, int O, int H, int A) {
        ldo = round_up(O, 4); lda = round_up(A, 4); ldh = round_up(H, 4);
        X = sm; sm += R * ldo;
        ACT = sm; sm += R * lda;
        H1 = sm; sm += 2 * R * ldh;
        H2 = sm; sm += 2 * R * ldh;
        MU = sm; sm += R * lda;
        VAL = sm; sm += R;
        return sm;
    }
};

// X must be loaded (zero padded to ldo) and visible.  Ends with a CTA barrier; MU and VAL are valid after.
template <int R, class WL = LdGlobal>
__device__ __forceinline__ void policy_tile_forward(const float* __restrict__ params, const PolicyLayout& L, int O,
                                                    int H, int A, const PolicyTile<R>& T, int tid) {
    const int half = tid >> 7, t = tid & (kHalf - 1);
    const float* W1 = params + (half ? L.cw1 : L.aw1);
    const float* B1 = params + (half ? L.cb1 : L.ab1);
    const float* W2 = params + (half ? L.cw2 : L.aw2);
    const float* B2 = params + (half ? L.cb2 : L.ab2);
    float* h1 = T.H1 + half * R * T.ldh;
    float* h2 = T.H2 + half * R * T.ldh;
    const int ldh = T.ldh;
    auto epi1 = [&](int r, int n, float s) { h1[r * ldh + n] = tanhf(s + WL::ld(B1 + n)); };
    if ((O & 3) == 0) gemm_xwT<R, 4, WL>(W1, T.X, T.ldo, H, O, t, kHalf, epi1);
    else gemm_xwT<R, 1, WL>(W1, T.X, T.ldo, H, O, t, kHalf, epi1);
    __syncthreads();
    auto epi2 = [&](int r, int n, float s) { h2[r * ldh + n] = tanhf(s + WL::ld(B2 + n)); };
    if ((H & 3) == 0) gemm_xwT<R, 4, WL>(W2, h1, ldh, H, H, t, kHalf, epi2);
    else gemm_xwT<R, 1, WL>(W2, h1, ldh, H, H, t, kHalf, epi2);
    __syncthreads();
    const float* WH = params + (half ? L.vw : L.mw);
    const float* BH = params + (half ? L.vb : L.mb);
    const int NH = half ? 1 : A;
    float* out = half ? T.VAL : T.MU;
    const int ldout = half ? 1 : T.lda;
    auto epih = [&](int r, int n, float s) { out[r * ldout + n] = s + WL::ld(BH + n); };
    if ((H & 3) == 0) gemm_xwT<R, 4, WL>(WH, h2, ldh, NH, H, t, kHalf, epih);
    else gemm_xwT<R, 1, WL>(WH, h2, ldh, NH, H, t, kHalf, epih);
    __syncthreads();
}

// log-prob of `act` under N(mu, exp(logstd)) summed over the action dim, in the op order of
// torch.distributions.Normal.log_prob (called from A2C/distributions.py:52-53).
template <class WL = LdGlobal>
__device__ __forceinline__ float gaussian_logp_row(const float* mu, const float* act, const float* logstd, int A) {
    float lp = 0.f;
    for (int a = 0; a < A; ++a) {
        const float sigma = expf(WL::ld(logstd + a));
        const float var = sigma * sigma;
        const float d = act[a] - mu[a];
        lp += -(d * d) / (2.f * var) - logf(sigma) - SG_LOG_SQRT_2PI;
    }
    return lp;
}

template <class WL = LdGlobal>
__device__ __forceinline__ float gaussian_entropy(const float* logstd, int A) {
    float e = 0.f;
    for (int a = 0; a < A; ++a) e += 0.5f + 0.5f * SG_LOG_2PI + logf(expf(WL::ld(logstd + a)));
    return e;
}

}  // namespace sg
