// PPO.update on the device (A2C/algo/ppo.py:65-157).
//
// One optimizer step = three grid-wide phases:
//   1. tile phase   every CTA owns tiles of R=8 minibatch rows: sampler gather (A2C/storage.py:169-185)
//                   -> actor/critic forward -> Gaussian log-prob -> clipped surrogate + clipped value
//                   loss -> hand-derived backward -> per-CTA partial gradient (P floats) in L2
//   2. reduce phase partial gradients summed in a fixed order -> flat gradient (+ loss sums)
//                   [data-parallel mode: the NCCL sum-allreduce of that vector happens here]
//   3. adam phase   global-norm clip (A2C/algo/ppo.py:143-144) + Adam (ppo.py:145) on a param slice
// mode 0 runs all steps of the call inside ONE persistent cooperative kernel with grid barriers between
// phases; mode 1 launches one kernel per phase (debug / data-parallel path).
#include "sg_common.cuh"
#include "sg_policy.cuh"

namespace sg {

struct PpoArgs {
    int O, H, A, S, P;
    int nmb, mbs, nsteps, row_begin, row_end, ntiles, nslots;
    int clipped_vloss, first_adam_step;
    float clip, ratio_lo, ratio_hi, c_v, c_e, max_norm;
    float one_minus_b1, b2, one_minus_b2, eps;
    PolicyLayout L;
    float *params, *m, *v;
    const float *obs, *actions, *vpred, *ret, *oldlp, *advstats;
    const int32_t* perm;
    const float *step_size, *bc2_sqrt;
    float* trace;
    float *gpart, *grad, *losspart, *scal;
    unsigned int* bar;
};

template <int R>
struct PpoSmem {
    PolicyTile<R> T;
    float *ROW, *DMUt, *DVt, *DLSt, *DZ2t, *DZ1t, *SCR;
    __host__ __device__ static int floats(int O, int H, int A) {
        return PolicyTile<R>::floats(O, H, A) + 8 * R + round_up(A, 4) * R + R + round_up(A, 4) * R + 4 * R * round_up(H, 4) +
               2 * kHalf * R * 4;
    }
    __device__ void carve(float* sm, int O, int H, int A) {
        sm = T.carve(sm, O, H, A);
        ROW = sm; sm += 8 * R;
        DMUt = sm; sm += round_up(A, 4) * R;
        DVt = sm; sm += R;
        DLSt = sm; sm += round_up(A, 4) * R;
        DZ2t = sm; sm += 2 * R * round_up(H, 4);
        DZ1t = sm; sm += 2 * R * round_up(H, 4);
        SCR = sm;
    }
};

// ---- phase 1: one tile of R minibatch rows ------------------------------------------------------
template <int R>
__device__ void ppo_tile(const PpoArgs& a, int step, int tile, float* __restrict__ gout, float* __restrict__ lossout,
                         PpoSmem<R>& sm, bool acc) {
    const int tid = threadIdx.x;
    const int O = a.O, H = a.H, A = a.A;
    const PolicyTile<R>& T = sm.T;
    const int epoch = step / a.nmb, mb = step - epoch * a.nmb;
    const int32_t* idx = a.perm + (size_t)epoch * a.S + (size_t)mb * a.mbs;
    const int row0 = a.row_begin + tile * R;
    float* rRet = sm.ROW;            float* rVp = sm.ROW + R;     float* rOlp = sm.ROW + 2 * R;
    float* rAdv = sm.ROW + 3 * R;    float* rValid = sm.ROW + 4 * R;
    float* rVl = sm.ROW + 5 * R;     float* rAl = sm.ROW + 6 * R;

    // gather this tile's rows (flat sample id = t*N+n, A2C/storage.py:169-181)
    for (int e = tid; e < R * T.ldo; e += kStepThreads) {
        const int r = e / T.ldo, k = e - r * T.ldo;
        const int row = row0 + r;
        T.X[e] = (row < a.row_end && k < O) ? a.obs[(size_t)idx[row] * O + k] : 0.f;
    }
    for (int e = tid; e < R * T.lda; e += kStepThreads) {
        const int r = e / T.lda, k = e - r * T.lda;
        const int row = row0 + r;
        T.ACT[e] = (row < a.row_end && k < A) ? a.actions[(size_t)idx[row] * A + k] : 0.f;
    }
    if (tid < R) {
        const int row = row0 + tid;
        const bool ok = row < a.row_end;
        const int i = ok ? idx[row] : 0;
        const float ret = ok ? a.ret[i] : 0.f, vp = ok ? a.vpred[i] : 0.f;
        rRet[tid] = ret; rVp[tid] = vp; rOlp[tid] = ok ? a.oldlp[i] : 0.f;
        // (adv - mean) / (std + 1e-5)   (A2C/algo/ppo.py:66-68)
        const float mean = a.advstats[0], sd = a.advstats[1];
        rAdv[tid] = ok ? __fdiv_rn(__fsub_rn(__fsub_rn(ret, vp), mean), __fadd_rn(sd, 1e-5f)) : 0.f;
        rValid[tid] = ok ? 1.f : 0.f;
    }
    __syncthreads();

    policy_tile_forward<R>(a.params, a.L, O, H, A, T, tid);

    // per-row losses and the gradient seeds d loss / d mu, d loss / d value
    const float* ls = a.params + a.L.ls;
    if (tid < R) {
        const int r = tid;
        float vl = 0.f, al = 0.f, dv = 0.f, coef = 0.f;
        const bool ok = rValid[r] != 0.f;
        const float invB = 1.f / (float)a.mbs;
        if (ok) {
            const float lp = gaussian_logp_row(T.MU + r * T.lda, T.ACT + r * T.lda, ls, A);
            const float ratio = expf(lp - rOlp[r]);
            const float adv = rAdv[r];
            const float s1 = ratio * adv;
            const float s2 = fminf(fmaxf(ratio, a.ratio_lo), a.ratio_hi) * adv;
            al = -fminf(s1, s2);
            const float inr = (ratio >= a.ratio_lo && ratio <= a.ratio_hi) ? 1.f : 0.f;
            // torch.min backward: all to the smaller side, 1/2 + 1/2 on exact ties; clamp passes grad on
            // inclusive bounds
            const float gsel = s1 < s2 ? 1.f : (s2 < s1 ? inr : 0.5f + 0.5f * inr);
            coef = -invB * gsel * adv * ratio;
            const float v = T.VAL[r], vp = rVp[r], ret = rRet[r];
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
        }
        rVl[r] = vl; rAl[r] = al;
        sm.DVt[r] = dv;
        for (int k = 0; k < A; ++k) {
            const float sigma = expf(ld_cg(ls + k));
            const float var = sigma * sigma;
            const float d = T.ACT[r * T.lda + k] - T.MU[r * T.lda + k];
            sm.DMUt[k * R + r] = ok ? coef * d / var : 0.f;          // d logp / d mu   = (a-mu)/var
            sm.DLSt[k * R + r] = ok ? coef * (d * d / var - 1.f) : 0.f;  // d logp / d logstd = (a-mu)^2/var - 1
        }
    }
    __syncthreads();
    if (tid == 0) {
        float svl = 0.f, sal = 0.f;
        for (int r = 0; r < R; ++r) { svl += rVl[r]; sal += rAl[r]; }
        if (acc) { svl += lossout[0]; sal += lossout[1]; }
        lossout[0] = svl; lossout[1] = sal;
    }

    // ---- backward -------------------------------------------------------------------------------
    const int half = tid >> 7, t = tid & (kHalf - 1);
    const int ldh = T.ldh;
    const bool vecH = (H & 3) == 0, vecO = (O & 3) == 0;
    const float* h1 = T.H1 + half * R * ldh;
    const float* h2 = T.H2 + half * R * ldh;
    float* dz2 = sm.DZ2t + half * R * ldh;     // [H][R]
    float* dz1 = sm.DZ1t + half * R * ldh;
    float* scr = sm.SCR + half * kHalf * R * 4;
    const float* Wh = a.params + (half ? a.L.vw : a.L.mw);
    const float* Dh = half ? sm.DVt : sm.DMUt;
    const int NH = half ? 1 : A;
    // head back-prop: dZ2 = (dHead . Whead) * (1 - h2^2)
    auto epi_h = [&](int r, int k, float s) { const float h = h2[r * ldh + k]; dz2[k * R + r] = s * (1.f - h * h); };
    if (vecH) gemm_yW<R, 4>(Wh, Dh, NH, H, scr, t, kHalf, epi_h);
    else gemm_yW<R, 1>(Wh, Dh, NH, H, scr, t, kHalf, epi_h);
    // head parameter gradients
    {
        float* gW = gout + (half ? a.L.vw : a.L.mw);
        float* gB = gout + (half ? a.L.vb : a.L.mb);
        if (vecH) outer_store<R, 4>(gW, Dh, h2, ldh, NH, H, t, kHalf, acc);
        else outer_store<R, 1>(gW, Dh, h2, ldh, NH, H, t, kHalf, acc);
        rowsum_store<R>(gB, Dh, NH, t, kHalf, acc);
        if (!half) rowsum_store<R>(gout + a.L.ls, sm.DLSt, A, t, kHalf, acc);
    }
    // layer 2 back-prop: dZ1 = (dZ2 . W2) * (1 - h1^2)
    const float* W2 = a.params + (half ? a.L.cw2 : a.L.aw2);
    auto epi_2 = [&](int r, int k, float s) { const float h = h1[r * ldh + k]; dz1[k * R + r] = s * (1.f - h * h); };
    if (vecH) gemm_yW<R, 4>(W2, dz2, H, H, scr, t, kHalf, epi_2);
    else gemm_yW<R, 1>(W2, dz2, H, H, scr, t, kHalf, epi_2);
    {
        float* gW2 = gout + (half ? a.L.cw2 : a.L.aw2);
        float* gB2 = gout + (half ? a.L.cb2 : a.L.ab2);
        if (vecH) outer_store<R, 4>(gW2, dz2, h1, ldh, H, H, t, kHalf, acc);
        else outer_store<R, 1>(gW2, dz2, h1, ldh, H, H, t, kHalf, acc);
        rowsum_store<R>(gB2, dz2, H, t, kHalf, acc);
        float* gW1 = gout + (half ? a.L.cw1 : a.L.aw1);
        float* gB1 = gout + (half ? a.L.cb1 : a.L.ab1);
        if (vecO) outer_store<R, 4>(gW1, dz1, T.X, T.ldo, H, O, t, kHalf, acc);
        else outer_store<R, 1>(gW1, dz1, T.X, T.ldo, H, O, t, kHalf, acc);
        rowsum_store<R>(gB1, dz1, H, t, kHalf, acc);
    }
    __syncthreads();   // smem is reused by the next tile
}

// ---- phase 2: deterministic reduction of the per-CTA partial gradients ----------------------------
__device__ void ppo_reduce(const PpoArgs& a, int cta, int ncta) {
    const int tid = threadIdx.x;
    for (int p = cta * kStepThreads + tid; p < a.P; p += ncta * kStepThreads) {
        float g = 0.f;
        for (int c = 0; c < a.nslots; ++c) g += ld_cg(a.gpart + (size_t)c * a.P + p);
        __stcg(a.grad + p, g);
    }
    if (cta == 0 && tid < 2) {
        float s = 0.f;
        for (int c = 0; c < a.nslots; ++c) s += ld_cg(a.losspart + c * 4 + tid);
        __stcg(a.grad + a.P + tid, s);
    }
    if (cta == 0 && tid == 2) a.scal[0] = gaussian_entropy(a.params + a.L.ls, a.A);   // before Adam touches logstd
}

// ---- phase 3: clip_grad_norm_ + Adam ---------------------------------------------------------------
__device__ void ppo_adam(const PpoArgs& a, int step, int cta, int ncta, double* red) {
    const int tid = threadIdx.x;
    // every CTA forms the same global norm in the same order (identical clip factor everywhere)
    double s = 0.0;
    for (int p = tid; p < a.P; p += kStepThreads) {
        float g = ld_cg(a.grad + p);
        if (p >= a.L.ls && p < a.L.ls + a.A) g -= a.c_e;      // d(-c_e * entropy)/d logstd = -c_e
        s += (double)g * (double)g;
    }
    s = warp_sum(s);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < kStepThreads / 32; ++w) tot += red[w];
    const float norm = (float)sqrt(tot);
    float clip = a.max_norm / (norm + 1e-6f);
    if (clip > 1.f) clip = 1.f;
    const float ss = a.step_size[step], bc2 = a.bc2_sqrt[step];
    for (int p = cta * kStepThreads + tid; p < a.P; p += ncta * kStepThreads) {
        float g = ld_cg(a.grad + p);
        if (p >= a.L.ls && p < a.L.ls + a.A) g -= a.c_e;
        g *= clip;
        float pv = a.params[p], mv = a.m[p], vv = a.v[p];
        adam_update(pv, mv, vv, g, a.one_minus_b1, a.b2, a.one_minus_b2, ss, bc2, a.eps);
        a.params[p] = pv; a.m[p] = mv; a.v[p] = vv;
    }
    if (cta == 0 && tid == 0) {
        const float invB = 1.f / (float)a.mbs;
        float* tr = a.trace + (size_t)step * 4;
        tr[0] = ld_cg(a.grad + a.P) * invB;
        tr[1] = ld_cg(a.grad + a.P + 1) * invB;
        tr[2] = a.scal[0];
        tr[3] = norm;
    }
    __syncthreads();
}

template <int R>
__device__ __forceinline__ void ppo_phase1_all(const PpoArgs& a, int step, int cta, int ncta, float* smem) {
    PpoSmem<R> sm;
    sm.carve(smem, a.O, a.H, a.A);
    bool acc = false;
    for (int tile = cta; tile < a.ntiles; tile += ncta) {
        ppo_tile<R>(a, step, tile, a.gpart + (size_t)cta * a.P, a.losspart + cta * 4, sm, acc);
        acc = true;
    }
}

template <int R>
__global__ void __launch_bounds__(kStepThreads, 1) ppo_persistent_kernel(PpoArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ double red[kStepThreads / 32];
    GridBarrier gb{a.bar, a.bar + 1, gridDim.x, 0};
    for (int step = 0; step < a.nsteps; ++step) {
        ppo_phase1_all<R>(a, step, blockIdx.x, gridDim.x, smem);
        gb.sync();
        ppo_reduce(a, blockIdx.x, gridDim.x);
        gb.sync();
        ppo_adam(a, step, blockIdx.x, gridDim.x, red);
        gb.sync();
    }
    // a timed-out grid barrier poisons the trace so the host raises instead of trusting the result
    if (blockIdx.x == 0 && threadIdx.x == 0 && *(volatile unsigned int*)(a.bar + 1) != 0u) a.trace[0] = __int_as_float(0x7fc00000);
}

template <int R>
__global__ void __launch_bounds__(kStepThreads, 1) ppo_phase1_kernel(PpoArgs a, int step) {
    extern __shared__ __align__(16) float smem[];
    ppo_phase1_all<R>(a, step, blockIdx.x, gridDim.x, smem);
}
__global__ void __launch_bounds__(kStepThreads) ppo_phase2_kernel(PpoArgs a) { ppo_reduce(a, blockIdx.x, gridDim.x); }
__global__ void __launch_bounds__(kStepThreads) ppo_phase3_kernel(PpoArgs a, int step) {
    __shared__ double red[kStepThreads / 32];
    ppo_adam(a, step, blockIdx.x, gridDim.x, red);
}

static int ppo_tiles(const sg_ppo_config* c) { return (c->row_end - c->row_begin + kRows - 1) / kRows; }

// number of CTAs of the tile phase == number of partial-gradient slots
static int ppo_grid(const sg_ppo_config* c, int* sm_count_out) {
    int sms = sg_device_sm_count();
    if (sms <= 0) sms = 148;
    if (sm_count_out) *sm_count_out = sms;
    int tiles = ppo_tiles(c);
    int g = tiles < sms ? tiles : sms;
    return g < 1 ? 1 : g;
}

static int ppo_validate(const sg_ppo_config* c) {
    SG_REQUIRE(c, "sg_ppo: null config");
    SG_REQUIRE(c->obs_dim > 0 && c->hidden > 0 && c->act_dim > 0, "sg_ppo: non-positive model dims");
    SG_REQUIRE(c->T > 0 && c->N > 0 && c->ppo_epoch > 0 && c->num_mini_batch > 0, "sg_ppo: non-positive sizes");
    SG_REQUIRE(c->mini_batch_size > 0 && (long long)c->mini_batch_size * c->num_mini_batch <= (long long)c->T * c->N,
               "sg_ppo: mini_batch_size*num_mini_batch exceeds T*N");
    SG_REQUIRE(c->row_begin >= 0 && c->row_begin < c->row_end && c->row_end <= c->mini_batch_size,
               "sg_ppo: shard [%d,%d) outside minibatch of %d rows", c->row_begin, c->row_end, c->mini_batch_size);
    SG_REQUIRE(c->first_adam_step >= 1, "sg_ppo: first_adam_step is 1-based");
    const size_t smem = (size_t)PpoSmem<kRows>::floats(c->obs_dim, c->hidden, c->act_dim) * sizeof(float);
    SG_REQUIRE(smem <= 220 * 1024, "sg_ppo: tile needs %zu bytes of shared memory (hidden too large)", smem);
    return SG_OK;
}

struct PpoWs {
    size_t gpart, grad, losspart, scal, bar, total;
};
static PpoWs ppo_ws(const sg_ppo_config* c, int grid) {
    PolicyLayout L = make_policy_layout(c->obs_dim, c->hidden, c->act_dim);
    PpoWs w;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 255) / 256 * 256; return at; };
    w.gpart = take((size_t)grid * L.total * sizeof(float));
    w.grad = take((size_t)(L.total + 4) * sizeof(float));
    w.losspart = take((size_t)grid * 4 * sizeof(float));
    w.scal = take(4 * sizeof(float));
    w.bar = take(2 * sizeof(unsigned int));
    w.total = o;
    return w;
}

}  // namespace sg

using namespace sg;

extern "C" {
#pragma GCC visibility push(default)

int64_t sg_ppo_workspace_bytes(const sg_ppo_config* cfg) {
    if (ppo_validate(cfg)) return -1;
    return (int64_t)ppo_ws(cfg, ppo_grid(cfg, nullptr)).total;
}

int sg_ppo_update(const sg_ppo_config* cfg, float* params, float* adam_m, float* adam_v, const float* obs,
                  const float* actions, const float* value_preds, const float* returns, const float* old_logp,
                  const float* adv_stats, const int32_t* perm, const float* step_size, const float* bc2_sqrt,
                  float* trace, void* workspace, sg_allreduce_fn allreduce_cb, void* allreduce_user, void* stream) {
    int rc = ppo_validate(cfg);
    if (rc) return rc;
    SG_REQUIRE(params && adam_m && adam_v && obs && actions && value_preds && returns && old_logp && adv_stats && perm &&
                   step_size && bc2_sqrt && trace && workspace, "sg_ppo_update: null pointer");
    SG_REQUIRE(!(allreduce_cb && cfg->mode == 0), "sg_ppo_update: the allreduce callback needs mode 1");
    cudaStream_t s = (cudaStream_t)stream;
    int sms = 0;
    const int grid = ppo_grid(cfg, &sms);
    const PpoWs w = ppo_ws(cfg, grid);
    char* ws = (char*)workspace;

    PpoArgs a;
    a.O = cfg->obs_dim; a.H = cfg->hidden; a.A = cfg->act_dim;
    a.S = cfg->T * cfg->N;
    a.L = make_policy_layout(a.O, a.H, a.A);
    a.P = a.L.total;
    a.nmb = cfg->num_mini_batch; a.mbs = cfg->mini_batch_size;
    a.nsteps = cfg->ppo_epoch * cfg->num_mini_batch;
    a.row_begin = cfg->row_begin; a.row_end = cfg->row_end;
    a.ntiles = ppo_tiles(cfg);
    a.nslots = grid < a.ntiles ? grid : a.ntiles;
    a.clipped_vloss = cfg->use_clipped_value_loss;
    a.first_adam_step = cfg->first_adam_step;
    a.clip = (float)cfg->clip_param;
    a.ratio_lo = (float)(1.0 - cfg->clip_param);
    a.ratio_hi = (float)(1.0 + cfg->clip_param);
    a.c_v = (float)cfg->value_loss_coef; a.c_e = (float)cfg->entropy_coef; a.max_norm = (float)cfg->max_grad_norm;
    a.one_minus_b1 = (float)(1.0 - cfg->beta1); a.b2 = (float)cfg->beta2; a.one_minus_b2 = (float)(1.0 - cfg->beta2);
    a.eps = (float)cfg->adam_eps;
    a.params = params; a.m = adam_m; a.v = adam_v;
    a.obs = obs; a.actions = actions; a.vpred = value_preds; a.ret = returns; a.oldlp = old_logp; a.advstats = adv_stats;
    a.perm = perm; a.step_size = step_size; a.bc2_sqrt = bc2_sqrt; a.trace = trace;
    a.gpart = (float*)(ws + w.gpart); a.grad = (float*)(ws + w.grad); a.losspart = (float*)(ws + w.losspart);
    a.scal = (float*)(ws + w.scal); a.bar = (unsigned int*)(ws + w.bar);

    const size_t smem = (size_t)PpoSmem<kRows>::floats(a.O, a.H, a.A) * sizeof(float);
    // partial-gradient padding lanes must be zero; barrier words must be zero
    SG_CUDA(cudaMemsetAsync(ws, 0, w.total, s));

    if (cfg->mode == 0) {
        SG_CUDA(cudaFuncSetAttribute(ppo_persistent_kernel<kRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        SG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ppo_persistent_kernel<kRows>, kStepThreads, smem));
        SG_REQUIRE(per_sm >= 1 && grid <= per_sm * sms, "sg_ppo_update: cooperative grid of %d CTAs does not fit", grid);
        void* kargs[] = {(void*)&a};
        SG_CUDA(cudaLaunchCooperativeKernel((const void*)ppo_persistent_kernel<kRows>, dim3(grid), dim3(kStepThreads), kargs, smem, s));
        count_launches(1);
    } else {
        SG_CUDA(cudaFuncSetAttribute(ppo_phase1_kernel<kRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int g2 = (a.P + kStepThreads - 1) / kStepThreads;
        for (int step = 0; step < a.nsteps; ++step) {
            ppo_phase1_kernel<kRows><<<grid, kStepThreads, smem, s>>>(a, step);
            ppo_phase2_kernel<<<g2, kStepThreads, 0, s>>>(a);
            if (allreduce_cb) {
                int cb = allreduce_cb(a.grad, a.P + 2, allreduce_user);
                SG_REQUIRE(cb == 0, "sg_ppo_update: allreduce callback failed with %d at step %d", cb, step);
            }
            ppo_phase3_kernel<<<g2, kStepThreads, 0, s>>>(a, step);
            count_launches(3);
        }
        SG_CUDA(cudaGetLastError());
    }
    return SG_OK;
}

#pragma GCC visibility pop
}  // extern "C"
