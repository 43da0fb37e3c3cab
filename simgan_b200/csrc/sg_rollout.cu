// Rollout-buffer kernels: return/GAE scan, advantage statistics, sampler row gathers, fused insert,
// and the batched actor-critic forward used by Policy.act / get_value / evaluate_actions.
#include "sg_common.cuh"
#include "sg_policy.cuh"

namespace sg {

// ------------------------------------------------------------------------------------------------
// compute_returns (A2C/storage.py:103-142): a first-order linear recurrence over t = T-1..0 per env column.
// Every arithmetic step uses explicit round-to-nearest mul/add in the reference's order so the result is
// bit-identical to the eager fp32 tensor ops (no FMA contraction).
//   mode 0: GAE + proper time limits   mode 1: GAE   mode 2: plain + proper limits   mode 3: plain
// A thread walking the chain cannot hide the latency of its own loads (16 columns x 2048 dependent steps took
// 0.46 ms with thread-level prefetch), so the inputs are staged:
// one CTA per 32 columns; ALL 256 threads stream chunks of kTC steps of the
// four input arrays into a kScanStages-deep ring of shared-memory buffers with cp.async (coalesced 128-byte row
// segments) while the first `cols` threads walk the recurrence out of shared memory; results are written back
// coalesced.  Three chunks in flight (~1.5 us of chain work) cover the global-memory latency of a chunk; with two
// buffers of 32 steps every chunk still waited ~1 us for its loads.
// ------------------------------------------------------------------------------------------------
constexpr int kTC = 64;
constexpr int kScanStages = 4;
constexpr size_t kScanSmemBytes = (size_t)(4 * kScanStages + 1) * kTC * 32 * sizeof(float);

__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<unsigned long long>(p) & 15ull) == 0; }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NKEEP>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NKEEP) : "memory"); }

template <int MODE>
__global__ void __launch_bounds__(256) returns_scan_staged_kernel(const float* __restrict__ rewards, float* __restrict__ vpred,
                                                                  const float* __restrict__ masks,
                                                                  const float* __restrict__ bad, float* __restrict__ ret,
                                                                  const float* __restrict__ next_value, int T, int N,
                                                                  float g, float gl) {
    constexpr bool GAE = (MODE <= 1), PROPER = (MODE == 0 || MODE == 2);
    extern __shared__ __align__(16) float scan_smem[];
    typedef float (*Stage)[kTC][32];
    Stage sR = reinterpret_cast<Stage>(scan_smem);
    Stage sV = sR + kScanStages, sM = sV + kScanStages, sB = sM + kScanStages;
    float (*sO)[32] = reinterpret_cast<float (*)[32]>(sB + kScanStages);
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * 32;
    const int cols = min(32, N - n0);
    const int nchunks = (T + kTC - 1) / kTC;
    // chunk k covers steps t in [T - (k+1)*kTC, T - k*kTC) (clipped at 0); slot tt <-> t = tbase + tt
    // 16-byte copies when every row segment is float4-aligned (N % 4 == 0 and aligned bases): 8 float4 per row,
    // 32 rows per sweep of the CTA, one 64-bit index per copy -- the scalar form spent more time computing
    // addresses for 4-byte copies than the chain spends on the recurrence
    const bool vec = (N & 3) == 0 && aligned16(rewards) && aligned16(vpred) && aligned16(masks) && (!PROPER || aligned16(bad));
    auto issue = [&](int k) {       // always commits a (possibly empty) group, so that group k <-> chunk k
        const int b = k % kScanStages;
        const int tbase = T - (k + 1) * kTC;
        if (k < nchunks && vec) {
            const int c4 = 4 * (tid & 7);
            if (c4 < cols) {
                for (int tt = tid >> 3; tt < kTC; tt += 32) {
                    const int t = tbase + tt;
                    if (t >= 0) {
                        const size_t i0 = (size_t)t * N + n0 + c4, i1 = i0 + N;
                        cp_async16(&sR[b][tt][c4], rewards + i0);
                        cp_async16(&sV[b][tt][c4], vpred + i0);
                        cp_async16(&sM[b][tt][c4], masks + i1);
                        if (PROPER) cp_async16(&sB[b][tt][c4], bad + i1);
                    }
                }
            }
        } else if (k < nchunks) {
            for (int e = tid; e < kTC * 32; e += 256) {
                const int tt = e >> 5, c = e & 31;
                const int t = tbase + tt;
                if (c < cols && t >= 0) {
                    const size_t i0 = (size_t)t * N + n0 + c, i1 = (size_t)(t + 1) * N + n0 + c;
                    cp_async4(&sR[b][tt][c], rewards + i0);
                    cp_async4(&sV[b][tt][c], vpred + i0);
                    cp_async4(&sM[b][tt][c], masks + i1);
                    if (PROPER) cp_async4(&sB[b][tt][c], bad + i1);
                }
            }
        }
        cp_async_commit();
    };
    float carry = 0.f, v_next = 0.f;
    if (tid < cols) {
        const float nv = next_value[n0 + tid];
        v_next = nv;
        if (GAE) { vpred[(size_t)T * N + n0 + tid] = nv; carry = 0.f; }
        else { ret[(size_t)T * N + n0 + tid] = nv; carry = nv; }
    }
    for (int k = 0; k < kScanStages - 1; ++k) issue(k);
    for (int k = 0; k < nchunks; ++k) {
        issue(k + kScanStages - 1);       // refills the buffer chunk k-1 was read from (all threads passed its barrier)
        cp_async_wait<kScanStages - 1>();
        __syncthreads();
        const int b = k % kScanStages;
        const int tbase = T - (k + 1) * kTC;
        if (tid < cols) {
            const int c = tid;
            // one step of the recurrence in the reference's op order (storage.py:109-142); everything that does not
            // depend on `carry` (delta, gl*m) is off the dependent chain
            auto step = [&](float r, float v, float m, float bb) {
                float out;
                if (GAE) {
                    const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(g, v_next), m)), v);
                    carry = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, m), carry));
                    if (PROPER) carry = __fmul_rn(carry, bb);
                    out = __fadd_rn(carry, v);
                    v_next = v;
                } else if (PROPER) {
                    const float a = __fmul_rn(__fadd_rn(__fmul_rn(__fmul_rn(carry, g), m), r), bb);
                    out = __fadd_rn(a, __fmul_rn(__fsub_rn(1.f, bb), v));
                    carry = out;
                } else {
                    out = __fadd_rn(__fmul_rn(__fmul_rn(carry, g), m), r);
                    carry = out;
                }
                return out;
            };
            const int lo = tbase < 0 ? -tbase : 0;      // first valid slot of this chunk
            int tt = kTC - 1;
            // groups of 8 steps: all shared-memory loads of a group are issued before its dependent chain starts
            for (; tt - 7 >= lo; tt -= 8) {
                float r[8], v[8], m[8], bb[8], o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    r[j] = sR[b][tt - j][c]; v[j] = sV[b][tt - j][c]; m[j] = sM[b][tt - j][c];
                    bb[j] = PROPER ? sB[b][tt - j][c] : 1.f;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = step(r[j], v[j], m[j], bb[j]);
#pragma unroll
                for (int j = 0; j < 8; ++j) sO[tt - j][c] = o[j];
            }
            for (; tt >= lo; --tt) sO[tt][c] = step(sR[b][tt][c], sV[b][tt][c], sM[b][tt][c], PROPER ? sB[b][tt][c] : 1.f);
        }
        __syncthreads();
        for (int e = tid; e < kTC * 32; e += 256) {
            const int tt = e >> 5, c = e & 31;
            const int t = tbase + tt;
            if (c < cols && t >= 0) ret[(size_t)t * N + n0 + c] = sO[tt][c];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// advantage statistics: mean and unbiased std of (returns - value_preds)[:S]  (A2C/algo/ppo.py:66-68)
// fp64 accumulation; stage 1 = per-CTA partial (sum, sumsq), stage 2 = one CTA finalises.
// ------------------------------------------------------------------------------------------------
constexpr int kStatBlocks = 296;

__global__ void __launch_bounds__(256) adv_partial_kernel(const float* __restrict__ ret, const float* __restrict__ vp,
                                                          int S, double* __restrict__ part) {
    double s = 0.0, q = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S; i += gridDim.x * blockDim.x) {
        const double a = (double)__fsub_rn(ret[i], vp[i]);
        s += a; q += a * a;
    }
    __shared__ double sh[2][8];
    s = warp_sum(s); q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0, tq = 0;
        for (int w = 0; w < 8; ++w) { ts += sh[0][w]; tq += sh[1][w]; }
        part[2 * blockIdx.x] = ts; part[2 * blockIdx.x + 1] = tq;
    }
}

__global__ void adv_final_kernel(const double* __restrict__ part, int nblocks, int S, float* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0, q = 0;
        for (int b = 0; b < nblocks; ++b) { s += part[2 * b]; q += part[2 * b + 1]; }
        const double mean = s / S;
        double var = (q - s * mean) / (S > 1 ? (S - 1) : 1);
        if (var < 0) var = 0;
        out[0] = (float)mean;
        out[1] = (S > 1) ? (float)sqrt(var) : __int_as_float(0x7fc00000);  // torch.std of 1 element = nan
    }
}

// ------------------------------------------------------------------------------------------------
// sampler row gathers / block copies
// ------------------------------------------------------------------------------------------------
constexpr int kMaxPtrs = 12;
struct GatherArgs {
    const float* src[kMaxPtrs];
    float* dst[kMaxPtrs];
    int dim[kMaxPtrs];
    int n;
};

__global__ void __launch_bounds__(256) gather_rows_kernel(GatherArgs a, const int64_t* __restrict__ idx, int n_rows) {
    const int i = blockIdx.y;
    const int d = a.dim[i];
    const long long total = (long long)n_rows * d;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(e / d), c = (int)(e - (long long)j * d);
        a.dst[i][e] = a.src[i][(size_t)idx[j] * d + c];
    }
}

__global__ void __launch_bounds__(256) copy_blocks_kernel(GatherArgs a) {
    const int i = blockIdx.y;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < a.dim[i]; e += gridDim.x * blockDim.x) a.dst[i][e] = a.src[i][e];
}

// ------------------------------------------------------------------------------------------------
// Policy forward (A2C/model.py:89-114)
// ------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(kStepThreads) policy_forward_kernel(const float* __restrict__ params, PolicyLayout L,
                                                                      int O, int H, int A, const float* __restrict__ obs,
                                                                      int B, const float* __restrict__ noise,
                                                                      const float* __restrict__ actions_in,
                                                                      float* __restrict__ value, float* __restrict__ action,
                                                                      float* __restrict__ logp, float* __restrict__ entropy) {
    extern __shared__ __align__(16) float smem[];
    PolicyTile<R> T;
    T.carve(smem, O, H, A);
    const int tid = threadIdx.x;
    for (int tile = blockIdx.x; tile * R < B; tile += gridDim.x) {
        const int row0 = tile * R;
        for (int e = tid; e < R * T.ldo; e += kStepThreads) {
            const int r = e / T.ldo, k = e - r * T.ldo;
            T.X[e] = (row0 + r < B && k < O) ? obs[(size_t)(row0 + r) * O + k] : 0.f;
        }
        __syncthreads();
        policy_tile_forward<R>(params, L, O, H, A, T, tid);
        // per (row, action): sample / copy the action
        for (int e = tid; e < R * A; e += kStepThreads) {
            const int r = e / A, a = e - r * A;
            const int row = row0 + r;
            float act = 0.f;
            if (row < B) {
                const float mu = T.MU[r * T.lda + a];
                if (actions_in) act = actions_in[(size_t)row * A + a];
                else if (noise) act = __fadd_rn(__fmul_rn(noise[(size_t)row * A + a], expf(ld_cg(params + L.ls + a))), mu);
                else act = mu;
                if (action) action[(size_t)row * A + a] = act;
            }
            T.ACT[r * T.lda + a] = act;
        }
        __syncthreads();
        if (tid < R && row0 + tid < B) {
            const int row = row0 + tid;
            if (value) value[row] = T.VAL[tid];
            if (logp) logp[row] = gaussian_logp_row(T.MU + tid * T.lda, T.ACT + tid * T.lda, params + L.ls, A);
        }
        if (entropy && blockIdx.x == 0 && tile == 0 && tid == 0) entropy[0] = gaussian_entropy(params + L.ls, A);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Rollout feed step (A2C/main_gail_dyn_ppo.py:209-236 + A2C/envs.py:199-210 + A2C/storage.py:70-84), one launch
// per environment step:
//   insert   the staged env outputs of step `step` (obs', sas_feat, reward, mask, bad_mask; one packed block that
//            arrived with ONE async H2D copy) go to slot step+1 (obs, obs_feat, masks, bad_masks, hxs) and slot
//            step (rewards); actions / log-probs / values of slot `step` were written when they were computed.
//   act      Policy.act on the new observations -> value_preds[step+1], actions[step+1], action_log_probs[step+1]
//            (when step+1 < T) and the action block the host envs read back with ONE async D2H copy.
// step = -1 is the start of a rollout: no insert, act on obs[0].
// ------------------------------------------------------------------------------------------------
struct FeedArgs {
    const float* params;
    PolicyLayout L;
    int O, H, A, F, N, T, step;
    const float *staged, *noise;
    float *obs, *obs_feat, *hxs, *rewards, *value_preds, *action_log_probs, *actions, *masks, *bad_masks, *action_out;
};

template <int R>
__global__ void __launch_bounds__(kStepThreads) rollout_feed_kernel(FeedArgs a) {
    extern __shared__ __align__(16) float smem[];
    PolicyTile<R> T;
    T.carve(smem, a.O, a.H, a.A);
    const int tid = threadIdx.x;
    const int O = a.O, A = a.A, F = a.F, N = a.N;
    const int slot = a.step + 1;                         // slot the new observations go to / are acted on
    const float* st_obs = a.staged;                      // [N*O | N*F | N | N | N]
    const float* st_feat = st_obs + (size_t)N * O;
    const float* st_rew = st_feat + (size_t)N * F;
    const float* st_mask = st_rew + N;
    const float* st_bad = st_mask + N;
    const bool do_act = slot < a.T || a.step < 0 || true;    // value of obs[T] doubles as next_value
    for (int tile = blockIdx.x; tile * R < N; tile += gridDim.x) {
        const int row0 = tile * R;
        const float* src_obs = a.step >= 0 ? st_obs : a.obs;       // step -1: act on obs[0] already in the buffer
        for (int e = tid; e < R * T.ldo; e += kStepThreads) {
            const int r = e / T.ldo, k = e - r * T.ldo;
            const int row = row0 + r;
            float x = 0.f;
            if (row < N && k < O) {
                x = src_obs[(size_t)row * O + k];
                if (a.step >= 0) a.obs[((size_t)slot * N + row) * O + k] = x;
            }
            T.X[e] = x;
        }
        if (a.step >= 0) {
            for (int e = tid; e < R * F; e += kStepThreads) {
                const int r = e / F, k = e - r * F;
                const int row = row0 + r;
                if (row < N) a.obs_feat[((size_t)slot * N + row) * F + k] = st_feat[(size_t)row * F + k];
            }
            if (tid < R && row0 + tid < N) {
                const int row = row0 + tid;
                a.rewards[(size_t)a.step * N + row] = st_rew[row];
                a.masks[(size_t)slot * N + row] = st_mask[row];
                a.bad_masks[(size_t)slot * N + row] = st_bad[row];
                a.hxs[(size_t)slot * N + row] = 0.f;                   // non-recurrent policies carry zeros (storage.py:78)
            }
        }
        __syncthreads();
        if (do_act) {
            policy_tile_forward<R>(a.params, a.L, O, a.H, A, T, tid);
            for (int e = tid; e < R * A; e += kStepThreads) {
                const int r = e / A, k = e - r * A;
                const int row = row0 + r;
                float act = 0.f;
                if (row < N) {
                    const float mu = T.MU[r * T.lda + k];
                    act = a.noise ? __fadd_rn(__fmul_rn(a.noise[(size_t)row * A + k], expf(ld_cg(a.params + a.L.ls + k))), mu) : mu;
                    a.action_out[(size_t)row * A + k] = act;
                    if (slot < a.T) a.actions[((size_t)slot * N + row) * A + k] = act;
                }
                T.ACT[r * T.lda + k] = act;
            }
            __syncthreads();
            if (tid < R && row0 + tid < N) {
                const int row = row0 + tid;
                a.value_preds[(size_t)slot * N + row] = T.VAL[tid];
                if (slot < a.T)
                    a.action_log_probs[(size_t)slot * N + row] = gaussian_logp_row(T.MU + tid * T.lda, T.ACT + tid * T.lda, a.params + a.L.ls, A);
            }
        }
        __syncthreads();
    }
}

}  // namespace sg

using namespace sg;

extern "C" {
#pragma GCC visibility push(default)

int sg_compute_returns(const float* rewards, float* value_preds, const float* masks, const float* bad_masks,
                       float* returns, const float* next_value, int T, int N, double gamma, double gae_lambda,
                       int use_gae, int use_proper_time_limits, void* stream) {
    SG_REQUIRE(T > 0 && N > 0, "sg_compute_returns: T and N must be positive (T=%d N=%d)", T, N);
    SG_REQUIRE(rewards && value_preds && masks && bad_masks && returns && next_value, "sg_compute_returns: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const float g = (float)gamma, gl = (float)(gamma * gae_lambda);
    const int mode = use_gae ? (use_proper_time_limits ? 0 : 1) : (use_proper_time_limits ? 2 : 3);
    const int sb = (N + 31) / 32;
    static SmemGrant grants[4];
    {
        int rc = mode == 0 ? grant_smem(grants[0], returns_scan_staged_kernel<0>, kScanSmemBytes)
               : mode == 1 ? grant_smem(grants[1], returns_scan_staged_kernel<1>, kScanSmemBytes)
               : mode == 2 ? grant_smem(grants[2], returns_scan_staged_kernel<2>, kScanSmemBytes)
                           : grant_smem(grants[3], returns_scan_staged_kernel<3>, kScanSmemBytes);
        if (rc) return rc;
    }
    switch (mode) {
        case 0: returns_scan_staged_kernel<0><<<sb, 256, kScanSmemBytes, s>>>(rewards, value_preds, masks, bad_masks, returns, next_value, T, N, g, gl); break;
        case 1: returns_scan_staged_kernel<1><<<sb, 256, kScanSmemBytes, s>>>(rewards, value_preds, masks, bad_masks, returns, next_value, T, N, g, gl); break;
        case 2: returns_scan_staged_kernel<2><<<sb, 256, kScanSmemBytes, s>>>(rewards, value_preds, masks, bad_masks, returns, next_value, T, N, g, gl); break;
        default: returns_scan_staged_kernel<3><<<sb, 256, kScanSmemBytes, s>>>(rewards, value_preds, masks, bad_masks, returns, next_value, T, N, g, gl); break;
    }
    count_launches(1);
    SG_CUDA(cudaGetLastError());
    return SG_OK;
}

int64_t sg_adv_stats_workspace_bytes(int S) { (void)S; return (int64_t)kStatBlocks * 2 * sizeof(double); }

int sg_adv_stats(const float* returns, const float* value_preds, int S, float* out_stats, void* workspace, void* stream) {
    SG_REQUIRE(S > 0 && returns && value_preds && out_stats && workspace, "sg_adv_stats: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    int blocks = (S + 255) / 256;
    if (blocks > kStatBlocks) blocks = kStatBlocks;
    adv_partial_kernel<<<blocks, 256, 0, s>>>(returns, value_preds, S, (double*)workspace);
    adv_final_kernel<<<1, 32, 0, s>>>((const double*)workspace, blocks, S, out_stats);
    count_launches(2);
    SG_CUDA(cudaGetLastError());
    return SG_OK;
}

int sg_gather_rows(const float* const* h_src, float* const* h_dst, const int* h_dims, int n_tensors,
                   const int64_t* idx, int n_rows, void* stream) {
    SG_REQUIRE(n_tensors > 0 && n_tensors <= kMaxPtrs, "sg_gather_rows: n_tensors must be in 1..%d", kMaxPtrs);
    if (n_rows == 0) return SG_OK;
    SG_REQUIRE(n_rows > 0 && idx, "sg_gather_rows: bad index array");
    GatherArgs a;
    a.n = n_tensors;
    int maxd = 1;
    for (int i = 0; i < n_tensors; ++i) {
        SG_REQUIRE(h_src[i] && h_dst[i] && h_dims[i] > 0, "sg_gather_rows: tensor %d invalid", i);
        a.src[i] = h_src[i]; a.dst[i] = h_dst[i]; a.dim[i] = h_dims[i];
        if (h_dims[i] > maxd) maxd = h_dims[i];
    }
    long long total = (long long)n_rows * maxd;
    int bx = (int)((total + 255) / 256);
    if (bx > 1184) bx = 1184;
    gather_rows_kernel<<<dim3(bx, n_tensors), 256, 0, (cudaStream_t)stream>>>(a, idx, n_rows);
    count_launches(1);
    SG_CUDA(cudaGetLastError());
    return SG_OK;
}

int sg_copy_blocks(const float* const* h_src, float* const* h_dst, const int* h_count, int n_copies, void* stream) {
    SG_REQUIRE(n_copies > 0 && n_copies <= kMaxPtrs, "sg_copy_blocks: n_copies must be in 1..%d", kMaxPtrs);
    GatherArgs a;
    a.n = n_copies;
    int maxc = 1;
    for (int i = 0; i < n_copies; ++i) {
        SG_REQUIRE(h_src[i] && h_dst[i] && h_count[i] >= 0, "sg_copy_blocks: block %d invalid", i);
        a.src[i] = h_src[i]; a.dst[i] = h_dst[i]; a.dim[i] = h_count[i];
        if (h_count[i] > maxc) maxc = h_count[i];
    }
    int bx = (maxc + 255) / 256;
    if (bx > 592) bx = 592;
    copy_blocks_kernel<<<dim3(bx, n_copies), 256, 0, (cudaStream_t)stream>>>(a);
    count_launches(1);
    SG_CUDA(cudaGetLastError());
    return SG_OK;
}

int sg_policy_forward(const float* params, int obs_dim, int hidden, int act_dim, const float* obs, int B,
                      const float* noise, const float* actions_in, float* value, float* action, float* logp,
                      float* entropy, void* stream) {
    SG_REQUIRE(params && obs && B > 0 && obs_dim > 0 && hidden > 0 && act_dim > 0, "sg_policy_forward: bad arguments");
    PolicyLayout L = make_policy_layout(obs_dim, hidden, act_dim);
    const size_t smem = (size_t)PolicyTile<kRows>::floats(obs_dim, hidden, act_dim) * sizeof(float);
    SG_REQUIRE(smem <= 200 * 1024, "sg_policy_forward: tile needs %zu bytes of shared memory", smem);
    static SmemGrant grant;
    if (int rc = grant_smem(grant, policy_forward_kernel<kRows>, smem)) return rc;
    int tiles = (B + kRows - 1) / kRows;
    int grid = tiles < 1184 ? tiles : 1184;
    policy_forward_kernel<kRows><<<grid, kStepThreads, smem, (cudaStream_t)stream>>>(params, L, obs_dim, hidden, act_dim, obs, B,
                                                                                     noise, actions_in, value, action, logp, entropy);
    count_launches(1);
    SG_CUDA(cudaGetLastError());
    return SG_OK;
}

int64_t sg_rollout_stage_floats(int obs_dim, int feat_dim, int N) { return (int64_t)N * (obs_dim + feat_dim + 3); }

int sg_rollout_feed(const float* params, int obs_dim, int hidden, int act_dim, int feat_dim, int N, int T, int step,
                    const float* staged, const float* noise, float* obs, float* obs_feat, float* hxs, float* rewards,
                    float* value_preds, float* action_log_probs, float* actions, float* masks, float* bad_masks,
                    float* action_out, void* stream) {
    SG_REQUIRE(params && obs && obs_feat && hxs && rewards && value_preds && action_log_probs && actions && masks && bad_masks &&
                   action_out, "sg_rollout_feed: null pointer");
    SG_REQUIRE(obs_dim > 0 && hidden > 0 && act_dim > 0 && feat_dim >= 0 && N > 0 && T > 0, "sg_rollout_feed: non-positive sizes");
    SG_REQUIRE(step >= -1 && step < T, "sg_rollout_feed: step %d outside [-1,%d)", step, T);
    SG_REQUIRE(step < 0 || staged, "sg_rollout_feed: staged block required for step >= 0");
    FeedArgs a;
    a.params = params; a.L = make_policy_layout(obs_dim, hidden, act_dim);
    a.O = obs_dim; a.H = hidden; a.A = act_dim; a.F = feat_dim; a.N = N; a.T = T; a.step = step;
    a.staged = staged; a.noise = noise;
    a.obs = obs; a.obs_feat = obs_feat; a.hxs = hxs; a.rewards = rewards; a.value_preds = value_preds;
    a.action_log_probs = action_log_probs; a.actions = actions; a.masks = masks; a.bad_masks = bad_masks; a.action_out = action_out;
    const size_t smem = (size_t)PolicyTile<kRows>::floats(obs_dim, hidden, act_dim) * sizeof(float);
    SG_REQUIRE(smem <= 200 * 1024, "sg_rollout_feed: tile needs %zu bytes of shared memory", smem);
    static SmemGrant grant;
    if (int rc = grant_smem(grant, rollout_feed_kernel<kRows>, smem)) return rc;
    int tiles = (N + kRows - 1) / kRows;
    int grid = tiles < 1184 ? tiles : 1184;
    rollout_feed_kernel<kRows><<<grid, kStepThreads, smem, (cudaStream_t)stream>>>(a);
    count_launches(1);
    SG_CUDA(cudaGetLastError());
    return SG_OK;
}

#pragma GCC visibility pop
}  // extern "C"
