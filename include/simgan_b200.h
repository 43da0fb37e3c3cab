/*
 * simgan_b200 C ABI  --  the drop-in boundary under the Python classes that mirror the reference's
 * Policy / PPO / gail.Discriminator / RolloutStorage.
 *
 * The reference (jyf588/SimGAN) is pure Python and has no FFI for this path; each entry point below
 * replaces the eager-PyTorch op sequence of the cited reference function.  Citations are relative to
 * the reference root; A2C = third_party/a2c_ppo_acktr.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with h_ (host); no ownership transfer
 *   - all tensors are dense fp32 in the reference's own layouts (time-major (T,N,D) rollout buffers,
 *     nn.Linear weights (out,in) row-major); indices are int32 flat sample ids t*N+n
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it
 *   - return value: 0 on success, non-zero on error; sg_last_error() gives the message
 *   - flat parameter vectors use the segment tables returned by sg_policy_layout / sg_disc_layout
 *     (every segment starts on a 16-byte boundary)
 */
#ifndef SIMGAN_B200_H
#define SIMGAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SG_OK 0
#define SG_ERR_INVALID 1
#define SG_ERR_CUDA 2
#define SG_ERR_WORKSPACE 3

const char* sg_last_error(void);
int sg_version(void);
/* Number of CUDA kernel launches this library has issued in this process (monotonic; bench.py reads
 * it before/after the timed region to report gpu_launches). */
long long sg_launch_count(void);
/* SM count / cooperative-launch capability of the current device (0 if no device). */
int sg_device_sm_count(void);

/* ---- flat parameter layouts ---------------------------------------------------------------- */
/* Policy(MLPBase)+DiagGaussian, 13 segments in nn.Module.parameters() order
 * (A2C/model.py:243-251, A2C/distributions.py:98-104):
 *   0 actor.0.weight (H,O)  1 actor.0.bias (H)  2 actor.2.weight (H,H)  3 actor.2.bias (H)
 *   4 critic.0.weight (H,O) 5 critic.0.bias (H) 6 critic.2.weight (H,H) 7 critic.2.bias (H)
 *   8 critic_linear.weight (1,H) 9 critic_linear.bias (1)
 *  10 dist.fc_mean.weight (A,H) 11 dist.fc_mean.bias (A) 12 dist.logstd._bias (A,1)
 * offsets[13] (in floats) receives the segment starts; returns the padded total length. */
#define SG_POLICY_SEGMENTS 13
int sg_policy_layout(int obs_dim, int hidden, int act_dim, int* offsets);
/* Discriminator trunk, 6 segments (A2C/algo/gail.py:40-43):
 *   0 trunk.0.weight (Hd,F) 1 trunk.0.bias (Hd) 2 trunk.2.weight (Hd,Hd) 3 trunk.2.bias (Hd)
 *   4 trunk.4.weight (1,Hd) 5 trunk.4.bias (1) */
#define SG_DISC_SEGMENTS 6
int sg_disc_layout(int feat_dim, int hidden, int* offsets);

/* ---- rollout buffer ------------------------------------------------------------------------ */
/* RolloutStorage.compute_returns, all four branches (A2C/storage.py:103-142).
 * rewards (T,N,1); value_preds/returns/masks/bad_masks (T+1,N,1); next_value (N,1).
 * GAE branches write value_preds[T] = next_value; the others write returns[T] = next_value.
 * gamma / gae_lambda are doubles because the reference forms gamma*gae_lambda in Python double
 * before narrowing to fp32.  Bit-exact with the reference's fp32 op order. */
int sg_compute_returns(const float* rewards, float* value_preds, const float* masks, const float* bad_masks,
                       float* returns, const float* next_value, int T, int N, double gamma, double gae_lambda,
                       int use_gae, int use_proper_time_limits, void* stream);

/* mean and UNBIASED std of (returns - value_preds)[:S] (A2C/algo/ppo.py:66-68).
 * out_stats: 2 floats {mean, std}. workspace: sg_adv_stats_workspace_bytes(S) bytes. */
int64_t sg_adv_stats_workspace_bytes(int S);
int sg_adv_stats(const float* returns, const float* value_preds, int S, float* out_stats, void* workspace,
                 void* stream);

/* Row gathers of RolloutStorage.feed_forward_generator (A2C/storage.py:169-185): for each of
 * n_tensors sources, dst[i][j,:] = src[i][idx[j],:] with row width dims[i].  idx is int64 (the
 * reference's sampler indices), h_src/h_dst/h_dims are HOST arrays of length n_tensors. */
int sg_gather_rows(const float* const* h_src, float* const* h_dst, const int* h_dims, int n_tensors,
                   const int64_t* idx, int n_rows, void* stream);

/* RolloutStorage.insert (A2C/storage.py:70-84): nine row-block copies in one launch.
 * h_src/h_dst/h_count: HOST arrays (n_copies) of device pointers and float counts. */
int sg_copy_blocks(const float* const* h_src, float* const* h_dst, const int* h_count, int n_copies, void* stream);

/* ---- actor-critic -------------------------------------------------------------------------- */
/* Policy.act / get_value / evaluate_actions forward (A2C/model.py:89-114).
 * params: flat policy vector.  obs (B,O).  noise (B,A) standard normal or NULL (deterministic /
 * evaluate).  actions_in (B,A) or NULL: when given, log-probs are evaluated for it (evaluate_actions)
 * instead of the sampled action.  Outputs (any may be NULL): value (B,1), action (B,A), logp (B,1),
 * entropy (1) = batch-mean entropy. */
int sg_policy_forward(const float* params, int obs_dim, int hidden, int act_dim, const float* obs, int B,
                      const float* noise, const float* actions_in, float* value, float* action, float* logp,
                      float* entropy, void* stream);

/* Rollout feed step = the body of the collection loop (A2C/main_gail_dyn_ppo.py:209-236) in ONE launch:
 * RolloutStorage.insert (A2C/storage.py:70-84) of the staged env outputs of step `step` + Policy.act
 * (A2C/model.py:89-101) on the new observations.
 *   staged     device copy of ONE packed block [obs' (N,O) | sas_feat (N,F) | reward (N) | mask (N) | bad_mask (N)]
 *              (sg_rollout_stage_floats floats; the host fills a pinned twin and issues one async H2D copy)
 *   noise      (N,A) standard normal draws or NULL for the deterministic mode
 *   step       slot being completed (0..T-1), or -1 at the start of a rollout (no insert, act on obs[0])
 *   effects    obs/obs_feat/masks/bad_masks/hxs[step+1], rewards[step]  <- staged
 *              value_preds[step+1]; actions/action_log_probs[step+1] (when step+1 < T)  <- act(obs[step+1])
 *              action_out (N,A): the actions for the host envs (one async D2H copy) */
int64_t sg_rollout_stage_floats(int obs_dim, int feat_dim, int N);
int sg_rollout_feed(const float* params, int obs_dim, int hidden, int act_dim, int feat_dim, int N, int T, int step,
                    const float* staged, const float* noise, float* obs, float* obs_feat, float* hxs, float* rewards,
                    float* value_preds, float* action_log_probs, float* actions, float* masks, float* bad_masks,
                    float* action_out, void* stream);

/* ---- PPO ----------------------------------------------------------------------------------- */
typedef struct sg_ppo_config {
    int obs_dim, hidden, act_dim;
    int T, N;                 /* rollout dims; S = T*N */
    int ppo_epoch;            /* epochs in this call */
    int num_mini_batch;       /* minibatches per epoch */
    int mini_batch_size;      /* rows per minibatch (S / num_mini_batch) */
    double clip_param, value_loss_coef, entropy_coef, max_grad_norm;   /* Python floats of PPO.__init__ */
    double beta1, beta2, adam_eps;
    int use_clipped_value_loss;
    int first_adam_step;      /* Adam step count of the first minibatch in this call (1-based) */
    int row_begin, row_end;   /* data-parallel shard: rows [row_begin,row_end) of every minibatch are
                                 processed locally (0, mini_batch_size on a single GPU) */
    int mode;                 /* 0 = auto (resident if the parameter image fits in shared memory, else
                                 persistent), 1 = one launch per phase, 2 = persistent (weights through L2),
                                 3 = resident (weights in shared memory), 4 = tensor cores: tcgen05 kind::tf32 MMAs
                                 with 3xTF32 operand splitting over 64/128-row jobs (hidden in {64,128,256}, act_dim <= 32,
                                 obs_dim <= 256); mode 0 picks them once a minibatch fills the SMs with such jobs */
    void* dp_ctx;             /* sg_dp_create context: fused peer-memory gradient exchange inside the persistent
                                 kernel (modes 0/2/3/4; every rank must pass shards of equal size); NULL = none */
} sg_ppo_config;

int64_t sg_ppo_workspace_bytes(const sg_ppo_config* cfg);
/* Diagnostics: byte offset inside the workspace of the per-CTA, per-phase clock64 totals of the last persistent
 * launch: int64 [n_ctas][8] with slots {param image, tile phase, barrier 1, reduce+ssq, barrier 2, clip+Adam,
 * barrier 3, -}; n_ctas = min(#tiles, #SMs). */
int64_t sg_ppo_phase_cycles_offset(const sg_ppo_config* cfg);
/* Diagnostics: 1 when sg_ppo_update will run this configuration on the tensor cores (mode 4, or mode 0 with a minibatch
 * shard large enough to fill the SMs with 64/128-row jobs), 0 for the CUDA-core tiles, -1 on an invalid configuration. */
int sg_ppo_uses_tensor_cores(const sg_ppo_config* cfg);

/* PPO.update (A2C/algo/ppo.py:65-157) for cfg->ppo_epoch epochs.
 *   params/adam_m/adam_v : flat policy vectors (updated in place)
 *   obs (T+1,N,O), actions (T,N,A), value_preds (T+1,N,1), returns (T+1,N,1), old_logp (T,N,1)
 *   adv_stats {mean,std} from sg_adv_stats
 *   perm  int32 (ppo_epoch, S): the reference sampler's torch.randperm(S) per epoch (A2C/storage.py:158-162)
 *   step_size / bc2_sqrt: device float arrays (n_steps) of Adam lr/(1-beta1^t) and sqrt(1-beta2^t),
 *     formed on the host in double like torch.optim.Adam does
 *   trace (n_steps,4) {value_loss, action_loss, entropy, grad_norm} per optimizer step
 * In mode 1 with allreduce_cb != NULL the callback is invoked on the host between the gradient
 * reduction and the clip+Adam epilogue of every step with (grad_ptr, n_floats, user) and must
 * enqueue an in-place sum-allreduce on `stream` (data-parallel minibatch sharding). */
typedef int (*sg_allreduce_fn)(float* grad, int n_floats, void* user);
int sg_ppo_update(const sg_ppo_config* cfg, float* params, float* adam_m, float* adam_v, const float* obs,
                  const float* actions, const float* value_preds, const float* returns, const float* old_logp,
                  const float* adv_stats, const int32_t* perm, const float* step_size, const float* bc2_sqrt,
                  float* trace, void* workspace, sg_allreduce_fn allreduce_cb, void* allreduce_user, void* stream);

/* ---- SplitPolicy (A2C/model_split.py:39-95, 157-238; what the shipped train_*.sh scripts use) ---------------- */
/* Flat layout: offsets[22] in nn.Module.parameters() order of SplitPolicy (base.actor_contact.{0,2}.{weight,bias},
 * base.actor_actuator.{0,2}.*, base.critic_full.{0,2,4}.*, dist.{contact_mean, actuator_mean, contact_logstd,
 * actuator_logstd}.{weight,bias}); inside the vector every actor's mean and log-std heads are adjacent so that
 * they form one (2*n, H) matrix.  Returns the padded total length. */
#define SG_SPLIT_SEGMENTS 22
int sg_split_layout(int obs_dim, int hidden, int num_feet, int* offsets);
/* SplitPolicy.act / get_value / evaluate_actions (A2C/model_split.py:69-95); as sg_policy_forward, except that the
 * entropy is state dependent: entropy_rows (B) receives the per-row entropies (their mean is dist_entropy). */
int sg_split_forward(const float* params, int obs_dim, int hidden, int num_feet, const float* obs, int B, const float* noise,
                     const float* actions_in, float* value, float* action, float* logp, float* entropy_rows, void* stream);
/* PPO.update (A2C/algo/ppo.py:65-157) for a SplitPolicy: arguments as sg_ppo_update (cfg->act_dim = 7*num_feet).
 * Persistent kernel only, no allreduce callback: cfg->mode 0 / 3 keep the three H x H matrices in a shared-memory copy
 * when it fits, mode 2 reads every weight through L2 (bit-identical results); cfg->dp_ctx enables the fused peer-memory
 * exchange as in sg_ppo_update (the value / action loss AND the entropy columns of `trace` are then per-rank partial sums).
 * trace columns: {value_loss, action_loss, dist_entropy (mean over the minibatch rows), grad_norm}. */
int64_t sg_split_ppo_workspace_bytes(const sg_ppo_config* cfg);
int sg_split_ppo_update(const sg_ppo_config* cfg, float* params, float* adam_m, float* adam_v, const float* obs,
                        const float* actions, const float* value_preds, const float* returns, const float* old_logp,
                        const float* adv_stats, const int32_t* perm, const float* step_size, const float* bc2_sqrt,
                        float* trace, void* workspace, void* stream);

/* ---- data-parallel exchange over NVLink peer memory (no reference counterpart) ---------- */
/* One context per optimizer per rank.  sg_dp_create allocates this rank's exchange buffer (room for a flat
 * gradient of max_floats); the 64-byte CUDA IPC handle from sg_dp_local_handle is exchanged between the ranks
 * by the caller (any transport), concatenated in rank order and handed to sg_dp_open_peers.  A context passed
 * as cfg->dp_ctx makes the persistent kernels exchange the locally reduced gradient slices with P2P stores,
 * system-scope flags and a rank-ordered sum inside their reduce phase (one kernel, no collective call).
 * In that mode the per-step loss columns of `trace` hold this rank's PARTIAL sums (add them over ranks). */
int sg_dp_create(int rank, int world, int max_floats, void** out_ctx);
int sg_dp_local_handle(void* ctx, unsigned char* out64);
int sg_dp_open_peers(void* ctx, const unsigned char* handles_world_x_64);
int sg_dp_capacity(void* ctx);
int sg_dp_destroy(void* ctx);

/* ---- GAIL discriminator -------------------------------------------------------------------- */
typedef struct sg_disc_config {
    int feat_dim, hidden;
    int batch_size;           /* rows per minibatch (gail_batch_size) */
    int n_steps;              /* zipped minibatches in this call */
    double gp_lambda;         /* 10.0 (A2C/algo/gail.py:70) */
    double beta1, beta2, adam_eps;
    int first_adam_step;
    int row_begin, row_end;   /* data-parallel shard of every minibatch */
    int mode;                 /* as sg_ppo_config.mode */
    void* dp_ctx;             /* as sg_ppo_config.dp_ctx */
} sg_disc_config;

int64_t sg_disc_workspace_bytes(const sg_disc_config* cfg);
/* Diagnostics, as sg_ppo_phase_cycles_offset: {param image, tile phase, barrier 1, reduce+Adam, barrier 2}. */
int64_t sg_disc_phase_cycles_offset(const sg_disc_config* cfg);

/* Discriminator.update_gail_dyn (A2C/algo/gail.py:154-193) for cfg->n_steps minibatches.
 *   expert (N_exp,F) expert rows; policy_feat = obs_feat[1:] viewed (S,F)
 *   expert_idx / policy_idx int32 (n_steps, batch): DataLoader / sampler index streams
 *   alpha (n_steps, batch): the torch.rand(B,1) mixup draws (gail.py:72)
 *   trace (n_steps,3) {total loss, expert loss, policy loss} */
int sg_disc_update(const sg_disc_config* cfg, float* params, float* adam_m, float* adam_v, const float* expert,
                   const float* policy_feat, const int32_t* expert_idx, const int32_t* policy_idx,
                   const float* alpha, const float* step_size, const float* bc2_sqrt, float* trace,
                   void* workspace, sg_allreduce_fn allreduce_cb, void* allreduce_user, void* stream);

/* Discriminator.predict_reward_combined (A2C/algo/gail.py:201-210) for one (N,F) block:
 * reward = log(s+1e-7)-log(1-s+1e-7)+offset; returns = has_returns ? returns*gamma*masks+reward : reward. */
int sg_disc_predict_reward(const float* params, int feat_dim, int hidden, const float* d_in, int n_rows,
                           double gamma, const float* masks, double offset, int has_returns, float* reward,
                           float* returns, void* stream);

/* Whole-rollout reward relabel = the T-step loop of A2C/main_gail_dyn_ppo.py:275-297 with the
 * RunningMeanStd update (A2C/baselines/common/running_mean_std.py:33-56) kept on the device.
 *   obs_feat (T+1,N,F), masks (T+1,N,1), rewards (T,N,1) out
 *   disc_returns (N): the discriminator's persistent running return (in/out); has_returns as above
 *   rms_state: 3 doubles {mean, var, count} (in/out)
 *   mean_returns (T): per-step torch.mean(returns) (the caller's gail_rewards deque)
 *   workspace: sg_relabel_workspace_bytes(T,N) */
int64_t sg_relabel_workspace_bytes(int T, int N);
int sg_disc_relabel(const float* params, int feat_dim, int hidden, const float* obs_feat, const float* masks,
                    float* rewards, int T, int N, double gamma, double offset, float* disc_returns, int has_returns,
                    double* rms_state, float* mean_returns, void* workspace, void* stream);

/* The device-resident tail of the relabel alone (steps after the discriminator forward): running
 * return scan, numpy-order float32 batch moments, float64 RunningMeanStd merge, normalise + clip.
 * raw_reward (T,N) is what predict_reward_combined returns as `reward`.  Bit-exact with the reference. */
int sg_relabel_normalize(const float* raw_reward, const float* masks, float* rewards, int T, int N, double gamma,
                         float* disc_returns, int has_returns, double* rms_state, float* mean_returns, void* workspace,
                         void* stream);

/* Diagnostics (no reference counterpart): the relabel's RunningMeanStd chain divides through a precomputed
 * correctly-rounded reciprocal plus two fused residual corrections instead of three plain IEEE divisions per step.
 * This compares that division with the plain one on blocks*256*per_thread pseudo-random operand pairs, divisors
 * log-uniform in [b_lo, b_hi]; *mismatches (DEVICE pointer, one uint64, caller-zeroed) receives the count. */
int sg_selftest_division(uint64_t seed, int blocks, int per_thread, double b_lo, double b_hi, uint64_t* mismatches,
                         void* stream);

/* Diagnostics (no reference counterpart): D (M,N) = A . B^T formed by ONE CTA on the 5th-generation tensor cores
 * (tcgen05.mma kind::tf32, accumulator in tensor memory) exactly the way the large-minibatch tiles issue their
 * contractions.  A is (M,K) row-major, or (K,M) row-major when a_mn_major != 0; B is (N,K), or (K,N) when b_mn_major != 0.
 * passes = 3: fp32 operands split hi + lo, hi*hi + hi*lo + lo*hi ("3xTF32"); passes = 1: plain TF32.
 * M in {64,128}, N a multiple of 16 in [16,256], K a multiple of 8.
 * An MN-major operand needs an MN extent that is a multiple of 32 (SWIZZLE_128B_BASE32B atoms).
 * a_mn_major == 2: A is (M,K) row-major and is kept UNSPLIT in shared memory with a padded row pitch (the layout of the
 * tiles' activation masters); the hi pass reads it in place, only the lo image is built (passes must be 3).
 * h_raw_strides (HOST, 8 ints {a_lbo, a_sbo, a_k_step, b_lbo, b_sbo, b_k_step} in bytes + {a_layout_type,
 * b_layout_type} descriptor bits [61,64), or NULL): descriptor probe -- A and B are then copied verbatim into shared
 * memory (M*K and N*K floats) and read through descriptors with these strides (passes must be 1). */
int sg_selftest_mma(int M, int N, int K, int a_mn_major, int b_mn_major, int passes, const float* A, const float* B,
                    float* D, const int* h_raw_strides, void* stream);

/* ---- host side of the minibatch sampler ---------------------------------------------------------------------------------------
 * Replaces: the `torch.randperm(batch_size)` draws BatchSampler(SubsetRandomSampler(range(batch_size)), ...) makes once per PPO
 * epoch (third_party/a2c_ppo_acktr/storage.py:158-162), when T*N is in the millions and the draw is longer than the epoch's
 * kernel.  Same stream, bit for bit: ATen's randperm_cpu (n < 2^32/20) is a Fisher-Yates walk over one 32-bit mt19937 draw per
 * element.  HOST pointers only; no CUDA call is made.
 *
 * sg_host_randperm_begin: start producing n_perms consecutive permutations of n elements into out (HOST, n_perms*n int32,
 * e.g. pinned staging) from the engine state (mt_key: the 624 state words, mt_pos: index of the next word, 624 = "regenerate
 * first" -- i.e. torch.get_rng_state()'s `state` and `next` fields).  One thread runs the engine, n_threads workers run the
 * walks of different permutations concurrently.  owned_mask: bit e set = permutation e is built here; for a clear bit the
 * engine only advances past its draws and out[e] is left untouched (the ranks of one node split the walks of an update's
 * permutations and exchange the results; every rank still ends with the same engine state).  n_perms <= 64.
 * Returns a handle, or NULL (sg_last_error).
 * sg_host_randperm_wait: block until permutation e is complete in out[e*n .. (e+1)*n).
 * sg_host_randperm_end: join, write the engine state after the last draw (what torch.set_rng_state must receive so that the
 * generator is where n_perms torch.randperm(n) calls would have left it), free the handle.
 * sg_host_randperm_prefix: synchronous; the first m elements of ONE torch.randperm(n) (what a zip() with a shorter loader
 * consumes of it, third_party/a2c_ppo_acktr/algo/gail.py:159-166) while the engine still advances by all n-1 draws. */
int sg_host_randperm_prefix(const uint32_t* mt_key, int mt_pos, int64_t n, int64_t m, int32_t* out, uint32_t* mt_key_out,
                            int* mt_pos_out);
void* sg_host_randperm_begin(const uint32_t* mt_key, int mt_pos, int64_t n, int n_perms, int32_t* out, int n_threads,
                             uint64_t owned_mask);
int sg_host_randperm_wait(void* handle, int e);
int sg_host_randperm_end(void* handle, uint32_t* mt_key_out, int* mt_pos_out);

#ifdef __cplusplus
}
#endif
#endif /* SIMGAN_B200_H */
