"""PyBullet-free twins of the reference's two training drivers.

TEST INFRASTRUCTURE.  ``gail_dyn_ppo`` restates the body of ``main()`` in third_party/a2c_ppo_acktr/main_gail_dyn_ppo.py
(lines cited inline) and ``policy_refinement`` the one of third_party/a2c_ppo_acktr/main.py, statement for statement,
with three things taken out: argument parsing (``args`` comes in as a namespace with the names of arguments.py),
``make_vec_envs`` (the vec-env comes in: tests/fake_env.py) and the ``gym.make(...).getSourceCode()`` dump / log-file
handlers.  Everything on the hot path is reached exactly as the caller reaches it: through the module namespace ``M``
(``M.Policy``, ``M.algo.PPO``, ``M.gail.Discriminator``, ``M.RolloutStorage``, ``M.utils``, ``M.RunningMeanStd``,
``M.gan_utils``), which the tests bind either to ``third_party.a2c_ppo_acktr.*`` after ``simgan_b200.compat.install()``
(this package, CUDA) or to the unmodified reference modules (CPU, dev container only).

tests/test_twin_vs_reference.py runs the REAL ``main()`` of the reference next to ``gail_dyn_ppo`` on the same fake
vec-env and requires identical logs and checkpoints, so the restatement cannot drift from the caller it stands for.
"""
import os
import types
from collections import deque

import numpy as np
import torch
from torch.utils.data import DataLoader, TensorDataset


def default_args(**over):
    """Defaults of third_party/a2c_ppo_acktr/arguments.py:28-262 (only what the two drivers read)."""
    a = dict(algo="ppo", lr=3e-4, eps=1e-5, alpha=0.99, gamma=0.99, use_gae=True, gae_lambda=0.95, entropy_coef=0.01,
             value_loss_coef=0.5, max_grad_norm=0.5, seed=1, num_processes=16, num_steps=5, ppo_epoch=10, num_mini_batch=32,
             clip_param=0.2, hidden_size=64, log_interval=10, save_interval=10, num_env_steps=10e6, num_episodes=None,
             env_name="FakeCombinedEnv-v1", save_dir="./trained_models_0/", cuda=False, no_proper_time_limits=False,
             recurrent_policy=False, use_linear_lr_decay=False, warm_start="", warm_start_logstd=None, gail=False,
             gail_dyn=False, gail_traj_path="", gail_batch_size=128, gail_epoch=5, gail_traj_num=20,
             gail_downsample_frequency=20, gail_dis_hdim=100, no_alive_bonus=False, use_split_pi=False, num_feet=1,
             dup_sym=False, loss_sym=0.0)
    a.update(over)
    return types.SimpleNamespace(**a)


def _tensor_ctor(cuda):
    """``Tensor = torch.cuda.FloatTensor if args.cuda else torch.FloatTensor`` (main_gail_dyn_ppo.py:66).  torch 2.x's legacy
    CUDA constructor no longer takes NumPy arrays (``Tensor(expert_merged_sas)``, ``Tensor(sas_feat)``, ``Tensor(rews)``), so
    on CUDA the same fp32 narrowing is done on the host and moved: same values, same device, same dtype."""
    if not cuda:
        return torch.FloatTensor
    return lambda data: torch.FloatTensor(data).cuda()


def gail_dyn_ppo(args, envs, M, log):
    """main_gail_dyn_ppo.py:50-343.  ``log(dict)`` receives what the reference formats into its log line."""
    torch.manual_seed(args.seed)                                                            # :53
    torch.cuda.manual_seed_all(args.seed)
    torch.set_num_threads(1)                                                                # :64
    device = torch.device("cuda:0" if args.cuda else "cpu")
    Tensor = _tensor_ctor(args.cuda)                                                        # :66

    if args.warm_start == '':                                                               # :71-84
        if args.use_split_pi:
            actor_critic = M.SplitPolicy(envs.observation_space.shape, envs.action_space,
                                         base_kwargs={'hidden_size': args.hidden_size, 'num_feet': args.num_feet})
        else:
            actor_critic = M.Policy(envs.observation_space.shape, envs.action_space,
                                    base_kwargs={'recurrent': args.recurrent_policy, 'hidden_size': args.hidden_size})
        actor_critic.to(device)
    else:                                                                                   # :85-94
        if args.cuda:
            actor_critic, _ = torch.load(args.warm_start, weights_only=False)
        else:
            actor_critic, _ = torch.load(args.warm_start, map_location='cpu', weights_only=False)
        if args.warm_start_logstd is not None:
            actor_critic.reset_variance(envs.action_space, args.warm_start_logstd)
            actor_critic.to(device)

    save_path = os.path.join(args.save_dir, args.algo)                                      # :97
    os.makedirs(save_path, exist_ok=True)

    if args.algo == 'ppo':                                                                  # :124-135
        agent = M.algo.PPO(actor_critic, args.clip_param, args.ppo_epoch, args.num_mini_batch, args.value_loss_coef,
                           args.entropy_coef, lr=args.lr, eps=args.eps, max_grad_norm=args.max_grad_norm)
    else:
        raise ValueError("only support PPO in gail dyn")
    assert len(envs.observation_space.shape) == 1                                           # :139

    expert_sas_w_past = M.gan_utils.load_sas_wpast_from_pickle(                             # :141-145
        args.gail_traj_path, downsample_freq=int(args.gail_downsample_frequency), load_num_trajs=args.gail_traj_num)
    s_dim = expert_sas_w_past[-1].shape[1]
    a_dim = expert_sas_w_past[-2].shape[1]
    s_idx = np.array([0])                                                                   # :152-153
    a_idx = np.array([0])
    info_length = len(s_idx) * s_dim + len(a_idx) * a_dim + s_dim                           # :159
    discr = M.gail.Discriminator(info_length, args.gail_dis_hdim, device)                   # :160-162
    expert_merged_sas = M.gan_utils.select_and_merge_sas(expert_sas_w_past, a_idx=a_idx, s_idx=s_idx)
    assert expert_merged_sas.shape[1] == info_length
    expert_dataset = TensorDataset(Tensor(expert_merged_sas))                               # :165
    gail_tar_length = expert_merged_sas.shape[0] * 1.0 / args.gail_traj_num * args.gail_downsample_frequency
    drop_last = len(expert_dataset) > args.gail_batch_size                                  # :170-175
    gail_train_loader = DataLoader(expert_dataset, batch_size=args.gail_batch_size, shuffle=True, drop_last=drop_last)

    obs = envs.reset()                                                                      # :177
    rollouts = M.RolloutStorage(args.num_steps, args.num_processes, envs.observation_space.shape, envs.action_space,
                                actor_critic.recurrent_hidden_state_size, info_length)     # :179-182
    rollouts.obs[0].copy_(obs)                                                              # :185-186
    rollouts.to(device)

    episode_rewards = deque(maxlen=10000)                                                   # :188-192
    gail_rewards = deque(maxlen=10)
    total_num_episodes = 0
    j = 0
    max_num_episodes = args.num_episodes if args.num_episodes else np.inf
    num_updates = int(args.num_env_steps) // args.num_steps // args.num_processes           # :195-196
    ret_rms = M.RunningMeanStd(shape=())                                                    # :198-199

    while j < num_updates and total_num_episodes < max_num_episodes:                        # :201
        if args.use_linear_lr_decay:                                                        # :203-207
            M.utils.update_linear_schedule(agent.optimizer, j, num_updates,
                                           agent.optimizer.lr if args.algo == "acktr" else args.lr)
        for step in range(args.num_steps):                                                  # :209-236
            with torch.no_grad():
                value, action, action_log_prob, recurrent_hidden_states = actor_critic.act(
                    rollouts.obs[step], rollouts.recurrent_hidden_states[step], rollouts.masks[step])
            obs, reward, done, infos = envs.step(action)
            sas_feat = np.zeros((args.num_processes, info_length))
            for core_idx, info in enumerate(infos):
                if 'episode' in info.keys():
                    episode_rewards.append(info['episode']['r'])
                sas_info = info["sas_window"]
                sas_feat[core_idx, :] = M.gan_utils.select_and_merge_sas(sas_info, s_idx=s_idx, a_idx=a_idx)
            masks = Tensor([[0.0] if done_ else [1.0] for done_ in done])
            bad_masks = Tensor([[0.0] if 'bad_transition' in info.keys() else [1.0] for info in infos])
            rollouts.insert(obs, recurrent_hidden_states, action, action_log_prob, value, reward, masks, bad_masks,
                            Tensor(sas_feat))

        with torch.no_grad():                                                               # :238-241
            next_value = actor_critic.get_value(rollouts.obs[-1], rollouts.recurrent_hidden_states[-1],
                                                rollouts.masks[-1]).detach()

        gail_loss, gail_loss_e, gail_loss_p = None, None, None                              # :243-256
        gail_epoch = args.gail_epoch
        for _ in range(gail_epoch):
            gail_loss, gail_loss_e, gail_loss_p = discr.update_gail_dyn(gail_train_loader, rollouts)

        num_of_dones = (1.0 - rollouts.masks).sum().cpu().numpy() + args.num_processes / 2  # :258-272
        num_of_expert_dones = (args.num_steps * args.num_processes) / gail_tar_length
        d_sa = 1 - num_of_dones / (num_of_dones + num_of_expert_dones)
        if args.no_alive_bonus:
            r_sa = 0
        else:
            r_sa = np.log(d_sa) - np.log(1 - d_sa)

        for step in range(args.num_steps):                                                  # :275-297
            rollouts.rewards[step], returns = discr.predict_reward_combined(
                rollouts.obs_feat[step + 1], args.gamma, rollouts.masks[step], offset=-r_sa)
            ret_rms.update(returns.view(-1).cpu().numpy())
            rews = rollouts.rewards[step].view(-1).cpu().numpy()
            rews = np.clip(rews / np.sqrt(ret_rms.var + 1e-7), -10.0, 10.0)
            rollouts.rewards[step] = Tensor(rews).view(-1, 1)
            gail_rewards.append(torch.mean(returns).cpu().data)

        rollouts.compute_returns(next_value, args.use_gae, args.gamma, args.gae_lambda,    # :299-300
                                 not args.no_proper_time_limits)
        value_loss, action_loss, dist_entropy = agent.update(rollouts)                      # :302
        rollouts.after_update()                                                             # :304

        if (j % args.save_interval == 0 or j == num_updates - 1) and args.save_dir != "":  # :307-320
            torch.save([actor_critic, getattr(M.utils.get_vec_normalize(envs), 'ob_rms', None)],
                       os.path.join(save_path, args.env_name + ".pt"))
            torch.save([actor_critic, getattr(M.utils.get_vec_normalize(envs), 'ob_rms', None)],
                       os.path.join(save_path, args.env_name + "_" + str(j) + ".pt"))
            if args.gail:
                torch.save(discr, os.path.join(save_path, args.env_name + "_D.pt"))
                torch.save(discr, os.path.join(save_path, args.env_name + "_" + str(j) + "_D.pt"))

        if j % args.log_interval == 0 and len(episode_rewards) > 1:                         # :322-338
            total_num_steps = (j + 1) * args.num_processes * args.num_steps
            log(dict(j=j, total_num_steps=total_num_steps, n_episodes=len(episode_rewards),
                     mean_reward=float(np.mean(episode_rewards)), median_reward=float(np.median(episode_rewards)),
                     min_reward=float(np.min(episode_rewards)), max_reward=float(np.max(episode_rewards)),
                     dist_entropy=float(dist_entropy), value_loss=float(value_loss), action_loss=float(action_loss),
                     recent_gail_r=float(np.mean(gail_rewards)), gail_loss=float(gail_loss),
                     gail_loss_e=float(gail_loss_e), gail_loss_p=float(gail_loss_p)))
        total_num_episodes += len(episode_rewards)                                          # :341-343
        episode_rewards.clear()
        j += 1
    return actor_critic, discr, agent, rollouts


def policy_refinement(args, envs, M, log):
    """The PPO-only driver of the second shipped command (third_party/a2c_ppo_acktr/main.py:69-88, 136-257 with
    --warm-start ... --warm-start-logstd ... --use-linear-lr-decay): a warm-started policy gets a fresh critic
    (``reset_critic``) and a fresh log-std (``reset_variance``), then trains on the env's own reward.  The
    obs_feat column carries the observation itself (``replace_obs_with_feat`` with the identity selection)."""
    torch.manual_seed(args.seed)
    torch.cuda.manual_seed_all(args.seed)
    torch.set_num_threads(1)
    device = torch.device("cuda:0" if args.cuda else "cpu")
    Tensor = _tensor_ctor(args.cuda)

    if args.warm_start == '':                                                               # main.py:72-77
        actor_critic = M.Policy(envs.observation_space.shape, envs.action_space,
                                base_kwargs={'recurrent': args.recurrent_policy, 'hidden_size': args.hidden_size})
        actor_critic.to(device)
    else:                                                                                   # main.py:78-88
        if args.cuda:
            actor_critic, _ = torch.load(args.warm_start, weights_only=False)
        else:
            actor_critic, _ = torch.load(args.warm_start, map_location='cpu', weights_only=False)
        actor_critic.reset_critic(envs.observation_space.shape)
        if args.warm_start_logstd is not None:
            actor_critic.reset_variance(envs.action_space, args.warm_start_logstd)
        actor_critic.to(device)

    save_path = os.path.join(args.save_dir, args.algo)
    os.makedirs(save_path, exist_ok=True)
    assert args.algo == 'ppo' and not args.loss_sym > 0.0 and not args.dup_sym              # the branches the shipped command takes
    agent = M.algo.PPO(actor_critic, args.clip_param, args.ppo_epoch, args.num_mini_batch, args.value_loss_coef,
                       args.entropy_coef, lr=args.lr, eps=args.eps, max_grad_norm=args.max_grad_norm)   # main.py:149-158

    def replace_obs_with_feat(obs):                 # my_pybullet_envs/utils.py:310-331 with feat_select_func=None
        feat_chunk = torch.Tensor(np.array([o for o in (obs.cpu() if args.cuda else obs).detach().numpy()]))
        return feat_chunk.cuda() if args.cuda else feat_chunk

    obs = envs.reset()                                                                      # main.py:166-168
    obs_feat = replace_obs_with_feat(obs)
    feat_len = obs_feat.size(1)
    rollouts = M.RolloutStorage(args.num_steps, args.num_processes, envs.observation_space.shape, envs.action_space,
                                actor_critic.recurrent_hidden_state_size, feat_len)        # main.py:174-178
    rollouts.to(device)
    rollouts.obs[0].copy_(obs)                                                              # main.py:186-187
    rollouts.obs_feat[0].copy_(obs_feat)

    episode_rewards = deque(maxlen=10000)
    total_num_episodes = 0
    j = 0
    max_num_episodes = args.num_episodes if args.num_episodes else np.inf
    num_updates = int(args.num_env_steps) // args.num_steps // args.num_processes
    while j < num_updates and total_num_episodes < max_num_episodes:                        # main.py:199
        if args.use_linear_lr_decay:                                                        # main.py:201-205
            M.utils.update_linear_schedule(agent.optimizer, j, num_updates,
                                           agent.optimizer.lr if args.algo == "acktr" else args.lr)
        for step in range(args.num_steps):                                                  # main.py:207-243
            with torch.no_grad():
                value, action, action_log_prob, recurrent_hidden_states = actor_critic.act(
                    rollouts.obs[step, :args.num_processes, :],
                    rollouts.recurrent_hidden_states[step, :args.num_processes, :],
                    rollouts.masks[step, :args.num_processes, :])
            obs, reward, done, infos = envs.step(action)
            obs_feat = replace_obs_with_feat(obs)
            for info in infos:
                if 'episode' in info.keys():
                    episode_rewards.append(info['episode']['r'])
            masks = Tensor([[0.0] if done_ else [1.0] for done_ in done])
            bad_masks = Tensor([[0.0] if 'bad_transition' in info.keys() else [1.0] for info in infos])
            rollouts.insert(obs, recurrent_hidden_states, action, action_log_prob, value, reward, masks, bad_masks,
                            obs_feat)
        with torch.no_grad():                                                               # main.py:245-248
            next_value = actor_critic.get_value(rollouts.obs[-1], rollouts.recurrent_hidden_states[-1],
                                                rollouts.masks[-1]).detach()
        rollouts.compute_returns(next_value, args.use_gae, args.gamma, args.gae_lambda,    # main.py:250-251
                                 not args.no_proper_time_limits)
        value_loss, action_loss, dist_entropy = agent.update(rollouts)                      # main.py:253
        rollouts.after_update()
        if (j % args.save_interval == 0 or j == num_updates - 1) and args.save_dir != "":  # main.py:258-268
            torch.save([actor_critic, getattr(M.utils.get_vec_normalize(envs), 'ob_rms', None)],
                       os.path.join(save_path, args.env_name + ".pt"))
            torch.save([actor_critic, getattr(M.utils.get_vec_normalize(envs), 'ob_rms', None)],
                       os.path.join(save_path, args.env_name + "_" + str(j) + ".pt"))
        if j % args.log_interval == 0 and len(episode_rewards) > 1:
            log(dict(j=j, n_episodes=len(episode_rewards), mean_reward=float(np.mean(episode_rewards)),
                     dist_entropy=float(dist_entropy), value_loss=float(value_loss), action_loss=float(action_loss),
                     lr=float(agent.optimizer.param_groups[0]["lr"])))
        total_num_episodes += len(episode_rewards)
        episode_rewards.clear()
        j += 1
    return actor_critic, agent, rollouts
