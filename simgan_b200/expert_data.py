"""Expert-trajectory input format of the discriminator (my_pybullet_envs/utils.py:170-199, 233-263).

The reference stores expert data as a pickle ``dict[int -> list[row]]`` where a row is ``2W+1`` lists
``[s_t .. s_{t-W+1}, a_t .. a_{t-W+1}, s_{t+1}]`` (written by third_party/a2c_ppo_acktr/collect_tarsim_traj.py:261-265).
``load_sas_wpast_from_pickle`` turns it into ``2W+1`` ``(N_exp, dim)`` arrays and ``select_and_merge_sas`` into the
``(N_exp, F)`` discriminator input ``[s_t | a_t | s_{t+1}]`` (main_gail_dyn_ppo.py:141-159).

Same names, arguments and generator consumption as the reference (one ``torch.randint(0, freq, (n_trajs,))`` on the
CPU default generator).  One deliberate difference: the reference builds the columns through a ragged
``np.array(sas)`` (utils.py:193), which raises on numpy >= 1.24; columns are stacked one by one instead, giving the
arrays numpy < 1.24 produced.  ``expert_tensor`` is the device-resident form the kernels read in place.
"""
import pickle

import numpy as np
import torch


def load_sas_wpast_from_pickle(pathname, downsample_freq=1, load_num_trajs=None):
    """List of 2W+1 arrays (N_exp, dim).  If ``load_num_trajs`` is None all trajectories are loaded."""
    with open(pathname, "rb") as handle:
        saved_file = pickle.load(handle)
    n_trajs = len(saved_file)
    start_idx = torch.randint(0, downsample_freq, size=(n_trajs,)).long()
    sas = []
    for traj_idx, traj_tuples in saved_file.items():
        sas.extend(traj_tuples[int(start_idx[traj_idx])::downsample_freq])
        if load_num_trajs and traj_idx >= load_num_trajs - 1:
            break
    if not sas:
        raise ValueError("no expert rows in %s" % pathname)
    width = len(sas[0])
    return [np.array([row[item] for row in sas]) for item in range(width)]


def select_and_merge_sas(sas, s_idx=np.array([0, ]), a_idx=np.array([0, ])):
    """``sas``: list of 2W+1 elements, each a length-l list (one transition) or an (N, l) array.
    Returns the (N, F) matrix -- or the length-F vector for a single transition -- of
    [s_{t-i} for i in s_idx | a_{t-j} for j in a_idx | s_{t+1}] in float64."""
    is_one_dim = np.array(sas[0]).ndim == 1
    cols = [np.array([c]) if is_one_dim else np.asarray(c) for c in sas]
    assert cols[-1].ndim == 2
    len_time_win = (len(sas) - 1) // 2
    merged = np.array([]).reshape((cols[0].shape[0], 0))
    for i in s_idx:
        merged = np.concatenate((merged, cols[i]), axis=1)
    for j in a_idx:
        merged = np.concatenate((merged, cols[len_time_win + j]), axis=1)
    merged = np.concatenate((merged, cols[-1]), axis=1)
    return merged[0, :] if is_one_dim else merged


def expert_tensor(pathname, device, downsample_freq=1, load_num_trajs=None, s_idx=np.array([0, ]), a_idx=np.array([0, ])):
    """(N_exp, F) fp32 tensor on ``device`` -- the matrix ``TensorDataset(Tensor(expert_merged))`` wraps in the
    caller (main_gail_dyn_ppo.py:165-175); the update kernels gather expert rows from it in place."""
    merged = select_and_merge_sas(load_sas_wpast_from_pickle(pathname, downsample_freq, load_num_trajs), s_idx, a_idx)
    return torch.as_tensor(merged, dtype=torch.float32).to(device)
