"""Mirror of the reanalyze call site of the search, ``MuZeroGameBuffer._compute_target_policy_reanalyzed``
(lzero/mcts/buffer/game_buffer_muzero.py:578-730): the replay buffer re-searches
``batch_size * (num_unroll_steps + 1)`` stored observations with the latest model and turns the root
visit counts into policy targets.  This is the second caller of ``MuZeroMCTSCtree.search`` in the
reference and its largest natural batch (1536 roots by default).  With an ``EfficientZeroMCTSCtree`` it follows
``EfficientZeroGameBuffer._compute_target_policy_reanalyzed`` (game_buffer_efficientzero.py:325-440): the same flow with
the (zero) reward hidden state of ``initial_inference`` handed to the value-prefix search.

Everything between the observation batch and the visit counts stays on the GPU: ``initial_inference``
in ``mini_infer_size`` slices (:612-633), root preparation with/without Dirichlet noise (:644-652), one
CUDA-graph search (:654-658); the per-segment target assembly (:676-730) is host-side bookkeeping on
the small [roots, A] result and follows the reference line by line.
"""
from typing import List, Optional, Sequence

import numpy as np
import torch

from .mcts_ctree import EfficientZeroMCTSCtree, MuZeroMCTSCtree


def compute_target_policy_reanalyzed(
        model, mcts: MuZeroMCTSCtree, policy_obs: torch.Tensor, action_mask: np.ndarray, to_play: Sequence[int],
        policy_mask: Sequence[int], pos_in_game_segment_list: Sequence[int], child_visits: Optional[List[list]] = None,
        num_unroll_steps: int = 5, action_space_size: Optional[int] = None, action_type: str = "fixed_action_space",
        reanalyze_noise: bool = True, mini_infer_size: int = 10240, noises: Optional[np.ndarray] = None,
        return_roots: bool = False):
    """policy_obs: [transition_batch_size, ...] observations (host or device); action_mask [T, A] 0/1;
    to_play [T]; policy_mask [T] (0 = padding beyond the episode end); pos_in_game_segment_list [segments];
    child_visits: per segment, the list the reference updates in place with the fresh distributions.
    Returns np.ndarray [segments, num_unroll_steps + 1, A] (game_buffer_muzero.py:728)."""
    cfg = mcts._cfg
    T = policy_obs.shape[0]
    action_mask = np.asarray(action_mask)
    A = action_space_size or action_mask.shape[1]
    legal_actions = [np.nonzero(action_mask[j])[0].tolist() for j in range(T)]          # :610
    dev = model.device
    lat, logits, hid0, hid1 = [], [], [], []
    ez = isinstance(mcts, EfficientZeroMCTSCtree)      # EfficientZeroGameBuffer._compute_target_policy_reanalyzed, game_buffer_efficientzero.py:325-440
    for beg in range(0, T, mini_infer_size):                                              # :612-633
        out = model.initial_inference(policy_obs[beg:beg + mini_infer_size].to(dev, non_blocking=True))
        lat.append(out.latent_state)
        logits.append(out.policy_logits)
        if ez:                                                                            # game_buffer_efficientzero.py:365-372
            hid0.append(out.reward_hidden_state[0])
            hid1.append(out.reward_hidden_state[1])
    latent_state_roots = torch.cat(lat) if len(lat) > 1 else lat[0]
    policy_logits_pool = torch.cat(logits) if len(logits) > 1 else logits[0]
    if noises is None:                                                                    # :638-641 (A entries per root)
        noises = np.stack([np.random.dirichlet([cfg.root_dirichlet_alpha] * A).astype(np.float32) for _ in range(T)])
    roots = mcts.roots(T, legal_actions)                                                  # :644
    reward_pool = [0.] * T                                                                # initial_inference rewards
    if reanalyze_noise:
        roots.prepare(cfg.root_noise_weight, noises, reward_pool, policy_logits_pool, list(to_play))   # :646
    else:
        roots.prepare_no_noise(reward_pool, policy_logits_pool, list(to_play))            # :648
    if ez:
        hidden_roots = (torch.cat(hid0, dim=1), torch.cat(hid1, dim=1))                   # game_buffer_efficientzero.py:375-376
        mcts.search(roots, model, latent_state_roots, hidden_roots, list(to_play))        # :398-400
    else:
        mcts.search(roots, model, latent_state_roots, list(to_play))                      # :654-658
    roots_distributions = roots.get_distributions()                                       # :674
    roots_values = roots.get_values()

    batch_target_policies_re = []
    policy_index = 0
    if child_visits is None:
        child_visits = [dict() for _ in pos_in_game_segment_list]
    for state_index, child_visit in zip(pos_in_game_segment_list, child_visits):         # :679
        target_policies = []
        for current_index in range(state_index, state_index + num_unroll_steps + 1):
            distributions = roots_distributions[policy_index]
            if policy_mask[policy_index] == 0:                                            # :686-688
                target_policies.append([0 for _ in range(A)])
            else:
                sum_visits = sum(distributions)
                child_visit[current_index] = [v / sum_visits for v in distributions]      # :691,704-705
                if action_type == "fixed_action_space":                                   # :707-711
                    target_policies.append([v / sum_visits for v in distributions])
                else:                                                                     # :712-721 (board games)
                    policy_tmp = [0 for _ in range(A)]
                    for index, legal_action in enumerate(legal_actions[policy_index]):
                        policy_tmp[legal_action] = distributions[index] / sum_visits
                    target_policies.append(policy_tmp)
            policy_index += 1
        batch_target_policies_re.append(target_policies)
    out = np.array(batch_target_policies_re)
    if return_roots:
        return out, roots_values, roots
    roots.clear()
    return out
