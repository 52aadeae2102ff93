"""oracle/search_ref.py -- restatement of the reference's per-simulation host loop
``MuZeroMCTSCtree.search`` (lzero/mcts/tree_search/mcts_ctree.py:267-368) and of the part of
``MuZeroPolicy._forward_collect`` that feeds it (lzero/policy/muzero.py:749-779).

TEST INFRASTRUCTURE ONLY.  `lzero.mcts.tree_search.mcts_ctree` cannot be imported here (it pulls
lzero.policy -> ding/easydict, absent), so the ~60-line loop is restated; the tree module it drives
is either the compiled UNMODIFIED reference (oracle/_ref/mz_tree, preferred) or oracle/ctree_port.

Differences from the reference loop, all deliberate and stated in DESIGN.md:
  * ``deterministic=True`` is passed to batch_traverse (the parity contract);
  * ``duplicate_inference`` reproduces the reference's discarded first recurrent_inference call
    (mcts_ctree.py:338 then :345) when timing the reference arm; parity runs set it False;
  * an optional ``recorder`` captures per-simulation (ix, iy, last_action, reward, value, logits)
    for replay-mode tree parity tests.
"""
import copy
import os
import sys

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def load_tree_module(prefer_ref: bool = True):
    """Returns (module, kind) with kind in {'reference', 'port'}."""
    if prefer_ref:
        ref_dir = os.path.join(_HERE, "_ref")
        if os.path.isdir(ref_dir) and any(f.startswith("mz_tree") for f in os.listdir(ref_dir)):
            if ref_dir not in sys.path:
                sys.path.insert(0, ref_dir)
            try:
                import mz_tree  # noqa
                return mz_tree, "reference"
            except ImportError:
                pass
    from oracle import ctree_port
    return ctree_port, "port"


class SearchRef:
    def __init__(self, tree_module, num_simulations=50, pb_c_base=19652, pb_c_init=1.25,
                 discount_factor=0.997, value_delta_max=0.01, env_type="not_board_games",
                 value_support_range=(-300., 301., 1.), reward_support_range=(-300., 301., 1.),
                 deterministic=True, duplicate_inference=False):
        from oracle.model_ref import DiscreteSupport, InverseScalarTransform
        self.tree = tree_module
        self.num_simulations = num_simulations
        self.pb_c_base, self.pb_c_init = pb_c_base, pb_c_init
        self.discount_factor, self.value_delta_max = discount_factor, value_delta_max
        self.env_type = env_type
        self.deterministic = deterministic
        self.duplicate_inference = duplicate_inference
        self.value_inv = InverseScalarTransform(DiscreteSupport(*value_support_range))
        self.reward_inv = InverseScalarTransform(DiscreteSupport(*reward_support_range))

    def roots(self, n, legal_actions, action_space_size=None):
        if self.tree.__name__.endswith("ctree_port"):
            return self.tree.Roots(n, legal_actions, action_space_size=action_space_size,
                                   max_sims=self.num_simulations)
        return self.tree.Roots(n, legal_actions)

    def search(self, roots, model, latent_state_roots, to_play_batch, recorder=None, replay=None):
        """mcts_ctree.py:281-368.  latent_state_roots: np.ndarray [B,C,H,W]."""
        tree = self.tree
        with torch.no_grad():
            if model is not None:
                model.eval()
            batch_size = roots.num
            latent_pool = [latent_state_roots]
            mm = tree.MinMaxStatsList(batch_size)
            mm.set_delta(self.value_delta_max)
            for simulation_index in range(self.num_simulations):
                results = tree.ResultsWrapper(batch_size)
                tp_arg = to_play_batch if self.env_type == "not_board_games" else copy.deepcopy(to_play_batch)
                ix_l, iy_l, last_actions, virtual_to_play = tree.batch_traverse(
                    roots, self.pb_c_base, self.pb_c_init, self.discount_factor, mm, results,
                    list(tp_arg), self.deterministic)
                if replay is not None:
                    rec = replay[simulation_index]
                    reward_batch, value_batch, policy_logits_batch = rec["reward"], rec["value"], rec["policy"]
                else:
                    latent_states = [latent_pool[ix][iy] for ix, iy in zip(ix_l, iy_l)]
                    latent_states = torch.from_numpy(np.asarray(latent_states))
                    actions = torch.from_numpy(np.asarray(last_actions)).long()
                    if self.duplicate_inference:
                        model.recurrent_inference(latent_states, actions)   # mcts_ctree.py:338 (discarded)
                    out = model.recurrent_inference(latent_states, actions)
                    latent_pool.append(out.latent_state.detach().cpu().numpy())
                    value = self.value_inv(out.value).detach().cpu().numpy()
                    reward = self.reward_inv(out.reward).detach().cpu().numpy()
                    reward_batch = reward.reshape(-1).tolist()
                    value_batch = value.reshape(-1).tolist()
                    policy_logits_batch = out.policy_logits.detach().cpu().numpy().tolist()
                if recorder is not None:
                    recorder.append(dict(ix=list(ix_l), iy=list(iy_l), last_action=list(last_actions),
                                         search_len=list(results.get_search_len()),
                                         reward=list(reward_batch), value=list(value_batch),
                                         policy=[list(p) for p in policy_logits_batch]))
                tree.batch_backpropagate(simulation_index + 1, self.discount_factor, reward_batch,
                                         value_batch, policy_logits_batch, mm, results, virtual_to_play)
        return latent_pool


def collect_step_ref(search: SearchRef, model, obs, action_mask, to_play, noise_weight=0.25,
                     noises=None, recorder=None):
    """The search-feeding part of MuZeroPolicy._forward_collect (policy/muzero.py:749-779)."""
    with torch.no_grad():
        model.eval()
        out = model.initial_inference(obs)
        pred_values = search.value_inv(out.value).detach().cpu().numpy()
        latent_roots = out.latent_state.detach().cpu().numpy()
        policy_logits = out.policy_logits.detach().cpu().numpy()
    B = obs.shape[0]
    legal_actions = [np.nonzero(action_mask[j])[0].tolist() for j in range(B)]
    roots = search.roots(B, legal_actions, action_space_size=policy_logits.shape[1])
    if noises is not None:
        roots.prepare(noise_weight, noises, [0.] * B, policy_logits.tolist(), list(to_play))
    else:
        roots.prepare_no_noise([0.] * B, policy_logits.tolist(), list(to_play))
    search.search(roots, model, latent_roots, list(to_play), recorder=recorder)
    return dict(distributions=roots.get_distributions(), values=roots.get_values(),
                pred_values=pred_values, policy_logits=policy_logits, latent_roots=latent_roots,
                roots=roots)
