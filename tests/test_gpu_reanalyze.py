"""the reanalyze caller (game_buffer_muzero.py:578-730) on the CUDA engine vs the same
function restated over the oracle pipeline (PyTorch-CPU model + reference ctree)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _oracle_targets(ref, obs, mask, to_play, policy_mask, pos_list, unroll, noises, S, board):
    from oracle.search_ref import SearchRef, load_tree_module
    tree, _ = load_tree_module()
    s = SearchRef(tree, num_simulations=S)
    T, A = mask.shape
    legal = [np.nonzero(mask[j])[0].tolist() for j in range(T)]
    with torch.no_grad():
        out = ref.initial_inference(obs)
    roots = s.roots(T, legal, action_space_size=A)
    roots.prepare(0.25, [n.tolist() for n in noises], [0.] * T, out.policy_logits.numpy().tolist(), list(to_play))
    s.search(roots, ref, out.latent_state.numpy(), list(to_play))
    dists = roots.get_distributions()
    res, pi = [], 0
    for st in pos_list:
        seg = []
        for _ in range(unroll + 1):
            d = dists[pi]
            if policy_mask[pi] == 0:
                seg.append([0] * A)
            elif not board:
                seg.append([v / sum(d) for v in d])
            else:
                tmp = [0] * A
                for i, a in enumerate(legal[pi]):
                    tmp[a] = d[i] / sum(d)
                seg.append(tmp)
            pi += 1
        res.append(seg)
    return np.array(res), dists


@pytest.mark.parametrize("board", [False, True])
def test_reanalyze_targets_match_reference_pipeline(board):
    import lightzero_b200 as lzb
    from lightzero_b200.reanalyze import compute_target_policy_reanalyzed
    from oracle.model_ref import MuZeroModelRef, emulate_trained_
    torch.manual_seed(1)
    A, S, unroll, segs = 6, 25, 5, 8
    T = segs * (unroll + 1)
    ref = emulate_trained_(MuZeroModelRef((4, 84, 84), A), 1)
    cu = lzb.MuZeroModel(observation_shape=(4, 84, 84), action_space_size=A).load_state_dict(ref.state_dict())
    rng = np.random.default_rng(2)
    obs = torch.rand(T, 4, 84, 84)
    if board:
        mask = (rng.random((T, A)) < 0.7).astype(np.int8)
        mask[np.arange(T), rng.integers(0, A, T)] = 1
        to_play = rng.integers(1, 3, T).tolist()
    else:
        mask = np.ones((T, A), np.int8)
        to_play = [-1] * T
    policy_mask = (rng.random(T) < 0.9).astype(int).tolist()
    pos_list = rng.integers(0, 50, segs).tolist()
    noises = np.stack([rng.dirichlet([0.3] * A).astype(np.float32) for _ in range(T)])
    mcts = lzb.MuZeroMCTSCtree(dict(num_simulations=S, deterministic=True, discount_factor=0.997,
                                    env_type="board_games" if board else "not_board_games"))
    child_visits = [dict() for _ in range(segs)]
    got = compute_target_policy_reanalyzed(cu, mcts, obs, mask, to_play, policy_mask, pos_list, child_visits, unroll, A,
                                           "varied_action_space" if board else "fixed_action_space", True, 32, noises)
    exp, _ = _oracle_targets(ref, obs, mask, to_play, policy_mask, pos_list, unroll, noises, S, board)
    assert got.shape == exp.shape == (segs, unroll + 1, A)
    flat_g, flat_e = got.reshape(T, A), exp.reshape(T, A)
    same = sum(np.array_equal(a, b) for a, b in zip(flat_g, flat_e))
    assert same >= int(0.85 * T), same
    for j in range(T):
        if policy_mask[j] == 0:
            assert not flat_g[j].any()
        else:
            assert abs(flat_g[j].sum() - 1.0) < 1e-12 and (flat_g[j][mask[j] == 0] == 0).all()
    # the reference updates child_visit[current_index] in place for every unmasked position (:691)
    assert sum(len(c) for c in child_visits) == sum(policy_mask)


def test_reanalyze_default_batch_size_runs():
    """256 x (5 + 1) = 1536 roots, the reference's default reanalyze batch."""
    import lightzero_b200 as lzb
    from lightzero_b200.reanalyze import compute_target_policy_reanalyzed
    from oracle.model_ref import MuZeroModelRef, emulate_trained_
    A, S, unroll, segs = 6, 50, 5, 256
    T = segs * (unroll + 1)
    ref = emulate_trained_(MuZeroModelRef((4, 84, 84), A), 0)
    cu = lzb.MuZeroModel(observation_shape=(4, 84, 84), action_space_size=A).load_state_dict(ref.state_dict())
    mcts = lzb.MuZeroMCTSCtree(dict(num_simulations=S, deterministic=True, discount_factor=0.997))
    obs = torch.rand(T, 4, 84, 84)
    out = compute_target_policy_reanalyzed(cu, mcts, obs, np.ones((T, A), np.int8), [-1] * T, [1] * T, [0] * segs, None,
                                           unroll, A, mini_infer_size=1024)
    assert out.shape == (segs, unroll + 1, A) and np.allclose(out.sum(-1), 1.0)


def test_reanalyze_efficientzero_targets_match_reference_pipeline():
    """EfficientZeroGameBuffer._compute_target_policy_reanalyzed (game_buffer_efficientzero.py:325-440): the same caller with
    the value-prefix search and the zero reward hidden state of initial_inference."""
    import lightzero_b200 as lzb
    from lightzero_b200.reanalyze import compute_target_policy_reanalyzed
    from oracle.model_ref import EfficientZeroModelRef, emulate_trained_
    from oracle.search_ref import SearchRefEZ, load_tree_module
    torch.manual_seed(4)
    A, S, unroll, segs, H = 6, 20, 5, 6, 3
    T = segs * (unroll + 1)
    ref = emulate_trained_(EfficientZeroModelRef((4, 96, 96), A), 4)
    cu = lzb.EfficientZeroModel(observation_shape=(4, 96, 96), action_space_size=A).load_state_dict(ref.state_dict())
    rng = np.random.default_rng(5)
    obs = torch.rand(T, 4, 96, 96)
    mask = np.ones((T, A), np.int8)
    to_play = [-1] * T
    policy_mask = (rng.random(T) < 0.9).astype(int).tolist()
    pos_list = rng.integers(0, 50, segs).tolist()
    noises = np.stack([rng.dirichlet([0.3] * A).astype(np.float32) for _ in range(T)])
    mcts = lzb.EfficientZeroMCTSCtree(dict(num_simulations=S, discount_factor=0.997, lstm_horizon_len=H))
    got = compute_target_policy_reanalyzed(cu, mcts, obs, mask, to_play, policy_mask, pos_list, None, unroll, A,
                                           "fixed_action_space", True, 16, noises)
    tree, _ = load_tree_module(name="ez_tree")
    s = SearchRefEZ(tree, lstm_horizon_len=H, num_simulations=S)
    with torch.no_grad():
        out = ref.initial_inference(obs)
    roots = s.roots(T, [list(range(A))] * T, action_space_size=A)
    roots.prepare(0.25, [n.tolist() for n in noises], [0.] * T, out.policy_logits.numpy().tolist(), to_play)
    s.search(roots, ref, out.latent_state.numpy(), (out.reward_hidden_state[0].numpy(), out.reward_hidden_state[1].numpy()), to_play)
    exp = np.array([[v / sum(d) for v in d] if policy_mask[j] else [0] * A for j, d in enumerate(roots.get_distributions())])
    flat = got.reshape(T, A)
    same = sum(np.array_equal(a, b) for a, b in zip(flat, exp))
    assert same >= int(0.85 * T), same
    assert got.shape == (segs, unroll + 1, A)
