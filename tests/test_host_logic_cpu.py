"""CPU tests of the host-side mirrors that need no device: select_action (lzero/policy/utils.py:637-661), the config object of
the MCTS mirrors (mcts_ctree.py:220-253); and that the scalar transform refuses to run without a device."""
import numpy as np
import torch
from scipy.stats import entropy


def test_select_action_follows_the_reference_formula():
    from lightzero_b200.collect import select_action
    rng = np.random.default_rng(0)
    for _ in range(50):
        v = rng.integers(0, 30, size=rng.integers(2, 19)).astype(np.int64)
        v[rng.integers(len(v))] += 1
        for temp in (1.0, 0.5, 0.25):
            pos, ent = select_action(v, temperature=temp, deterministic=True)
            probs = v.astype(np.float64) ** (1 / temp)
            probs /= probs.sum()
            assert pos == int(np.argmax(v))
            assert abs(ent - entropy(probs, base=2)) < 1e-12       # lzero/policy/utils.py:660
    np.random.seed(0)
    draws = [select_action(np.array([1, 0, 3]), temperature=1.0, deterministic=False)[0] for _ in range(400)]
    assert 1 not in draws and 0.15 < draws.count(0) / 400 < 0.35


def test_mcts_config_defaults_and_overrides():
    from lightzero_b200.mcts_ctree import EfficientZeroMCTSCtree, MuZeroMCTSCtree
    cfg = MuZeroMCTSCtree.default_config()
    assert cfg.pb_c_base == 19652 and cfg.pb_c_init == 1.25 and cfg.root_dirichlet_alpha == 0.3 and cfg.root_noise_weight == 0.25
    assert cfg.value_delta_max == 0.01 and cfg.env_type == "not_board_games" and cfg.cfg_type == "MuZeroMCTSCtreeDict"
    m = MuZeroMCTSCtree(dict(num_simulations=7, pb_c_init=2.0, model=dict(value_support_range=(-10., 11., 1.))))
    assert m._cfg.num_simulations == 7 and m._cfg.pb_c_init == 2.0 and m._cfg.model.value_support_range == (-10., 11., 1.)
    assert m._params() == (19652, 2.0, 0.997, 0.01) and m.deterministic is False
    e = EfficientZeroMCTSCtree(dict(lstm_horizon_len=3))
    assert e._cfg.lstm_horizon_len == 3 and EfficientZeroMCTSCtree({})._cfg.lstm_horizon_len == 5


def test_inverse_scalar_transform_has_no_cpu_path():
    """The product package has no CPU fallback: without a device the transform raises instead of computing on the host."""
    import pytest
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    from lightzero_b200.cabi import LzError
    from lightzero_b200.scaling_transform import DiscreteSupport, InverseScalarTransform
    with pytest.raises((RuntimeError, LzError, AssertionError)):
        InverseScalarTransform(DiscreteSupport(-300., 301., 1., device="cpu"))(torch.zeros(2, 601))


def test_unizero_and_efficientzero_tiebreak_defaults():
    """UniZeroMCTSCtree: deterministic=False by default (mcts_ctree.py:41-42); EfficientZeroMCTSCtree: first maximum unless the config says
    otherwise (the reference tree has no switch; lz_tree_set_tiebreak)."""
    from lightzero_b200.mcts_ctree import EfficientZeroMCTSCtree, UniZeroMCTSCtree
    assert UniZeroMCTSCtree.default_config().deterministic is False and UniZeroMCTSCtree.default_config().cfg_type == "UniZeroMCTSCtreeDict"
    assert UniZeroMCTSCtree({}).deterministic is False and UniZeroMCTSCtree(dict(deterministic=True)).deterministic is True
    assert EfficientZeroMCTSCtree({}).deterministic is True and EfficientZeroMCTSCtree(dict(deterministic=False)).deterministic is False


def test_segment_pack_unpack_round_trip():
    """collector.pack_segments / unpack_segments: the packed buffer of the finished-segment all-gather is lossless."""
    from lightzero_b200.collector import pack_segments, unpack_segments
    g = torch.Generator().manual_seed(0)
    B, T, A = 5, 7, 18
    cv, rv = torch.rand(B, T, A, generator=g), torch.rand(B, T, generator=g)
    ln = torch.randint(0, T + 1, (B,), generator=g, dtype=torch.int32)
    packed = pack_segments(cv, rv, ln)
    assert packed.shape == (B, T * A + T + 1) and packed.is_contiguous()
    cv2, rv2, ln2 = unpack_segments(packed, T, A)
    assert torch.equal(cv2, cv) and torch.equal(rv2, rv) and torch.equal(ln2, ln)


def test_collector_state_has_no_cpu_path():
    import pytest
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    from lightzero_b200.collector import FrameStack, SegmentStats
    with pytest.raises(Exception):
        FrameStack(2, 4, 84, 84)
    with pytest.raises(Exception):
        SegmentStats(2, 8, 6)
