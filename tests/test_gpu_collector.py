"""GPU tests of the device-resident collector state (csrc/collector.cu) through the C ABI mirrors."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_frame_stack_matches_the_reference_obs_window():
    """GameSegment.get_obs() over a reset([init] * stack) + append() history (game_segment.py:140-181,
    muzero_collector.py:451-457) == FrameStack after the same pushes, including mid-run episode resets."""
    from lightzero_b200.collector import FrameStack
    B, S, H, W = 37, 4, 84, 84
    rng = np.random.default_rng(0)
    fs = FrameStack(B, S, H, W)
    first = rng.integers(0, 256, (B, H, W), dtype=np.uint8)
    fs.push(first, reset=np.ones(B, np.uint8))
    obs_segment = [[first[b]] * S for b in range(B)]           # the reference's per-env obs list after reset()
    assert np.array_equal(fs.get_obs().cpu().numpy(), np.stack([np.stack(o[-S:]) for o in obs_segment]))
    for step in range(7):
        new = rng.integers(0, 256, (B, H, W), dtype=np.uint8)
        reset = (rng.random(B) < 0.2).astype(np.uint8) if step % 2 else None
        if step == 3:                                          # device-resident input path
            fs.push(torch.from_numpy(new).cuda(), None if reset is None else torch.from_numpy(reset).cuda())
        else:
            fs.push(new, reset)
        for b in range(B):
            if reset is not None and reset[b]:
                obs_segment[b] = [new[b]] * S
            else:
                obs_segment[b].append(new[b])
        assert np.array_equal(fs.get_obs().cpu().numpy(), np.stack([np.stack(o[-S:]) for o in obs_segment])), step


def test_frame_stack_feeds_the_uint8_collect_entry_point():
    """The stacked device batch is what lz_search_collect_u8 consumes: same visit counts as uploading the full uint8 stack."""
    import lightzero_b200 as lzb
    from lightzero_b200.collect import MuZeroCollectPolicy
    from lightzero_b200.collector import FrameStack
    from lightzero_b200.synthetic_weights import synthetic_state_dict
    B, A, S = 32, 6, 10
    model = lzb.MuZeroModel(observation_shape=(4, 84, 84), action_space_size=A).load_state_dict(synthetic_state_dict((4, 84, 84), A))
    rng = np.random.default_rng(1)
    frames = rng.integers(0, 256, (5, B, 84, 84), dtype=np.uint8)
    fs = FrameStack(B, 4, 84, 84)
    fs.push(frames[0], reset=np.ones(B, np.uint8))
    for t in range(1, 5):
        fs.push(frames[t])
    stacked = fs.get_obs()
    assert np.array_equal(stacked.cpu().numpy(), np.transpose(frames[1:5], (1, 0, 2, 3)))
    pol = MuZeroCollectPolicy(model, dict(num_simulations=S, deterministic=True, discount_factor=0.997))
    mask = np.ones((B, A), np.uint8)
    noise = rng.dirichlet([0.3] * A, size=B).astype(np.float32)
    out_a = pol.search_batch(stacked.cpu(), mask, noise)                         # full uint8 stack uploaded from the host
    va = np.asarray(out_a["visits"]).copy()
    out_b = pol.search_batch(stacked, torch.from_numpy(mask).cuda(), torch.from_numpy(noise).cuda())   # the device-resident stack
    assert np.array_equal(va, np.asarray(out_b["visits"]))
    assert (va.sum(1) == S).all()


def test_store_search_stats_matches_game_segment():
    """GameSegment.store_search_stats (game_segment.py:241-263): float64 division stored as float32, 1e-6 for an all-zero row,
    appended only for the environments that stepped, full segments left alone, reset() per environment."""
    from lightzero_b200.collector import SegmentStats
    B, T, A = 50, 6, 18
    rng = np.random.default_rng(2)
    seg = SegmentStats(B, T, A)
    ref_cv = [[] for _ in range(B)]
    ref_rv = [[] for _ in range(B)]
    for step in range(9):
        nlegal = rng.integers(1, A + 1, B)
        visits = np.full((B, A), -1, np.int32)
        for b in range(B):
            visits[b, :nlegal[b]] = rng.multinomial(50, rng.dirichlet([0.3] * nlegal[b]))
        if step == 4:
            visits[3, :nlegal[3]] = 0                                            # the reference's 1e-6 branch
        values = rng.standard_normal(B).astype(np.float32)
        active = (rng.random(B) < 0.8).astype(np.uint8)
        seg.store_search_stats(torch.from_numpy(visits).cuda(), torch.from_numpy(values).cuda(), active)
        for b in range(B):
            if active[b] and len(ref_cv[b]) < T:
                vc = [int(v) for v in visits[b, :nlegal[b]]]
                s = sum(vc)
                if s == 0:
                    s = 1e-6
                ref_cv[b].append([v / s for v in vc] + [0.0] * (A - nlegal[b]))
                ref_rv[b].append(values[b])
        if step == 5:
            done = (rng.random(B) < 0.3).astype(np.uint8)
            seg.reset(done)
            for b in range(B):
                if done[b]:
                    ref_cv[b], ref_rv[b] = [], []
    cv, rv, ln = (t.cpu().numpy() for t in seg.tensors())
    for b in range(B):
        n = len(ref_cv[b])
        assert ln[b] == n
        if n:
            assert np.array_equal(cv[b, :n], np.asarray(ref_cv[b], np.float64).astype(np.float32)), b
            assert np.array_equal(rv[b, :n], np.asarray(ref_rv[b], np.float32)), b


def test_collect_step_without_read_back_returns_device_actions():
    """The collector loop of INTEGRATION.md section 5: search_batch(read_back=False, select=...) leaves visits / values / actions on the
    device (no synchronisation inside the step) and they equal the read-back variant."""
    import lightzero_b200 as lzb
    from lightzero_b200.collect import MuZeroCollectPolicy
    from lightzero_b200.collector import FrameStack, SegmentStats
    from lightzero_b200.synthetic_weights import synthetic_state_dict
    B, A, S = 24, 6, 8
    model = lzb.MuZeroModel(observation_shape=(4, 84, 84), action_space_size=A).load_state_dict(synthetic_state_dict((4, 84, 84), A))
    pol = MuZeroCollectPolicy(model, dict(num_simulations=S, deterministic=True, discount_factor=0.997))
    rng = np.random.default_rng(7)
    fs = FrameStack(B, 4, 84, 84)
    fs.push(rng.integers(0, 256, (B, 84, 84), dtype=np.uint8), reset=np.ones(B, np.uint8))
    mask = torch.ones(B, A, dtype=torch.uint8).cuda()
    noise = torch.from_numpy(rng.dirichlet([0.3] * A, size=B).astype(np.float32)).cuda()
    dev = pol.search_batch(fs.view(), mask, noise, read_back=False, select=(1.0, True, 3))
    assert dev["action"].is_cuda and dev["visits"].is_cuda
    host = pol.search_batch(fs.view(), mask, noise, read_back=True, select=(1.0, True, 3))
    assert np.array_equal(dev["visits"].cpu().numpy(), np.asarray(host["visits"]))
    assert np.array_equal(dev["action"].cpu().numpy(), np.asarray(host["action"]))
    seg = SegmentStats(B, 4, A)
    seg.store_search_stats(dev["visits"], dev["values"])
    cv, rv, ln = seg.tensors()
    assert (ln == 1).all() and torch.allclose(cv[:, 0].sum(1), torch.ones(B, device=cv.device))
