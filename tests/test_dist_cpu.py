"""N>1 host logic on CPU: world_size-2 gloo processes shard a root batch and all-gather results."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    from lightzero_b200.dist import gather_search_results, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    G, A = 11, 5          # uneven split on purpose
    lo, hi = shard_range(G, rank, world)
    g = torch.Generator().manual_seed(0)
    full_v = torch.randint(0, 50, (G, A), generator=g, dtype=torch.int32)
    full_f = torch.rand(G, generator=g)
    v, f = gather_search_results(full_v[lo:hi].clone(), full_f[lo:hi].clone(), G)
    ok = torch.equal(v, full_v) and torch.equal(f, full_f)
    open(os.path.join(out_dir, f"ok{rank}"), "w").write(str(ok))
    dist.destroy_process_group()


def test_shard_range_covers_batch():
    from lightzero_b200.dist import shard_range
    for G in (1, 7, 1024, 1025):
        for W in (1, 2, 3, 8):
            spans = [shard_range(G, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == G
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_gloo_world2_gather(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok0").read() == "True" and open(tmp_path / "ok1").read() == "True"


def _seg_worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    from lightzero_b200.collector import gather_segments
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, T, A = 3, 4, 5
    g = torch.Generator().manual_seed(1)
    cv = torch.rand(world * B, T, A, generator=g)
    rv = torch.rand(world * B, T, generator=g)
    ln = torch.randint(0, T + 1, (world * B,), generator=g, dtype=torch.int32)
    sl = slice(rank * B, (rank + 1) * B)
    gcv, grv, gln = gather_segments(cv[sl].clone(), rv[sl].clone(), ln[sl].clone())
    ok = torch.equal(gcv, cv) and torch.equal(grv, rv) and torch.equal(gln, ln)
    open(os.path.join(out_dir, f"seg{rank}"), "w").write(str(ok))
    dist.destroy_process_group()


def test_gloo_world2_segment_all_gather(tmp_path):
    """The finished-segment all-gather of the collector: one collective on one packed buffer."""
    port = 31500 + os.getpid() % 2000
    mp.spawn(_seg_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "seg0").read() == "True" and open(tmp_path / "seg1").read() == "True"
