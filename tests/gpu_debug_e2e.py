"""Why does the chunked H2D not overlap?  Times search_batch from pinned host memory with different chunk counts."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lightzero_b200 as lzb
from lightzero_b200.collect import MuZeroCollectPolicy
from oracle.model_ref import MuZeroModelRef, emulate_trained_

B, A, S = 1024, 18, 50
torch.manual_seed(0)
ref = emulate_trained_(MuZeroModelRef((4, 84, 84), A), 0)
cu = lzb.MuZeroModel(observation_shape=(4, 84, 84), action_space_size=A).load_state_dict(ref.state_dict())
pol = MuZeroCollectPolicy(cu, dict(num_simulations=S, deterministic=True, discount_factor=0.997))
h_obs = torch.rand(B, 4, 84, 84).pin_memory()
d_obs = h_obs.cuda()
mask = torch.ones(B, A, dtype=torch.uint8).pin_memory()
noise = torch.rand(B, A).pin_memory()
print("pinned:", h_obs.is_pinned(), mask.is_pinned(), noise.is_pinned())

def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (time.perf_counter() - t0) * 1e3 / n

d = torch.empty_like(d_obs)
print("torch H2D copy of 115 MB: dev %.3f ms wall %.3f ms" % timeit(lambda: d.copy_(h_obs, non_blocking=True)))
print("device-resident step:     dev %.3f ms wall %.3f ms" % timeit(lambda: pol.search_batch(d_obs, mask.cuda(), noise.cuda(), None, deterministic=True, read_back=False)))
for ch in (1, 2, 4, 8):
    pol.h2d_chunks = ch
    t0 = time.perf_counter()
    pol.search_batch(h_obs, mask, noise, None, deterministic=True, read_back=False)
    call_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    print("host step, chunks=%d:      dev %.3f ms wall %.3f ms  (launch-call returns after %.3f ms)" % (ch, *timeit(lambda: pol.search_batch(h_obs, mask, noise, None, deterministic=True, read_back=True)), call_ms))

print("---- same under a non-default (non-blocking) torch stream")
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for ch in (1, 2, 4):
        pol.h2d_chunks = ch
        for _ in range(2):
            pol.search_batch(h_obs, mask, noise, None, deterministic=True, read_back=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(5):
            pol.search_batch(h_obs, mask, noise, None, deterministic=True, read_back=True)
        b.record(st)
        torch.cuda.synchronize()
        print("side stream, chunks=%d: dev %.3f ms" % (ch, a.elapsed_time(b) / 5))
