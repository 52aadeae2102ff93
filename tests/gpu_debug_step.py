"""Where does a collect step spend its time outside the search kernel?  CUDA-event timings of the pieces (bench size)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lightzero_b200 as lzb
from lightzero_b200 import cabi
from lightzero_b200.collect import MuZeroCollectPolicy
from lightzero_b200.synthetic_weights import synthetic_state_dict

B, S, A = 1024, 50, 18
dev = torch.device("cuda", 0)
model = lzb.MuZeroModel(observation_shape=(4, 84, 84), action_space_size=A).load_state_dict(synthetic_state_dict((4, 84, 84), A))
policy = MuZeroCollectPolicy(model, dict(num_simulations=S, deterministic=True, discount_factor=0.997))
lib = cabi.load()
obs = [torch.rand(B, 4, 84, 84, device=dev) for _ in range(3)]
mask = torch.ones(B, A, dtype=torch.uint8, device=dev)
noise = torch.from_numpy(np.random.default_rng(0).dirichlet([0.3] * A, size=B).astype(np.float32)).to(dev)


def timed(fn, n=10, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    evs = []
    t0 = time.perf_counter()
    for i in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(i); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    return sum(a.elapsed_time(b) for a, b in evs) / n, wall


print("initial_inference (python API)        dev %.3f ms  wall %.3f ms" % timed(lambda i: model.initial_inference(obs[i % 3])))
print("search_batch device obs, no read-back dev %.3f ms  wall %.3f ms" % timed(lambda i: policy.search_batch(obs[i % 3], mask, noise, None, deterministic=True, read_back=False)))
# back-to-back initial inferences in one event pair: launch gaps amortised?
def many(i):
    for j in range(4):
        model.initial_inference(obs[j % 3])
d, w = timed(many)
print("4 x initial_inference back to back    dev %.3f ms per call  wall %.3f ms per call" % (d / 4, w / 4))
