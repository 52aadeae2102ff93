"""CUPTI trace (torch.profiler; nsys is not installed in this image) of (1) one initial_inference and (2) one fused
MuZeroMCTSCtree.search(): the kernel timeline with the idle gaps between kernels, and every CUDA runtime / driver API call made
inside search() -- the evidence for "one cudaGraphLaunch, no cudaStreamSynchronize / cudaMemcpy inside search()" (SURVEY 8d).
Run on the GPU box; writes gpurun_out/trace_step.json and prints a summary."""
import json
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lightzero_b200 as lzb
from lightzero_b200.synthetic_weights import synthetic_state_dict

B, S, A = 1024, 50, 18
model = lzb.MuZeroModel(observation_shape=(4, 84, 84), action_space_size=A).load_state_dict(synthetic_state_dict((4, 84, 84), A))
obs = torch.rand(B, 4, 84, 84).cuda()
mcts = lzb.MuZeroMCTSCtree(dict(num_simulations=S, deterministic=True, discount_factor=0.997))
noise = torch.from_numpy(np.random.default_rng(0).dirichlet([0.3] * A, size=B).astype(np.float32)).cuda()
mask = torch.ones(B, A, dtype=torch.uint8).cuda()      # device-resident legal-action mask: nothing to upload inside search()


def prep():
    out0 = model.initial_inference(obs)
    roots = mcts.roots(B, mask)
    roots.prepare(0.25, noise, None, out0.policy_logits, None)
    roots._materialize(S, mcts._params())
    return out0, roots


for _ in range(3):                       # warm-up: graph capture, allocations
    out0, roots = prep()
    mcts.search(roots, model, out0.latent_state, None)
torch.cuda.synchronize()


def events(prof):
    ev = []
    for e in prof.events():
        dt = str(getattr(e, "device_type", ""))
        ev.append(dict(name=e.name, cuda="CUDA" in dt, start_us=e.time_range.start, dur_us=e.time_range.end - e.time_range.start))
    return ev


summary = {}
# ---- (1) initial_inference: kernel timeline
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    model.initial_inference(obs)
    torch.cuda.synchronize()
ev = events(prof)
kern = sorted([e for e in ev if e["cuda"]], key=lambda e: e["start_us"])
rows, prev_end = [], None
for e in kern:
    gap = (e["start_us"] - prev_end) if prev_end is not None else 0.0
    rows.append(dict(kernel=e["name"][:60], dur_us=round(e["dur_us"], 1), gap_before_us=round(gap, 1)))
    prev_end = e["start_us"] + e["dur_us"]
span = (kern[-1]["start_us"] + kern[-1]["dur_us"] - kern[0]["start_us"]) if kern else 0.0
summary["initial_inference"] = dict(kernels=rows, busy_us=round(sum(e["dur_us"] for e in kern), 1), span_us=round(span, 1))
print("initial_inference: %d device activities, busy %.1f us, span %.1f us" % (len(kern), summary["initial_inference"]["busy_us"], span))
for r in rows:
    print("   %-60s %9.1f us   gap before %7.1f us" % (r["kernel"], r["dur_us"], r["gap_before_us"]))

# ---- (2) search(): runtime API calls between entering and leaving search()
out0, roots = prep()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    with torch.profiler.record_function("lz_search_call"):
        mcts.search(roots, model, out0.latent_state, None)
    torch.cuda.synchronize()            # outside search(): closes the trace
ev = events(prof)
rng = [e for e in ev if e["name"] == "lz_search_call"][0]
lo, hi = rng["start_us"], rng["start_us"] + rng["dur_us"]
api = {}
for e in ev:      # CUDA runtime / driver API calls issued between entering and leaving search()
    if not e["cuda"] and (e["name"].startswith("cuda") or e["name"].startswith("cu")) and lo <= e["start_us"] <= hi:
        api[e["name"]] = api.get(e["name"], 0) + 1
dev = sorted([e for e in ev if e["cuda"]], key=lambda e: e["start_us"])
summary["search"] = dict(runtime_api_calls=api, device_activities=[dict(name=e["name"][:60], dur_us=round(e["dur_us"], 1)) for e in dev])
print("search(): CUDA API calls:", api)
print("search(): device activities:", [(e["name"][:40], round(e["dur_us"], 1)) for e in dev])
n_sync = sum(v for k, v in api.items() if "Synchronize" in k)
n_cpy = sum(v for k, v in api.items() if "Memcpy" in k)
summary["search"]["syncs_inside_search"] = n_sync
summary["search"]["memcpys_inside_search"] = n_cpy
print("search(): cudaGraphLaunch x%d, synchronisations inside search(): %d, memcpy calls inside search(): %d"
      % (api.get("cudaGraphLaunch", 0), n_sync, n_cpy))
# ---- (3) one end-to-end collect step from pinned host uint8 frames: copies and kernels on one timeline
from lightzero_b200.collect import MuZeroCollectPolicy
policy = MuZeroCollectPolicy(model, dict(num_simulations=S, deterministic=True, discount_factor=0.997))
policy.h2d_chunks = int(os.environ.get("H2D_CHUNKS", 2))
h_u8 = torch.randint(0, 256, (B, 4, 84, 84), dtype=torch.uint8).pin_memory()
h_mask = torch.ones(B, A, dtype=torch.uint8).pin_memory()
h_noise = noise.cpu().pin_memory()
for _ in range(3):
    policy.search_batch(h_u8, h_mask, h_noise, None, deterministic=True, read_back=True)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    policy.search_batch(h_u8, h_mask, h_noise, None, deterministic=True, read_back=True)
    torch.cuda.synchronize()
ev = events(prof)
dev = sorted([e for e in ev if e["cuda"]], key=lambda e: e["start_us"])
t0 = dev[0]["start_us"]
rows = [dict(name=e["name"][:56], start_us=round(e["start_us"] - t0, 1), dur_us=round(e["dur_us"], 1)) for e in dev]
summary["e2e_u8"] = dict(activities=rows, span_us=round(dev[-1]["start_us"] + dev[-1]["dur_us"] - t0, 1))
print("end-to-end step (uint8 host frames, %d H2D chunks): span %.1f us" % (policy.h2d_chunks, summary["e2e_u8"]["span_us"]))
for r in rows:
    print("   %9.1f us  +%9.1f us  %s" % (r["start_us"], r["dur_us"], r["name"]))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(summary, open(os.path.join(ROOT, "gpurun_out", "trace_step.json"), "w"), indent=1)
