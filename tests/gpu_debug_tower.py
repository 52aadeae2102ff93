"""clock64 phase stamps of the middle CTA of every tcgen05 tower layer (bring-up; run on the GPU box: python tests/gpu_debug_tower.py)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["LZ_TC_DEBUG"] = "1"
import lightzero_b200 as lzb
from lightzero_b200 import cabi
from lightzero_b200.synthetic_weights import synthetic_state_dict

B, A = int(os.environ.get("DBG_B", 1024)), 18
model = lzb.MuZeroModel(observation_shape=(4, 84, 84), action_space_size=A).load_state_dict(synthetic_state_dict((4, 84, 84), A))
lib = cabi.load()
obs = torch.rand(B, 4, 84, 84).cuda()
names = ["rb1.conv1 <32>", "rb1.conv2 <32> (+x, phase-split out)", "downsample conv1|conv3 <128> (stride 2)", "downsample conv2 <64> (+id)",
         "rb2.conv1 <64>", "rb2.conv2 <64> (+x)", "rb3.conv1 <64>", "rb3.conv2 <64> (+x)"]
for _ in range(2):
    model.initial_inference(obs)
torch.cuda.synchronize()
print(f"B = {B}: cycles of the middle CTA per layer: [start -> input band in shared memory] [MMA issue: 9 taps] [-> accumulators complete] [epilogue]")
for k in range(8):
    os.environ["LZ_CONV_DEBUG"] = str(k)
    for _ in range(8 - 0):          # the launch counter is global: 8 launches per tower, so one tower per setting keeps the phase
        pass
    model.initial_inference(obs)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 64)()
    cabi.check(lib.lz_debug_tc_stamps(buf), "stamps")
    s = list(buf)
    nt, grid = s[63] >> 32, s[63] & 0xffffffff
    print(f"  {names[k]:42s} grid {grid:5d} NT {nt}:  load {s[59] - s[58]:6d}  mma issue {s[60] - s[59]:6d}  acc wait {s[61] - s[60]:6d}  epilogue {s[62] - s[61]:6d}  total {s[62] - s[58]:6d}")
