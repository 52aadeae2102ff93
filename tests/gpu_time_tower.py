"""Times initial_inference (tower + tail) with CUDA events (A/B tool for the conv_tc variants)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lightzero_b200 as lzb
from lightzero_b200.synthetic_weights import synthetic_state_dict
B, A = int(os.environ.get("DBG_B", 1024)), 18
model = lzb.MuZeroModel(observation_shape=(4, 84, 84), action_space_size=A).load_state_dict(synthetic_state_dict((4, 84, 84), A))
obs = [torch.rand(B, 4, 84, 84).cuda() for _ in range(3)]
ms = []
for i in range(12):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); model.initial_inference(obs[i % 3]); b.record(); torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
print(f"initial_inference B={B}: min {min(ms[3:]):.3f} median {sorted(ms[3:])[len(ms[3:]) // 2]:.3f} ms")
