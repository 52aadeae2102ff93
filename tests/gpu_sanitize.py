"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python tests/gpu_sanitize.py
Sizes are tiny: the sanitizer slows kernels 10-100x."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lightzero_b200 as lzb
from lightzero_b200.collect import MuZeroCollectPolicy
from oracle.model_ref import MuZeroModelRef, emulate_trained_, MuZeroModelMLPRef, emulate_trained_mlp_

torch.manual_seed(0)
A, B, S = 6, 9, 6
ref = emulate_trained_(MuZeroModelRef((4, 84, 84), A), 0)
for math in ("tc3", "fp32"):
    cu = lzb.MuZeroModel(observation_shape=(4, 84, 84), action_space_size=A).load_state_dict(ref.state_dict()).set_math(math)
    pol = MuZeroCollectPolicy(cu, dict(num_simulations=S, deterministic=True, discount_factor=0.997))
    obs = torch.rand(B, 4, 84, 84)
    mask = (np.random.default_rng(0).random((B, A)) < 0.7).astype(np.uint8); mask[:, 0] = 1
    noise = np.random.default_rng(1).dirichlet([0.3] * A, size=B).astype(np.float32)
    r = pol.search_batch(obs.pin_memory(), mask, noise, None)
    assert int(r["visits"].clamp(min=0).sum()) == B * S
    r2 = pol.search_batch(obs.cuda(), mask, noise, None, deterministic=False)
    assert int(r2["visits"].clamp(min=0).sum()) == B * S
    print(math, "ok", r["values"][:3].tolist())
mref = emulate_trained_mlp_(MuZeroModelMLPRef(4, 2), 0)
mcu = lzb.MuZeroModelMLP(4, 2, res_connection_in_dynamics=True).load_state_dict(mref.state_dict())
mcts = lzb.MuZeroMCTSCtree(dict(num_simulations=5, deterministic=True))
o = mcu.initial_inference((torch.rand(5, 4) * 2 - 1).cuda())
roots = mcts.roots(5, [[0, 1]] * 5)
roots.prepare_no_noise([0.] * 5, o.policy_logits, [1, 2, 1, 2, 1])     # two-player sign flips
mcts.search(roots, mcu, o.latent_state, [1, 2, 1, 2, 1])
print("mlp ok", roots.get_distributions(), roots.get_trajectories()[:2])
# EfficientZero: value-prefix trees, LSTM value-prefix head, multi-kernel search graph, collect from host buffers
from lightzero_b200.collect import EfficientZeroCollectPolicy
from oracle.model_ref import EfficientZeroModelRef
eref = emulate_trained_(EfficientZeroModelRef((4, 96, 96), A), 1)
ecu = lzb.EfficientZeroModel(observation_shape=(4, 96, 96), action_space_size=A).load_state_dict(eref.state_dict())
epol = EfficientZeroCollectPolicy(ecu, dict(num_simulations=8, discount_factor=0.997, lstm_horizon_len=2))
eobs = torch.rand(B, 4, 96, 96)
r = epol.search_batch(eobs.pin_memory(), mask, noise, None)
assert int(r["visits"].clamp(min=0).sum()) == B * 8
emcts = lzb.EfficientZeroMCTSCtree(dict(num_simulations=8, lstm_horizon_len=2))
eo = ecu.initial_inference(eobs.cuda())
eroots = emcts.roots(B, [list(range(A))] * B)
eroots.prepare(0.25, noise, [0.] * B, eo.policy_logits, [1, 2] * (B // 2) + [1])     # two-player branch
emcts.search(eroots, ecu, eo.latent_state, eo.reward_hidden_state, [1, 2] * (B // 2) + [1])
print("efficientzero ok", r["values"][:3].tolist(), eroots.get_distributions()[:2])
torch.cuda.synchronize()
print("sanitize script finished")
