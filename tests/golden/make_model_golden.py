"""Pin the PyTorch model restatements (oracle/model_ref.py) to the REFERENCE'S OWN model classes and generate the
committed model fixtures tests/golden/model_*.npz.

Run in the build container (needs /root/reference):   python tests/golden/make_model_golden.py

The reference files lzero/model/{muzero_model,efficientzero_model,muzero_model_mlp,common,utils}.py are imported from
where they lie, WITHOUT executing lzero/__init__.py (which pulls the whole framework): `lzero` and `lzero.model` are
registered as bare namespace modules.  Their two missing third-party imports (DI-engine `ding`, `ditk`) are served by
tests/golden/ding_stub (see its README).  For each case the restated model is built under a fixed seed, its state_dict
is loaded into the reference class (strict on every key the restatement has), both run the same seeded inputs and must
agree BIT FOR BIT; the reference's outputs are then stored.  tests/test_oracle_model.py rebuilds the restatement under
the same seed on any machine and checks it against the stored vectors (and against a checksum of the weights).
"""
import hashlib
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("LZ_REFERENCE", "/root/reference")


def import_reference_models():
    sys.path.insert(0, os.path.join(HERE, "ding_stub"))
    for name, path in (("lzero", "lzero"), ("lzero.model", "lzero/model")):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, path)]
        sys.modules[name] = m
    return (importlib.import_module("lzero.model.muzero_model"), importlib.import_module("lzero.model.efficientzero_model"),
            importlib.import_module("lzero.model.muzero_model_mlp"))


def weights_digest(model) -> str:
    h = hashlib.sha256()
    for k, v in sorted(model.state_dict().items()):
        if v.dtype.is_floating_point:
            h.update(k.encode())
            h.update(v.detach().cpu().numpy().tobytes())
    return h.hexdigest()


def build_restated(kind, obs_shape, A, nres, seed):
    from oracle import model_ref as R
    torch.manual_seed(seed)
    if kind == "muzero":
        return R.emulate_trained_(R.MuZeroModelRef(tuple(obs_shape), A, num_res_blocks=nres), seed)
    if kind == "efficientzero":
        return R.emulate_trained_(R.EfficientZeroModelRef(tuple(obs_shape), A, num_res_blocks=nres), seed)
    if kind == "muzero_mlp":
        return R.emulate_trained_mlp_(R.MuZeroModelMLPRef(int(obs_shape), A), seed)
    raise ValueError(kind)


def case_inputs(kind, obs_shape, A, B, seed):
    g = torch.Generator().manual_seed(1000 + seed)
    if kind == "muzero_mlp":
        obs = torch.rand(B, int(obs_shape), generator=g) * 2 - 1
    else:
        obs = torch.rand((B,) + tuple(obs_shape), generator=g)
    action = torch.randint(0, A, (B,), generator=g)
    return obs, action, g


CASES = [
    # name,                     kind,            obs,          A,  res blocks, B, seed
    # (4, 84, 84) cannot be pinned this way: the reference MuZeroModel raises UnboundLocalError for it
    # (muzero_model.py:122-128 defines latent_size for 96 and 64 only); the restatement follows sampled_muzero_model.py:144-145
    ("model_muzero_96_a6",      "muzero",        (4, 96, 96),   6, 1, 4, 0),
    ("model_muzero_96_a18",     "muzero",        (4, 96, 96),  18, 1, 3, 1),
    ("model_muzero_96_a6_r2",   "muzero",        (4, 96, 96),   6, 2, 2, 2),
    ("model_ez_96_a6",          "efficientzero", (4, 96, 96),   6, 1, 4, 3),
    ("model_ez_96_a18",         "efficientzero", (4, 96, 96),  18, 1, 3, 4),
    ("model_mlp_cartpole",      "muzero_mlp",    4,             2, 1, 8, 5),
]


def run_case(mods, name, kind, obs_shape, A, nres, B, seed):
    mz_mod, ez_mod, mlp_mod = mods
    mine = build_restated(kind, obs_shape, A, nres, seed)
    if kind == "muzero":
        theirs = mz_mod.MuZeroModel(observation_shape=obs_shape, action_space_size=A, num_res_blocks=nres, downsample=True)
    elif kind == "efficientzero":
        theirs = ez_mod.EfficientZeroModel(observation_shape=obs_shape, action_space_size=A, num_res_blocks=nres, downsample=True)
    else:
        theirs = mlp_mod.MuZeroModelMLP(observation_shape=obs_shape, action_space_size=A, latent_state_dim=128,
                                        res_connection_in_dynamics=True)
    missing, unexpected = theirs.load_state_dict(mine.state_dict(), strict=False)
    assert not unexpected, unexpected                      # every restated key exists in the reference model
    assert all(k.startswith(("projection", "prediction_head")) for k in missing), missing   # only the training-time SSL heads
    theirs.eval()
    obs, action, g = case_inputs(kind, obs_shape, A, B, seed)
    out = {}
    with torch.no_grad():
        a0, b0 = theirs.initial_inference(obs), mine.initial_inference(obs)
        for f in ("value", "policy_logits", "latent_state"):
            assert torch.equal(getattr(a0, f), getattr(b0, f)), (name, "initial", f)
            out["init_" + f] = getattr(a0, f).numpy()
        latent = a0.latent_state * 1.0
        if kind == "efficientzero":
            hc = (torch.randn(1, B, 512, generator=g) * 0.3, torch.randn(1, B, 512, generator=g) * 0.3)
            a1, b1 = theirs.recurrent_inference(latent, hc, action), mine.recurrent_inference(latent, hc, action)
            for f in ("value", "value_prefix", "policy_logits", "latent_state"):
                assert torch.equal(getattr(a1, f), getattr(b1, f)), (name, "recurrent", f)
                out["rec_" + f] = getattr(a1, f).numpy()
            for i in range(2):
                assert torch.equal(a1.reward_hidden_state[i], b1.reward_hidden_state[i]), (name, "hidden", i)
                out[f"rec_hidden{i}"] = a1.reward_hidden_state[i].numpy()
                out[f"in_hidden{i}"] = hc[i].numpy()
        else:
            a1, b1 = theirs.recurrent_inference(latent, action), mine.recurrent_inference(latent, action)
            for f in ("value", "reward", "policy_logits", "latent_state"):
                assert torch.equal(getattr(a1, f), getattr(b1, f)), (name, "recurrent", f)
                out["rec_" + f] = getattr(a1, f).numpy()
    out.update(obs=obs.numpy(), action=action.numpy().astype(np.int64), kind=kind, A=A, nres=nres, seed=seed,
               obs_shape=np.asarray(obs_shape), weights_sha256=weights_digest(mine))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "restatement == reference class bit for bit; fixture written")


def main():
    mods = import_reference_models()
    for case in CASES:
        run_case(mods, *case)


if __name__ == "__main__":
    main()
