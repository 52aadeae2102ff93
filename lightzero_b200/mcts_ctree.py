"""Mirror of ``lzero.mcts.tree_search.mcts_ctree.MuZeroMCTSCtree`` (mcts_ctree.py:211-368): same
class-level ``config``, ``default_config()``, ``__init__(cfg)``, ``roots(n, legal_actions)`` and
``search(roots, model, latent_state_roots, to_play_batch, task_id=None)``.

``search`` has two execution modes, both entirely on the GPU:
  * fused (``model`` is a ``lightzero_b200.MuZeroModel``): the whole num_simulations loop is one CUDA
    graph launch (``lz_search_run``) -- zero host<->device synchronisations inside the search;
  * step-wise (any other object with ``recurrent_inference``, e.g. the user's own torch module on
    CUDA): the device trees are driven one simulation at a time around the caller's model; the tree
    hands the gather indices to torch on device, so this mode also never syncs with the host.

Parity contract: ``deterministic=True`` reproduces the reference C++ ctree bit for bit (first legal
action attaining the maximum, cnode.cpp:592); the reference's default mode seeds rand() from the wall
clock on every traverse (cnode.cpp:770, utils.cpp:12-26) and is not reproducible by construction, so
``deterministic=False`` draws from the same tie list with a counter-based device RNG instead.
"""
import copy
from typing import Any, List, Optional, Union

import numpy as np
import torch

from . import cabi, mz_tree
from .muzero_model import MuZeroModel
from .muzero_model_mlp import MuZeroModelMLP
from .scaling_transform import DiscreteSupport, InverseScalarTransform


class ConfigDict(dict):
    """Minimal attribute dict (the reference uses easydict.EasyDict, not installed here)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            v = ConfigDict(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def update(self, other=None, **kw):
        for k, v in dict(other or {}, **kw).items():
            self[k] = v


class MuZeroMCTSCtree(object):
    # mcts_ctree.py:220-232
    config = dict(
        root_dirichlet_alpha=0.3,
        root_noise_weight=0.25,
        pb_c_base=19652,
        pb_c_init=1.25,
        value_delta_max=0.01,
        env_type='not_board_games',
    )
    # extension: tie-breaking mode (the reference MuZero search passes no flag, i.e. False)
    deterministic_default = False

    @classmethod
    def default_config(cls) -> ConfigDict:
        cfg = ConfigDict(copy.deepcopy(cls.config))
        cfg.cfg_type = cls.__name__ + 'Dict'
        return cfg

    def __init__(self, cfg=None) -> None:
        default_config = self.default_config()
        default_config.update(cfg or {})
        self._cfg = default_config
        self._cfg.setdefault("num_simulations", 50)
        self._cfg.setdefault("discount_factor", 0.997)
        self._cfg.setdefault("device", "cuda")
        self.deterministic = bool(self._cfg.get("deterministic", self.deterministic_default))
        self._inv = None    # lazily: InverseScalarTransform for the step-wise mode (mcts_ctree.py:250-253)

    @classmethod
    def roots(cls, active_collect_env_num: int, legal_actions: List[Any]) -> "mz_tree.Roots":
        """mcts_ctree.py:255-265"""
        return mz_tree.Roots(active_collect_env_num, legal_actions)

    def _params(self):
        c = self._cfg
        return (c.pb_c_base, c.pb_c_init, c.discount_factor, c.value_delta_max)

    def search(self, roots: "mz_tree.Roots", model, latent_state_roots, to_play_batch: Union[int, List[Any]],
               task_id: Optional[int] = None) -> None:
        """mcts_ctree.py:267-368.  ``latent_state_roots``: np.ndarray or CUDA tensor [B,C,H,W]."""
        S = int(self._cfg.num_simulations)
        roots._materialize(S, self._params())     # reset + prepare on device, MinMax stats fresh (:291-292)
        t = roots._tree
        dev = roots.device
        if isinstance(latent_state_roots, torch.Tensor):
            lat = latent_state_roots.to(dev, torch.float32, non_blocking=True).contiguous()
        else:
            lat = torch.from_numpy(np.ascontiguousarray(latent_state_roots, dtype=np.float32)).to(dev, non_blocking=True)
        if isinstance(model, (MuZeroModel, MuZeroModelMLP)):
            q = t.search_for(model, S)
            with torch.cuda.device(dev):
                cabi.check(t.lib.lz_search_run(q, lat.data_ptr(), int(self.deterministic), cabi.stream_ptr()),
                           "lz_search_run")
            self.last_num_kernels = t.lib.lz_search_num_kernels(q)
            return
        self._search_stepwise(roots, model, lat, S)

    def _search_stepwise(self, roots, model, lat, S):
        t = roots._tree
        dev = roots.device
        if self._inv is None:
            m = self._cfg.get("model", None)
            rng = tuple(m.value_support_range) if m is not None and "value_support_range" in m else (-300., 301., 1.)
            self._inv = InverseScalarTransform(DiscreteSupport(*rng, device=dev))
        B = roots.num
        pool = torch.empty((S + 1,) + tuple(lat.shape), device=dev, dtype=torch.float32)
        pool[0] = lat
        rows = torch.arange(B, device=dev)
        with torch.no_grad(), torch.cuda.device(dev):
            if hasattr(model, "eval"):
                model.eval()
            for sim in range(S):
                cabi.check(t.lib.lz_tree_traverse(t.h, int(self.deterministic), t.ix.data_ptr(), t.iy.data_ptr(),
                                                  t.action.data_ptr(), t.search_len.data_ptr(), t.vtp.data_ptr(),
                                                  cabi.stream_ptr()), "lz_tree_traverse")
                latent_states = pool[t.ix.long(), rows]                       # mcts_ctree.py:323-324 on device
                out = model.recurrent_inference(latent_states, t.action.long())
                pool[sim + 1] = out.latent_state
                value = self._inv(out.value).reshape(-1).contiguous()        # :349
                reward = self._inv(out.reward).reshape(-1).contiguous()      # :350
                pol = out.policy_logits.to(torch.float32).contiguous()
                cabi.check(t.lib.lz_tree_backpropagate(t.h, sim + 1, reward.data_ptr(), value.data_ptr(),
                                                       pol.data_ptr(), None, cabi.stream_ptr()),
                           "lz_tree_backpropagate")
