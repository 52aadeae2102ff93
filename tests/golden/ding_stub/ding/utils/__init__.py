"""ding.utils names used by lzero/model/*.py: registry decorator, SequenceType, rank helpers."""
from typing import List, Tuple, Union

SequenceType = Union[List, Tuple]


class _Registry:
    def register(self, name, *a, **k):
        return lambda cls: cls


MODEL_REGISTRY = _Registry()


def get_rank():
    return 0


def get_world_size():
    return 1


def set_pkg_seed(seed, use_cuda=True):
    pass
