"""ding.torch_utils.network.normalization.build_normalization (DI-engine v0.5.x): norm type + dimensionality -> class."""
import torch.nn as nn


def build_normalization(norm_type: str, dim=None):
    if dim is None:
        key = norm_type
    elif norm_type in ('BN', 'IN'):
        key = norm_type + str(dim)
    elif norm_type in ('LN', 'SyncBN'):
        key = norm_type
    else:
        raise NotImplementedError(norm_type)
    table = {'BN1': nn.BatchNorm1d, 'BN2': nn.BatchNorm2d, 'LN': nn.LayerNorm, 'IN1': nn.InstanceNorm1d,
             'IN2': nn.InstanceNorm2d, 'SyncBN': nn.SyncBatchNorm}
    return table[key]
