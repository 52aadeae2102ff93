"""ding.torch_utils.{MLP, ResBlock} (DI-engine v0.5.x nn_module.py / res_block.py semantics)."""
import torch
import torch.nn as nn

from .network.normalization import build_normalization


def conv2d_block(in_channels, out_channels, kernel_size, stride=1, padding=0, activation=None, norm_type=None, bias=True):
    block = [nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=bias)]
    if norm_type is not None:
        block.append(build_normalization(norm_type, dim=2)(out_channels))
    if activation is not None:
        block.append(activation)
    return nn.Sequential(*block)


class ResBlock(nn.Module):
    def __init__(self, in_channels, activation=nn.ReLU(), norm_type='BN', res_type='basic', bias=True, out_channels=None):
        super().__init__()
        self.act = activation
        assert res_type in ['basic', 'bottleneck', 'downsample']
        self.res_type = res_type
        if out_channels is None:
            out_channels = in_channels
        if res_type == 'basic':
            self.conv1 = conv2d_block(in_channels, out_channels, 3, 1, 1, activation=self.act, norm_type=norm_type, bias=bias)
            self.conv2 = conv2d_block(out_channels, out_channels, 3, 1, 1, activation=None, norm_type=norm_type, bias=bias)
        elif res_type == 'bottleneck':
            self.conv1 = conv2d_block(in_channels, out_channels, 1, 1, 0, activation=self.act, norm_type=norm_type, bias=bias)
            self.conv2 = conv2d_block(out_channels, out_channels, 3, 1, 1, activation=self.act, norm_type=norm_type, bias=bias)
            self.conv3 = conv2d_block(out_channels, out_channels, 1, 1, 0, activation=None, norm_type=norm_type, bias=bias)
        else:
            self.conv1 = conv2d_block(in_channels, out_channels, 3, 2, 1, activation=self.act, norm_type=norm_type, bias=bias)
            self.conv2 = conv2d_block(out_channels, out_channels, 3, 1, 1, activation=None, norm_type=norm_type, bias=bias)
            self.conv3 = conv2d_block(in_channels, out_channels, 3, 2, 1, activation=None, norm_type=None, bias=bias)

    def forward(self, x):
        identity = x
        x = self.conv1(x)
        x = self.conv2(x)
        if self.res_type == 'bottleneck':
            x = self.conv3(x)
        elif self.res_type == 'downsample':
            identity = self.conv3(identity)
        x = self.act(x + identity)
        return x


def MLP(in_channels, hidden_channels, out_channels, layer_num, layer_fn=None, activation=None, norm_type=None,
        use_dropout=False, dropout_probability=0.5, output_activation=True, output_norm=True,
        last_linear_layer_init_zero=False):
    assert layer_num >= 0
    if layer_num == 0:
        return nn.Sequential(nn.Identity())
    channels = [in_channels] + [hidden_channels] * (layer_num - 1) + [out_channels]
    if layer_fn is None:
        layer_fn = nn.Linear
    block = []
    for i, (cin, cout) in enumerate(zip(channels[:-2], channels[1:-1])):
        block.append(layer_fn(cin, cout))
        if norm_type is not None:
            block.append(build_normalization(norm_type, dim=1)(cout))
        if activation is not None:
            block.append(activation)
        if use_dropout:
            block.append(nn.Dropout(dropout_probability))
    block.append(layer_fn(channels[-2], channels[-1]))
    if output_norm and norm_type is not None:
        block.append(build_normalization(norm_type, dim=1)(channels[-1]))
    if output_activation and activation is not None:
        block.append(activation)
        if use_dropout:
            block.append(nn.Dropout(dropout_probability))
    if last_linear_layer_init_zero:
        for layer in reversed(block):
            if isinstance(layer, nn.Linear):
                nn.init.zeros_(layer.weight)
                nn.init.zeros_(layer.bias)
                break
    return nn.Sequential(*block)
