"""Random, non-degenerate weights in the reference's ``state_dict`` key layout, for benchmarking and
smoke runs where no checkpoint is available (there is no network for checkpoints; bench.py says ``data: synthetic``).

Key names / shapes follow ``MuZeroModel`` / ``EfficientZeroModel`` (lzero/model/muzero_model.py:140-184, 465-502;
common.py:300-332, 755-758, 1130-1187; efficientzero_model.py:511-525) with DI-engine's ``ResBlock`` convention
(``conv1.0.weight`` / ``conv1.1.{weight,bias,running_mean,running_var}``).  Like BASELINE.md's recipe the last Linear of
every head is drawn N(0, 0.02) instead of the reference's zero init (all-zero logits would make every PUCT score tie) and the
BatchNorm statistics are randomised, so the network behaves like a trained one for timing purposes.
"""
import math
from typing import Dict, Sequence

import torch


def synthetic_state_dict(observation_shape: Sequence[int] = (4, 84, 84), action_space_size: int = 18, num_res_blocks: int = 1,
                         num_channels: int = 64, head_channels: int = 16, head_hidden: int = 32, support_size: int = 601,
                         efficientzero: bool = False, lstm_hidden_size: int = 512, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    C, c2, A = num_channels, num_channels // 2, action_space_size
    P = 36 if observation_shape[-1] in (84, 96) else 64          # latent pixels (6x6 / 8x8)

    def conv(name, cout, cin, k=3, bias=False):
        bound = 1.0 / math.sqrt(cin * k * k)                       # nn.Conv2d default (kaiming_uniform, a = sqrt(5))
        sd[name + ".weight"] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
        if bias:
            sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    def bn(name, n):
        sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(n, generator=g)
        sd[name + ".bias"] = 0.1 * torch.randn(n, generator=g)
        sd[name + ".running_mean"] = 0.1 * torch.randn(n, generator=g)
        sd[name + ".running_var"] = torch.rand(n, generator=g) + 0.5

    def linear(name, nout, nin, std=None):
        if std is None:
            bound = 1.0 / math.sqrt(nin)
            sd[name + ".weight"] = (torch.rand(nout, nin, generator=g) * 2 - 1) * bound
            sd[name + ".bias"] = (torch.rand(nout, generator=g) * 2 - 1) * bound
        else:
            sd[name + ".weight"] = torch.randn(nout, nin, generator=g) * std
            sd[name + ".bias"] = torch.randn(nout, generator=g) * std

    def resblock(name, cin, cout, downsample=False):
        conv(name + ".conv1.0", cout, cin); bn(name + ".conv1.1", cout)
        conv(name + ".conv2.0", cout, cout); bn(name + ".conv2.1", cout)
        if downsample:
            conv(name + ".conv3.0", cout, cin)

    def head(name, nin, nout):
        linear(name + ".0", head_hidden, nin); bn(name + ".1", head_hidden); linear(name + ".3", nout, head_hidden, std=0.02)

    R = "representation_network.downsample_net."
    conv(R + "conv1", c2, observation_shape[0]); bn(R + "norm1", c2)
    resblock(R + "resblocks1.0", c2, c2)
    resblock(R + "downsample_block", c2, C, downsample=True)
    resblock(R + "resblocks2.0", C, C)
    resblock(R + "resblocks3.0", C, C)
    D, Q = "dynamics_network.", "prediction_network."
    conv(D + "conv", C, C + A); bn(D + "norm_common", C)
    for i in range(num_res_blocks):
        resblock(f"representation_network.resblocks.{i}", C, C)
        resblock(f"{D}resblocks.{i}", C, C)
        resblock(f"{Q}resblocks.{i}", C, C)
    conv(D + "conv1x1_reward", head_channels, C, k=1, bias=True); bn(D + "norm_reward", head_channels)
    conv(Q + "conv1x1_value", head_channels, C, k=1, bias=True); bn(Q + "norm_value", head_channels)
    conv(Q + "conv1x1_policy", head_channels, C, k=1, bias=True); bn(Q + "norm_policy", head_channels)
    nflat = head_channels * P
    if efficientzero:
        H = lstm_hidden_size
        k = 1.0 / math.sqrt(H)                                    # nn.LSTM default init
        for n, shape in (("weight_ih_l0", (4 * H, nflat)), ("weight_hh_l0", (4 * H, H)), ("bias_ih_l0", (4 * H,)), ("bias_hh_l0", (4 * H,))):
            sd[D + "lstm." + n] = (torch.rand(*shape, generator=g) * 2 - 1) * k
        bn(D + "norm_value_prefix", H)
        head(D + "fc_reward_head", H, support_size)
    else:
        head(D + "fc_reward_head", nflat, support_size)
    head(Q + "fc_value", nflat, support_size)
    head(Q + "fc_policy", nflat, A)
    return sd
