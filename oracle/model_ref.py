"""oracle/model_ref.py -- plain-PyTorch fp32 restatement of the reference MuZero conv model.

TEST INFRASTRUCTURE ONLY: the parity checker for the CUDA network kernels.  Never imported by
the product package ``lightzero_b200``.

Pinning: tests/golden/make_model_golden.py imports the reference's OWN model classes
(lzero/model/{muzero_model,efficientzero_model,muzero_model_mlp,common}.py from /root/reference, without executing
lzero/__init__.py) and asserts that every class below reproduces them BIT FOR BIT on seeded inputs with identical
state_dicts (initial_inference and recurrent_inference: value / reward or value_prefix / policy logits / latents / LSTM
state) before it writes the committed fixtures tests/golden/model_*.npz; tests/test_oracle_model.py re-checks the
restatement against those vectors anywhere.  The reference's two third-party imports that are absent here (DI-engine
``ding`` -- pinned ``DI-engine>=0.5.3`` in requirements.txt:1 -- and ``ditk``) are served by tests/golden/ding_stub, which
restates ``ding.torch_utils.{MLP,ResBlock}`` / ``build_normalization`` from their published v0.5.x semantics: those two
building blocks are the only part of the model arithmetic that is NOT executed from reference source ("parity unpinned"
for them; everything lzero/model does with them is pinned).  (4,84,84) observations: the reference MuZeroModel itself
raises for them (muzero_model.py:122-128 defines latent_size for 96 and 64 only); the restatement follows
sampled_muzero_model.py:144-145.  ``InverseScalarTransform`` is pinned against the reference's
``inverse_scalar_transform`` (policy/tests/test_scaling_transform.py:7-19).

DI-engine pieces restated from DI-engine v0.5.x ``ding/torch_utils/network/{res_block,nn_module}.py``
semantics: ``conv2d_block`` = nn.Sequential(Conv2d[, norm][, act]) and
``ResBlock`` basic / downsample.
"""
import math
from dataclasses import dataclass
from typing import List, Sequence

import torch
import torch.nn as nn


@dataclass
class MZNetworkOutput:
    """lzero/model/common.py:131-141"""
    value: torch.Tensor
    reward: torch.Tensor
    policy_logits: torch.Tensor
    latent_state: torch.Tensor


def conv2d_block(cin, cout, k, s, p, activation=None, norm=True, bias=False):
    layers = [nn.Conv2d(cin, cout, k, s, p, bias=bias)]
    if norm:
        layers.append(nn.BatchNorm2d(cout))
    if activation is not None:
        layers.append(activation)
    return nn.Sequential(*layers)


class ResBlock(nn.Module):
    """DI-engine ResBlock (res_type 'basic' / 'downsample'), as used at
    lzero/model/common.py:308-329,755-758,1130-1136 and muzero_model.py:475-481."""

    def __init__(self, in_channels, out_channels=None, res_type="basic", bias=False):
        super().__init__()
        out_channels = out_channels or in_channels
        self.res_type = res_type
        self.act = nn.ReLU(inplace=False)
        if res_type == "basic":
            self.conv1 = conv2d_block(in_channels, out_channels, 3, 1, 1, nn.ReLU(), True, bias)
            self.conv2 = conv2d_block(out_channels, out_channels, 3, 1, 1, None, True, bias)
        elif res_type == "downsample":
            self.conv1 = conv2d_block(in_channels, out_channels, 3, 2, 1, nn.ReLU(), True, bias)
            self.conv2 = conv2d_block(out_channels, out_channels, 3, 1, 1, None, True, bias)
            self.conv3 = conv2d_block(in_channels, out_channels, 3, 2, 1, None, False, bias)
        else:
            raise ValueError(res_type)

    def forward(self, x):
        identity = x
        x = self.conv1(x)
        x = self.conv2(x)
        if self.res_type == "downsample":
            identity = self.conv3(identity)
        return self.act(x + identity)


def MLP_V2(in_channels, hidden_channels: List[int], out_channels, last_linear_layer_init_zero=True):
    """lzero/model/common.py:28-99 with activation=ReLU, norm_type='BN', output_activation=False,
    output_norm=False (the only way the MuZero heads call it: muzero_model.py:493-502,
    common.py:1165-1187).  Sequential indices: 0 Linear, 1 BatchNorm1d, 2 ReLU, 3 Linear."""
    layers = []
    chans = [in_channels] + list(hidden_channels) + [out_channels]
    for i in range(len(chans) - 1):
        layers.append(nn.Linear(chans[i], chans[i + 1]))
        if i != len(chans) - 2:
            layers.append(nn.BatchNorm1d(chans[i + 1]))
            layers.append(nn.ReLU())
    if last_linear_layer_init_zero:
        nn.init.zeros_(layers[-1].weight)
        nn.init.zeros_(layers[-1].bias)
    return nn.Sequential(*layers)


class DownSample(nn.Module):
    """lzero/model/common.py:266-366"""

    def __init__(self, observation_shape, out_channels):
        super().__init__()
        self.observation_shape = observation_shape
        self.conv1 = nn.Conv2d(observation_shape[0], out_channels // 2, 3, 2, 1, bias=False)
        self.norm1 = nn.BatchNorm2d(out_channels // 2)
        self.resblocks1 = nn.ModuleList([ResBlock(out_channels // 2)])
        self.downsample_block = ResBlock(out_channels // 2, out_channels, res_type="downsample")
        self.resblocks2 = nn.ModuleList([ResBlock(out_channels)])
        self.pooling1 = nn.AvgPool2d(3, 2, 1)
        self.resblocks3 = nn.ModuleList([ResBlock(out_channels)])
        self.pooling2 = nn.AvgPool2d(3, 2, 1)

    def forward(self, x):
        x = torch.relu(self.norm1(self.conv1(x)))
        for b in self.resblocks1:
            x = b(x)
        x = self.downsample_block(x)
        for b in self.resblocks2:
            x = b(x)
        x = self.pooling1(x)
        for b in self.resblocks3:
            x = b(x)
        h = self.observation_shape[1]
        if h == 64:
            return x
        if h in (84, 96):
            return self.pooling2(x)
        raise NotImplementedError(h)


class RepresentationNetwork(nn.Module):
    """lzero/model/common.py:706-787 (downsample=True, use_sim_norm=False)"""

    def __init__(self, observation_shape, num_res_blocks, num_channels):
        super().__init__()
        self.downsample_net = DownSample(observation_shape, num_channels)
        self.resblocks = nn.ModuleList([ResBlock(num_channels) for _ in range(num_res_blocks)])

    def forward(self, x):
        x = self.downsample_net(x)
        for b in self.resblocks:
            x = b(x)
        return x


class DynamicsNetwork(nn.Module):
    """lzero/model/muzero_model.py:419-538.  `num_channels` here is the LATENT channel count; the conv
    input has num_channels + action_encoding_dim planes (muzero_model.py:158,465)."""

    def __init__(self, action_encoding_dim, num_res_blocks, num_channels, reward_head_channels,
                 reward_head_hidden_channels, output_support_size, flatten_size, last_zero=True):
        super().__init__()
        self.action_encoding_dim = action_encoding_dim
        self.conv = nn.Conv2d(num_channels + action_encoding_dim, num_channels, 3, 1, 1, bias=False)
        self.norm_common = nn.BatchNorm2d(num_channels)
        self.resblocks = nn.ModuleList([ResBlock(num_channels) for _ in range(num_res_blocks)])
        self.conv1x1_reward = nn.Conv2d(num_channels, reward_head_channels, 1)
        self.norm_reward = nn.BatchNorm2d(reward_head_channels)
        self.fc_reward_head = MLP_V2(flatten_size, reward_head_hidden_channels, output_support_size, last_zero)

    def forward(self, state_action_encoding):
        state_encoding = state_action_encoding[:, :-self.action_encoding_dim, :, :]
        x = self.norm_common(self.conv(state_action_encoding))
        x = torch.relu(x + state_encoding)
        for b in self.resblocks:
            x = b(x)
        next_latent_state = x
        x = torch.relu(self.norm_reward(self.conv1x1_reward(next_latent_state)))
        x = x.reshape(x.shape[0], -1)
        return next_latent_state, self.fc_reward_head(x)


class PredictionNetwork(nn.Module):
    """lzero/model/common.py:1081-1215"""

    def __init__(self, action_space_size, num_res_blocks, num_channels, value_head_channels,
                 policy_head_channels, value_hidden, policy_hidden, output_support_size,
                 flat_value, flat_policy, last_zero=True):
        super().__init__()
        self.resblocks = nn.ModuleList([ResBlock(num_channels) for _ in range(num_res_blocks)])
        self.conv1x1_value = nn.Conv2d(num_channels, value_head_channels, 1)
        self.conv1x1_policy = nn.Conv2d(num_channels, policy_head_channels, 1)
        self.norm_value = nn.BatchNorm2d(value_head_channels)
        self.norm_policy = nn.BatchNorm2d(policy_head_channels)
        self.fc_value = MLP_V2(flat_value, value_hidden, output_support_size, last_zero)
        self.fc_policy = MLP_V2(flat_policy, policy_hidden, action_space_size, last_zero)

    def forward(self, latent_state):
        for b in self.resblocks:
            latent_state = b(latent_state)
        value = torch.relu(self.norm_value(self.conv1x1_value(latent_state)))
        policy = torch.relu(self.norm_policy(self.conv1x1_policy(latent_state)))
        value = value.reshape(value.shape[0], -1)
        policy = policy.reshape(policy.shape[0], -1)
        return self.fc_policy(policy), self.fc_value(value)


def latent_hw(h: int) -> int:
    """Spatial size of the latent for downsample=True: 96->6, 84->6 (ceil(84/14), the fix the
    reference applies in sampled_muzero_model.py:144-145; muzero_model.py:122-125 only defines 96/64), 64->8."""
    if h == 96:
        return math.ceil(h / 16)
    if h == 84:
        return math.ceil(h / 14)
    if h == 64:
        return math.ceil(h / 8)
    raise NotImplementedError(h)


class MuZeroModelRef(nn.Module):
    """lzero/model/muzero_model.py:20-272 (conv, downsample=True, BN, one_hot, categorical)."""

    def __init__(self, observation_shape: Sequence[int] = (4, 84, 84), action_space_size: int = 6,
                 num_res_blocks: int = 1, num_channels: int = 64, reward_head_channels: int = 16,
                 value_head_channels: int = 16, policy_head_channels: int = 16,
                 reward_head_hidden_channels=(32,), value_head_hidden_channels=(32,),
                 policy_head_hidden_channels=(32,), reward_support_range=(-300., 301., 1.),
                 value_support_range=(-300., 301., 1.), last_linear_layer_init_zero: bool = True):
        super().__init__()
        self.action_space_size = action_space_size
        self.reward_support_size = len(torch.arange(*reward_support_range))
        self.value_support_size = len(torch.arange(*value_support_range))
        hw = latent_hw(observation_shape[1])
        latent_size = hw * hw
        self.latent_hw = hw
        self.representation_network = RepresentationNetwork(observation_shape, num_res_blocks, num_channels)
        self.dynamics_network = DynamicsNetwork(
            action_space_size, num_res_blocks, num_channels, reward_head_channels,
            list(reward_head_hidden_channels), self.reward_support_size,
            reward_head_channels * latent_size, last_linear_layer_init_zero)
        self.prediction_network = PredictionNetwork(
            action_space_size, num_res_blocks, num_channels, value_head_channels, policy_head_channels,
            list(value_head_hidden_channels), list(policy_head_hidden_channels), self.value_support_size,
            value_head_channels * latent_size, policy_head_channels * latent_size,
            last_linear_layer_init_zero)

    def initial_inference(self, obs):
        """muzero_model.py:210-240"""
        latent_state = self.representation_network(obs)
        policy_logits, value = self.prediction_network(latent_state)
        return MZNetworkOutput(value, [0. for _ in range(obs.size(0))], policy_logits, latent_state)

    def recurrent_inference(self, latent_state, action):
        """muzero_model.py:242-272 with _dynamics one-hot encoding (:331-374)"""
        if action.dim() == 1:
            action = action.unsqueeze(-1)
        one_hot = torch.zeros(action.shape[0], self.action_space_size, device=action.device)
        one_hot.scatter_(1, action.long(), 1)
        enc = one_hot.unsqueeze(-1).unsqueeze(-1).expand(
            latent_state.shape[0], self.action_space_size, latent_state.shape[2], latent_state.shape[3])
        next_latent_state, reward = self.dynamics_network(torch.cat((latent_state, enc), dim=1))
        policy_logits, value = self.prediction_network(next_latent_state)
        return MZNetworkOutput(value, reward, policy_logits, next_latent_state)


@dataclass
class EZNetworkOutput:
    """lzero/model/common.py:119-128"""
    value: torch.Tensor
    value_prefix: torch.Tensor
    policy_logits: torch.Tensor
    latent_state: torch.Tensor
    reward_hidden_state: tuple


class DynamicsNetworkEZ(nn.Module):
    """lzero/model/efficientzero_model.py:427-570: the MuZero dynamics trunk, then conv1x1_reward -> BN -> ReLU ->
    flatten -> nn.LSTM(flatten, lstm_hidden) (one step) -> BatchNorm1d -> ReLU -> DI-engine MLP(lstm_hidden -> hidden ->
    support) = Linear, BN, ReLU, Linear (efficientzero_model.py:515-525)."""

    def __init__(self, action_encoding_dim, num_res_blocks, num_channels, reward_head_channels,
                 reward_head_hidden_channels, output_support_size, flatten_size, lstm_hidden_size=512, last_zero=True):
        super().__init__()
        self.action_encoding_dim = action_encoding_dim
        self.flatten_size = flatten_size
        self.conv = nn.Conv2d(num_channels + action_encoding_dim, num_channels, 3, 1, 1, bias=False)
        self.norm_common = nn.BatchNorm2d(num_channels)
        self.resblocks = nn.ModuleList([ResBlock(num_channels) for _ in range(num_res_blocks)])
        self.conv1x1_reward = nn.Conv2d(num_channels, reward_head_channels, 1)
        self.norm_reward = nn.BatchNorm2d(reward_head_channels)
        self.lstm = nn.LSTM(input_size=flatten_size, hidden_size=lstm_hidden_size)
        self.norm_value_prefix = nn.BatchNorm1d(lstm_hidden_size)
        self.fc_reward_head = DingMLP(lstm_hidden_size, reward_head_hidden_channels[0], output_support_size,
                                      len(reward_head_hidden_channels) + 1, output_activation=False, output_norm=False)
        if last_zero:
            last = [l for l in self.fc_reward_head if isinstance(l, nn.Linear)][-1]
            nn.init.zeros_(last.weight)
            nn.init.zeros_(last.bias)

    def forward(self, state_action_encoding, reward_hidden_state):
        state_encoding = state_action_encoding[:, :-self.action_encoding_dim, :, :]
        x = self.norm_common(self.conv(state_action_encoding))
        x = x + state_encoding
        x = torch.relu(x)
        for b in self.resblocks:
            x = b(x)
        next_latent_state = x
        x = torch.relu(self.norm_reward(self.conv1x1_reward(next_latent_state)))
        x = x.reshape(-1, self.flatten_size).unsqueeze(0)
        value_prefix, next_reward_hidden_state = self.lstm(x, reward_hidden_state)
        value_prefix = torch.relu(self.norm_value_prefix(value_prefix.squeeze(0)))
        return next_latent_state, next_reward_hidden_state, self.fc_reward_head(value_prefix)


class EfficientZeroModelRef(nn.Module):
    """lzero/model/efficientzero_model.py:20-272 (conv, downsample=True, BN, one_hot, categorical, state_norm=False);
    the SSL projection heads (:184-201) are training-only and not restated."""

    def __init__(self, observation_shape: Sequence[int] = (4, 96, 96), action_space_size: int = 6,
                 num_res_blocks: int = 1, num_channels: int = 64, lstm_hidden_size: int = 512,
                 reward_head_channels: int = 16, value_head_channels: int = 16, policy_head_channels: int = 16,
                 reward_head_hidden_channels=(32,), value_head_hidden_channels=(32,),
                 policy_head_hidden_channels=(32,), reward_support_range=(-300., 301., 1.),
                 value_support_range=(-300., 301., 1.), last_linear_layer_init_zero: bool = True):
        super().__init__()
        self.action_space_size = action_space_size
        self.lstm_hidden_size = lstm_hidden_size
        self.reward_support_size = len(torch.arange(*reward_support_range))
        self.value_support_size = len(torch.arange(*value_support_range))
        hw = latent_hw(observation_shape[1])
        latent_size = hw * hw
        self.latent_hw = hw
        self.representation_network = RepresentationNetwork(observation_shape, num_res_blocks, num_channels)
        self.dynamics_network = DynamicsNetworkEZ(
            action_space_size, num_res_blocks, num_channels, reward_head_channels,
            list(reward_head_hidden_channels), self.reward_support_size,
            reward_head_channels * latent_size, lstm_hidden_size, last_linear_layer_init_zero)
        self.prediction_network = PredictionNetwork(
            action_space_size, num_res_blocks, num_channels, value_head_channels, policy_head_channels,
            list(value_head_hidden_channels), list(policy_head_hidden_channels), self.value_support_size,
            value_head_channels * latent_size, policy_head_channels * latent_size,
            last_linear_layer_init_zero)

    def initial_inference(self, obs):
        """efficientzero_model.py:203-238: zero reward hidden state (h, c)"""
        B = obs.size(0)
        latent_state = self.representation_network(obs)
        policy_logits, value = self.prediction_network(latent_state)
        hidden = (torch.zeros(1, B, self.lstm_hidden_size, device=obs.device), torch.zeros(1, B, self.lstm_hidden_size, device=obs.device))
        return EZNetworkOutput(value, [0. for _ in range(B)], policy_logits, latent_state, hidden)

    def recurrent_inference(self, latent_state, reward_hidden_state, action):
        """efficientzero_model.py:240-272 with _dynamics one-hot encoding (:310-381)"""
        if action.dim() == 1:
            action = action.unsqueeze(-1)
        one_hot = torch.zeros(action.shape[0], self.action_space_size, device=action.device)
        one_hot.scatter_(1, action.long(), 1)
        enc = one_hot.unsqueeze(-1).unsqueeze(-1).expand(
            latent_state.shape[0], self.action_space_size, latent_state.shape[2], latent_state.shape[3])
        next_latent_state, hidden, value_prefix = self.dynamics_network(torch.cat((latent_state, enc), dim=1), reward_hidden_state)
        policy_logits, value = self.prediction_network(next_latent_state)
        return EZNetworkOutput(value, value_prefix, policy_logits, next_latent_state, hidden)


def emulate_trained_(model: nn.Module, seed: int = 0) -> nn.Module:
    """Make a randomly initialised model non-degenerate (BASELINE.md s.3 / SURVEY.md 8d config 3):
    the reference zero-initialises the last Linear of every head (common.py:91-97) so all logits
    would be 0 and every PUCT score ties.  Re-draw those layers N(0, 0.02) and randomise the BN
    running statistics / affine parameters, all under `seed`."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
                m.weight.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
        for seq in _head_mlps(model):
            last = [l for l in seq if isinstance(l, nn.Linear)][-1]
            last.weight.copy_(torch.randn(last.weight.shape, generator=g) * 0.02)
            last.bias.copy_(torch.randn(last.bias.shape, generator=g) * 0.02)
    model.eval()
    return model


def _head_mlps(model):
    out = []
    for name in ("fc_reward_head", "fc_value", "fc_policy"):
        for m in model.modules():
            if hasattr(m, name):
                out.append(getattr(m, name))
    return out


class DiscreteSupport:
    """lzero/policy/scaling_transform.py:6-13"""

    def __init__(self, start, stop, step=1., device="cpu"):
        self.arange = torch.arange(start, stop, step, dtype=torch.float32).unsqueeze(0).to(device)
        self.size = self.arange.shape[1]
        self.step = step


class InverseScalarTransform:
    """lzero/policy/scaling_transform.py:64-92"""

    def __init__(self, scalar_support: DiscreteSupport, categorical_distribution: bool = True):
        self.value_support = scalar_support.arange
        self.categorical_distribution = categorical_distribution

    def __call__(self, logits, epsilon: float = 0.001):
        if self.categorical_distribution:
            value_probs = torch.softmax(logits, dim=1)
            value = value_probs.mul_(self.value_support).sum(1, keepdim=True)
        else:
            value = logits
        tmp = ((torch.sqrt(1 + 4 * epsilon * (torch.abs(value) + 1 + epsilon)) - 1) / (2 * epsilon))
        return torch.sign(value) * (tmp * tmp - 1)


def inverse_scalar_transform(logits, scalar_support, epsilon=0.001, categorical_distribution=True):
    """lzero/policy/scaling_transform.py:33-61 (the function form the reference test compares with)"""
    if categorical_distribution:
        value_probs = torch.softmax(logits, dim=1)
        value = (scalar_support.arange.to(value_probs.device) * value_probs).sum(1, keepdim=True)
    else:
        value = logits
    return torch.sign(value) * (
        ((torch.sqrt(1 + 4 * epsilon * (torch.abs(value) + 1 + epsilon)) - 1) / (2 * epsilon)) ** 2 - 1)


# --------------------------------------------------------------------------------------------------
# MuZeroModelMLP (vector observations; BASELINE config 1: CartPole, latent 128, A=2)
# --------------------------------------------------------------------------------------------------
def DingMLP(in_channels, hidden_channels, out_channels, layer_num, output_activation=True, output_norm=True):
    """DI-engine ``ding.torch_utils.MLP`` with norm_type='BN', activation=ReLU as called at
    muzero_model_mlp.py:367-404: ``layer_num`` Linear layers [in] + [hidden]*(layer_num-1) + [out], each
    followed by BatchNorm1d + ReLU (last layer per output_norm / output_activation).  Sequential indices
    for layer_num=2: 0 Linear, 1 BN, 2 ReLU, 3 Linear, 4 BN, 5 ReLU."""
    chans = [in_channels] + [hidden_channels] * (layer_num - 1) + [out_channels]
    block = []
    for i in range(layer_num):
        last = i == layer_num - 1
        block.append(nn.Linear(chans[i], chans[i + 1]))
        if (not last) or output_norm:
            block.append(nn.BatchNorm1d(chans[i + 1]))
        if (not last) or output_activation:
            block.append(nn.ReLU())
    return nn.Sequential(*block)


def MLP_V2_general(in_channels, hidden_channels, out_channels, activation, output_activation, output_norm,
                   last_linear_layer_init_zero):
    """lzero/model/common.py:28-99 with norm_type='BN'."""
    layers = []
    chans = [in_channels] + list(hidden_channels) + [out_channels]
    n = len(chans) - 1
    for i in range(n):
        layers.append(nn.Linear(chans[i], chans[i + 1]))
        if i != n - 1:
            layers.append(nn.BatchNorm1d(chans[i + 1]))
            layers.append(activation())
        else:
            if output_norm:
                layers.append(nn.BatchNorm1d(chans[i + 1]))
            if output_activation:
                layers.append(activation())
    if last_linear_layer_init_zero:
        last = [l for l in layers if isinstance(l, nn.Linear)][-1]
        nn.init.zeros_(last.weight)
        nn.init.zeros_(last.bias)
    return nn.Sequential(*layers)


class RepresentationNetworkMLP(nn.Module):
    """lzero/model/common.py:790-850: Linear -> BN1d -> GELU(tanh) -> Linear(zero-init) -> LayerNorm
    (MuZeroModelMLP passes no activation, so the default nn.GELU(approximate='tanh') applies,
    muzero_model_mlp.py:104-106)."""

    def __init__(self, observation_shape, hidden_channels):
        super().__init__()
        self.fc_representation = MLP_V2_general(observation_shape, [hidden_channels], hidden_channels,
                                                lambda: nn.GELU(approximate='tanh'), False, False, True)
        self.norm = nn.LayerNorm(hidden_channels)

    def forward(self, x):
        return self.norm(self.fc_representation(x.float()))


class DynamicsNetworkMLP(nn.Module):
    """lzero/model/muzero_model_mlp.py:328-442"""

    def __init__(self, action_encoding_dim, latent_dim, reward_hidden, support_size, res_connection, last_zero=True):
        super().__init__()
        self.action_encoding_dim = action_encoding_dim
        self.res_connection_in_dynamics = res_connection
        nc = latent_dim + action_encoding_dim
        if res_connection:
            self.fc_dynamics_1 = DingMLP(nc, latent_dim, latent_dim, 2)
            self.fc_dynamics_2 = DingMLP(latent_dim, latent_dim, latent_dim, 2)
        else:
            self.fc_dynamics = DingMLP(nc, latent_dim, latent_dim, 2)
        self.fc_reward_head = MLP_V2_general(latent_dim, list(reward_hidden), support_size, nn.ReLU, False, False, last_zero)

    def forward(self, sa):
        if self.res_connection_in_dynamics:
            latent = sa[:, :-self.action_encoding_dim]
            nxt = self.fc_dynamics_1(sa) + latent
            enc = self.fc_dynamics_2(nxt)
        else:
            nxt = self.fc_dynamics(sa)
            enc = nxt
        return nxt, self.fc_reward_head(enc)


class PredictionNetworkMLP(nn.Module):
    """lzero/model/common.py:1218-1292"""

    def __init__(self, action_space_size, latent_dim, value_hidden, policy_hidden, support_size, last_zero=True):
        super().__init__()
        self.fc_prediction_common = MLP_V2_general(latent_dim, [latent_dim], latent_dim, nn.ReLU, True, True, False)
        self.fc_value_head = MLP_V2_general(latent_dim, list(value_hidden), support_size, nn.ReLU, False, False, last_zero)
        self.fc_policy_head = MLP_V2_general(latent_dim, list(policy_hidden), action_space_size, nn.ReLU, False, False, last_zero)

    def forward(self, latent):
        x = self.fc_prediction_common(latent)
        return self.fc_policy_head(x), self.fc_value_head(x)


class MuZeroModelMLPRef(nn.Module):
    """lzero/model/muzero_model_mlp.py:21-295 (one_hot actions, categorical, BN, state_norm=False)."""

    def __init__(self, observation_shape: int = 4, action_space_size: int = 2, latent_state_dim: int = 128,
                 reward_head_hidden_channels=(32,), value_head_hidden_channels=(32,), policy_head_hidden_channels=(32,),
                 reward_support_range=(-300., 301., 1.), value_support_range=(-300., 301., 1.),
                 res_connection_in_dynamics: bool = True, last_linear_layer_init_zero: bool = True):
        super().__init__()
        self.action_space_size = action_space_size
        self.latent_state_dim = latent_state_dim
        self.reward_support_size = len(torch.arange(*reward_support_range))
        self.value_support_size = len(torch.arange(*value_support_range))
        self.representation_network = RepresentationNetworkMLP(observation_shape, latent_state_dim)
        self.dynamics_network = DynamicsNetworkMLP(action_space_size, latent_state_dim, reward_head_hidden_channels,
                                                   self.reward_support_size, res_connection_in_dynamics,
                                                   last_linear_layer_init_zero)
        self.prediction_network = PredictionNetworkMLP(action_space_size, latent_state_dim, value_head_hidden_channels,
                                                       policy_head_hidden_channels, self.value_support_size,
                                                       last_linear_layer_init_zero)

    def initial_inference(self, obs):
        latent = self.representation_network(obs)
        policy_logits, value = self.prediction_network(latent)
        return MZNetworkOutput(value, [0. for _ in range(obs.size(0))], policy_logits, latent)

    def recurrent_inference(self, latent_state, action):
        if action.dim() == 1:
            action = action.unsqueeze(-1)
        one_hot = torch.zeros(action.shape[0], self.action_space_size, device=action.device)
        one_hot.scatter_(1, action.long(), 1)
        nxt, reward = self.dynamics_network(torch.cat((latent_state, one_hot.float()), dim=1))
        policy_logits, value = self.prediction_network(nxt)
        return MZNetworkOutput(value, reward, policy_logits, nxt)


def emulate_trained_mlp_(model: nn.Module, seed: int = 0) -> nn.Module:
    """Same idea as emulate_trained_: un-zero every zero-initialised last layer, randomise norm statistics."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, nn.BatchNorm1d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
                m.weight.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
            if isinstance(m, nn.LayerNorm):
                m.weight.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
            if isinstance(m, nn.Linear) and float(m.weight.abs().sum()) == 0.0:
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * 0.05)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)
    model.eval()
    return model
