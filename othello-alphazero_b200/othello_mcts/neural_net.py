"""Host-side mirror of the reference network interface, for tests and benchmarks that must run where
/root/reference does not exist (reference: python/othello_alphazero/neural_net.py:138-172).

Same constructor, same forward contract and the same parameter names (so reference checkpoints load
with `load_state_dict`), written from the architecture description: ConvBlock -> R residual blocks ->
policy head (1x1 conv 2ch, BN, ReLU, FC 128->65, softmax) + value head (1x1 conv 1ch, BN, ReLU,
FC 64->Hv, ReLU, FC Hv->1, tanh).  This is the fp32 torch reference the native evaluator is compared to.
"""
from __future__ import annotations

import torch
from torch import nn


def _conv_bn(cin: int, cout: int, k: int) -> tuple[nn.Conv2d, nn.BatchNorm2d]:
    return nn.Conv2d(cin, cout, k, padding="same" if k == 3 else 0), nn.BatchNorm2d(cout)


class _ConvBlock(nn.Module):
    def __init__(self, cin: int, cout: int) -> None:
        super().__init__()
        self.conv, self.norm = _conv_bn(cin, cout, 3)

    def forward(self, x):
        return torch.relu(self.norm(self.conv(x)))


class _ResidualBlock(nn.Module):
    def __init__(self, c: int) -> None:
        super().__init__()
        self.conv1, self.norm1 = _conv_bn(c, c, 3)
        self.conv2, self.norm2 = _conv_bn(c, c, 3)

    def forward(self, x):
        y = torch.relu(self.norm1(self.conv1(x)))
        return torch.relu(self.norm2(self.conv2(y)) + x)


class _PolicyHead(nn.Module):
    def __init__(self, c: int, squares: int, actions: int) -> None:
        super().__init__()
        self.conv, self.norm = _conv_bn(c, 2, 1)
        self.linear = nn.Linear(2 * squares, actions)

    def forward(self, x):
        x = torch.relu(self.norm(self.conv(x))).flatten(1)
        return torch.softmax(self.linear(x), dim=1)


class _ValueHead(nn.Module):
    def __init__(self, c: int, squares: int, hidden: int) -> None:
        super().__init__()
        self.conv, self.norm = _conv_bn(c, 1, 1)
        self.linear1 = nn.Linear(squares, hidden)
        self.linear2 = nn.Linear(hidden, 1)

    def forward(self, x):
        x = torch.relu(self.norm(self.conv(x))).flatten(1)
        return torch.tanh(self.linear2(torch.relu(self.linear1(x))).squeeze(1))


class AlphaZeroNet(nn.Module):
    def __init__(self, in_channels: int, num_squares: int = 64, num_actions: int = 65, conv_channels: int = 128,
                 num_residual_blocks: int = 9, value_head_hidden_channels: int = 128) -> None:
        super().__init__()
        self.conv_block = _ConvBlock(in_channels, conv_channels)
        self.residual_blocks = nn.Sequential(*(_ResidualBlock(conv_channels) for _ in range(num_residual_blocks)))
        self.policy_head = _PolicyHead(conv_channels, num_squares, num_actions)
        self.value_head = _ValueHead(conv_channels, num_squares, value_head_hidden_channels)

    def forward(self, x: torch.Tensor) -> dict[str, torch.Tensor]:
        x = self.residual_blocks(self.conv_block(x))
        return {"policy": self.policy_head(x), "value": self.value_head(x)}


def randomize_batchnorm(net: nn.Module, seed: int = 0) -> None:
    """Give every BatchNorm non-trivial running statistics and affine terms (a fresh module has
    mean 0 / var 1 / gamma 1 / beta 0, which would not exercise the folding)."""
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, nn.BatchNorm2d):
            with torch.no_grad():
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
                m.weight.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.1)
