"""Policy/value networks of the data-generation path (PyTorch; reference architectures.py).

The three PyTorch families of the reference are provided with IDENTICAL parameter names and
shapes, so `state_dict`s are interchangeable with the reference's `<n>.pt` checkpoints:
  AlphaSame      (architectures.py:60-142, config :8-17)   — BASELINE's blocks=10 filters=16
  BaseResNet     (architectures.py:159-271, config :150-156)
  AuxBaseResNet  (architectures.py:279-353)                 — the Config default
Each also has `forward_packed(grids, extras)`, the device-resident entry used by self-play:
grids [2B,1,40,10] (side to move first, opponent second) and extras [B,105] exactly as the
feature-encode kernel writes them (csrc/features.cu), so no per-feature host tensors exist.
The Keras twins of the reference (architectures.py:356-718) are out of scope (TensorFlow).
"""
from dataclasses import dataclass

import torch
from torch import nn

from .const import COLS, MINOS, POLICY_SIZE, PREVIEWS, ROWS

PIECE_FEATS = (2 + PREVIEWS) * len(MINOS)  # 49 one-hot piece inputs per player
SCALAR_FEATS = 3                           # b2b, combo, garbage
SIDE_FEATS = PIECE_FEATS + SCALAR_FEATS    # 52
CELLS = ROWS * COLS


@dataclass
class AlphaSameConfig:
    blocks: int = 10
    pooling_blocks: int = 2
    filters: int = 16
    cpool: int = 4
    dropout: float = 0.25
    kernels: int = 1
    o_side_neurons: int = 16
    value_head_neurons: int = 16


@dataclass
class BaseResNetConfig:
    blocks: int = 8
    filters: int = 32
    opp_hidden: int = 128
    own_kernels: int = 4
    value_hidden: int = 16


@dataclass
class AuxBaseResNetConfig(BaseResNetConfig):
    aux_hidden: int = 16
    aux_weight: float = 1.5


def _conv(cin, cout, k):
    return nn.Conv2d(cin, cout, kernel_size=k, padding="same", bias=False)


def _side(pieces, b2b, combo, garbage):
    """[B,7,7] + three [B] scalars -> [B,52] in torch.cat order."""
    return torch.cat([pieces.flatten(1), b2b.unsqueeze(1), combo.unsqueeze(1), garbage.unsqueeze(1)], dim=1)


def pack_inputs(a_grid, a_pieces, a_b2b, a_combo, a_garbage, o_grid, o_pieces, o_b2b, o_combo, o_garbage, color):
    """The reference's 11 batched inputs -> (grids [2B,1,40,10], extras [B,105])."""
    grids = torch.cat([a_grid, o_grid], dim=0)
    extras = torch.cat([_side(a_pieces, a_b2b, a_combo, a_garbage), _side(o_pieces, o_b2b, o_combo, o_garbage),
                        color.unsqueeze(1)], dim=1)
    return grids, extras.to(grids.dtype)


class ResidualBlock(nn.Module):
    """Pre-activation block: BN-ReLU-Conv3x3, BN-Dropout-ReLU-Conv3x3, + skip."""

    def __init__(self, model_config):
        super().__init__()
        f = model_config.filters
        self.conv_block1 = nn.Sequential(nn.BatchNorm2d(f), nn.ReLU(), _conv(f, f, 3))
        self.conv_block2 = nn.Sequential(nn.BatchNorm2d(f), nn.Dropout(p=model_config.dropout), nn.ReLU(), _conv(f, f, 3))

    def forward(self, x):
        return x + self.conv_block2(self.conv_block1(x))


class AlphaSame(nn.Module):
    def __init__(self, model_config=None, use_tanh=False):
        super().__init__()
        cfg = model_config or AlphaSameConfig()
        self.conv1 = _conv(1, cfg.filters, 5)
        self.res_blocks = nn.Sequential(*[ResidualBlock(cfg) for _ in range(cfg.blocks)])
        self.batchnorm1 = nn.BatchNorm2d(cfg.filters)
        self.relu1 = nn.ReLU()
        self.kernel1 = _conv(cfg.filters, cfg.kernels, 1)
        self.batchnorm2 = nn.BatchNorm2d(cfg.kernels)
        self.relu2 = nn.ReLU()
        self.flatten1 = nn.Flatten()
        grid_out = CELLS * cfg.kernels
        self.osidedense = nn.Sequential(nn.Linear(grid_out, cfg.o_side_neurons),
                                        nn.BatchNorm1d(cfg.o_side_neurons), nn.ReLU())
        head_in = grid_out + cfg.o_side_neurons + 2 * SIDE_FEATS + 1
        self.policy_head = nn.Linear(head_in, POLICY_SIZE)
        self.value_head = nn.Sequential(
            nn.Linear(head_in, cfg.value_head_neurons), nn.BatchNorm1d(cfg.value_head_neurons), nn.ReLU(),
            nn.Linear(cfg.value_head_neurons, 1), nn.Dropout(cfg.dropout), nn.Tanh() if use_tanh else nn.Sigmoid())

    def grid_features(self, grids):
        x = self.res_blocks(self.conv1(grids))
        x = self.kernel1(self.relu1(self.batchnorm1(x)))
        return self.flatten1(self.relu2(self.batchnorm2(x)))

    def forward_packed(self, grids, extras):
        b = extras.shape[0]
        if self.training:
            # the reference runs process_grid on the two boards separately (architectures.py:128-129): BatchNorm
            # sees two batches of B (statistics and running-stat updates), not one of 2B
            feats = torch.cat([self.grid_features(grids[:b]), self.grid_features(grids[b:])], dim=0)
        else:
            feats = self.grid_features(grids)
        x = torch.cat([feats[:b], extras[:, :SIDE_FEATS], self.osidedense(feats[b:]), extras[:, SIDE_FEATS:]], dim=1)
        return self.value_head(x), self.policy_head(x)

    def forward(self, a_grid, a_pieces, a_b2b, a_combo, a_garbage, o_grid, o_pieces, o_b2b, o_combo, o_garbage, color):
        return self.forward_packed(*pack_inputs(a_grid, a_pieces, a_b2b, a_combo, a_garbage,
                                                o_grid, o_pieces, o_b2b, o_combo, o_garbage, color))


class _BaseResBlock(nn.Module):
    """Post-activation block: Conv-BN-ReLU-Conv-BN, + skip, ReLU."""

    def __init__(self, filters):
        super().__init__()
        self.conv1, self.bn1 = _conv(filters, filters, 3), nn.BatchNorm2d(filters)
        self.conv2, self.bn2 = _conv(filters, filters, 3), nn.BatchNorm2d(filters)

    def forward(self, x):
        y = self.bn2(self.conv2(torch.relu(self.bn1(self.conv1(x)))))
        return torch.relu(x + y)


class BaseResNet(nn.Module):
    """Shared trunk on both boards; the opponent summary becomes a per-channel bias on the own
    feature map (FiLM-add) before the 1x1 collapse that feeds the heads."""

    def __init__(self, model_config=None, use_tanh=False):
        super().__init__()
        cfg = model_config or BaseResNetConfig()
        f = cfg.filters
        self.stem = nn.Sequential(_conv(1, f, 3), nn.BatchNorm2d(f), nn.ReLU())
        self.trunk = nn.Sequential(*[_BaseResBlock(f) for _ in range(cfg.blocks)])
        self.opp_collapse = nn.Sequential(nn.Conv2d(f, 1, kernel_size=1, bias=False), nn.BatchNorm2d(1), nn.ReLU())
        self.opp_encode = nn.Sequential(nn.Linear(CELLS + SIDE_FEATS, cfg.opp_hidden),
                                        nn.BatchNorm1d(cfg.opp_hidden), nn.ReLU())
        self.bias_project = nn.Linear(cfg.opp_hidden + SIDE_FEATS + 1, f)
        self.own_collapse = nn.Sequential(nn.Conv2d(f, cfg.own_kernels, kernel_size=1, bias=False),
                                          nn.BatchNorm2d(cfg.own_kernels), nn.ReLU())
        self._head_in = cfg.own_kernels * CELLS + SIDE_FEATS + 1
        self.policy_head = nn.Linear(self._head_in, POLICY_SIZE)
        self.value_head = nn.Sequential(nn.Linear(self._head_in, cfg.value_hidden), nn.BatchNorm1d(cfg.value_hidden),
                                        nn.ReLU(), nn.Linear(cfg.value_hidden, 1),
                                        nn.Tanh() if use_tanh else nn.Sigmoid())

    def _process_grid(self, grid):
        return self.trunk(self.stem(grid))

    def head_input(self, grids, extras):
        b = extras.shape[0]
        if self.training:   # two BatchNorm batches like the reference (architectures.py:236-237), see AlphaSame
            feats = torch.cat([self._process_grid(grids[:b]), self._process_grid(grids[b:])], dim=0)
        else:
            feats = self._process_grid(grids)
        own, opp, color = extras[:, :SIDE_FEATS], extras[:, SIDE_FEATS:2 * SIDE_FEATS], extras[:, 2 * SIDE_FEATS:]
        opp_repr = self.opp_encode(torch.cat([self.opp_collapse(feats[b:]).flatten(1), opp], dim=1))
        bias = self.bias_project(torch.cat([opp_repr, own, color], dim=1))
        flat = self.own_collapse(feats[:b] + bias[:, :, None, None]).flatten(1)
        return torch.cat([flat, own, color], dim=1)

    def forward_packed(self, grids, extras):
        x = self.head_input(grids, extras)
        return self.value_head(x), self.policy_head(x)

    def forward(self, a_grid, a_pieces, a_b2b, a_combo, a_garbage, o_grid, o_pieces, o_b2b, o_combo, o_garbage, color):
        return self.forward_packed(*pack_inputs(a_grid, a_pieces, a_b2b, a_combo, a_garbage,
                                                o_grid, o_pieces, o_b2b, o_combo, o_garbage, color))


def compute_aux_targets(own_grid_batch):
    """(B,1,40,10) 0/1 grids -> (B,2) [holes, aggregate height], both / 400 (architectures.py:285-305)."""
    filled = (own_grid_batch.squeeze(1) > 0.5).float()
    covered = (torch.cumsum(filled, dim=1) > 0).float()
    holes = ((1.0 - filled) * covered).sum(dim=(1, 2))
    top = filled.argmax(dim=1)
    heights = torch.where(filled.sum(dim=1) > 0, (ROWS - top).float(), torch.zeros_like(top, dtype=torch.float))
    return torch.stack([holes, heights.sum(dim=1)], dim=1) / float(CELLS)


class AuxBaseResNet(BaseResNet):
    """BaseResNet + a 2-output sigmoid head (holes, height of the own board); inference ignores it."""

    def __init__(self, model_config=None, use_tanh=False):
        cfg = model_config or AuxBaseResNetConfig()
        super().__init__(cfg, use_tanh=use_tanh)
        self.aux_head = nn.Sequential(nn.Linear(self._head_in, cfg.aux_hidden), nn.BatchNorm1d(cfg.aux_hidden),
                                      nn.ReLU(), nn.Linear(cfg.aux_hidden, 2), nn.Sigmoid())

    def forward_packed(self, grids, extras, with_aux=False):
        x = self.head_input(grids, extras)
        if with_aux:
            return self.value_head(x), self.policy_head(x), self.aux_head(x)
        return self.value_head(x), self.policy_head(x)

    def forward(self, a_grid, a_pieces, a_b2b, a_combo, a_garbage, o_grid, o_pieces, o_b2b, o_combo, o_garbage, color):
        return self.forward_packed(*pack_inputs(a_grid, a_pieces, a_b2b, a_combo, a_garbage,
                                                o_grid, o_pieces, o_b2b, o_combo, o_garbage, color), with_aux=True)


def build_network(model_config, use_tanh=False):
    """Type of `model_config` selects the family (reference dispatch, ai.py:1047-1054)."""
    if isinstance(model_config, AuxBaseResNetConfig):
        return AuxBaseResNet(model_config, use_tanh)
    if isinstance(model_config, BaseResNetConfig):
        return BaseResNet(model_config, use_tanh)
    return AlphaSame(model_config, use_tanh)
