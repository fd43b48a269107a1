"""Deterministic weight synthesis for tests / bench (no checkpoints exist offline).

Follows the reference's init *distributions* (``ctools/torch_utils/network/nn_module.py:17-46`` xavier-normal
fc/conv weights, torch-default Linear/Conv bias, ``torch.randn`` LSTM weights ``model/lstm.py:125-126``,
xavier-uniform learned embeddings ``scalar_encoder.py:76``, ``value_fc`` gain 0.1 ``nn_module.py:305-308``)
but draws every tensor from its own seeded CPU generator, so the same ``state_dict`` can be rebuilt on any
machine with this torch build and loaded into the reference, the oracle and the CUDA model alike.
``perturb`` jitters LayerNorm affine parameters away from (1, 0) so that tests exercise them.
"""
import math
import zlib
from typing import Dict

import torch

from .spec import param_specs


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device='cpu')
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
    return g


def _binary_table(rows: int, bits: int) -> torch.Tensor:
    v = torch.arange(rows).unsqueeze(1)
    s = torch.arange(bits - 1, -1, -1).unsqueeze(0)
    return ((v >> s) & 1).float()


def init_state_dict(seed: int = 0, spatial_x: int = 128, spatial_y: int = 128, baselines=('winloss',),
                    perturb: float = 0.1, use_value_feature: bool = False) -> Dict[str, torch.Tensor]:
    sd = {}
    for name, shape, kind in param_specs(spatial_x, spatial_y, baselines, use_value_feature):
        g = _gen(seed, name)
        if kind == 'xavier_normal':
            rf = 1
            for d in shape[2:]:
                rf *= d
            std = math.sqrt(2.0 / (shape[1] * rf + shape[0] * rf))
            t = torch.randn(shape, generator=g) * std
        elif kind.startswith('xavier_uniform'):
            gain = float(kind.split(':')[1]) if ':' in kind else 1.0
            a = gain * math.sqrt(6.0 / (shape[0] + shape[1]))
            t = (torch.rand(shape, generator=g) * 2 - 1) * a
        elif kind.startswith('bias:'):
            b = 1.0 / math.sqrt(int(kind.split(':')[1]))
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif kind.startswith('uniform:'):
            b = 1.0 / math.sqrt(int(kind.split(':')[1]))
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif kind == 'randn':
            t = torch.randn(shape, generator=g)
        elif kind == 'ln_weight':
            t = 1.0 + perturb * torch.randn(shape, generator=g)
        elif kind == 'ln_bias':
            t = perturb * torch.randn(shape, generator=g)
        elif kind.startswith('const:'):
            t = torch.full(shape, float(kind.split(':')[1]))
        elif kind == 'zeros':
            t = torch.zeros(shape)
        elif kind == 'frozen_eye':
            t = torch.eye(shape[0])
        elif kind == 'frozen_binary':
            t = _binary_table(shape[0], shape[1])
        elif kind == 'position_array':
            i = torch.arange(0, shape[0], dtype=torch.float)
            t = 1.0 / torch.pow(10000., (i // 2 * 2) / shape[0])   # scalar_encoder.py:11-16
        else:
            raise KeyError(kind)
        sd[name] = t.contiguous()
    return sd
