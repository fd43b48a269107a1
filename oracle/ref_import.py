"""Import the UNMODIFIED reference (/root/reference) in the authoring container.

TEST INFRASTRUCTURE ONLY.  Used by `oracle/make_golden.py` and by the `-m "not gpu"`
tests that pin the restatement in `oracle/alphastar_ref.py` against the real reference.
`/root/reference` does not exist on the GPU box, so nothing that runs there may call this.

Recipe follows SURVEY.md Appendix D: six import shims, no source edits.
"""
import math
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('DISTAR_REFERENCE_ROOT', '/root/reference')
_SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'refshim')


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'distar'))


def install_shims():
    import numpy as np
    from unittest import mock
    if _SHIM_DIR not in sys.path:
        sys.path[:0] = [_SHIM_DIR, REFERENCE_ROOT]
    if 'torch._six' not in sys.modules:
        m = types.ModuleType('torch._six')
        m.inf = math.inf
        m.string_classes = (str,)
        sys.modules['torch._six'] = m
    for a, b in [('int', int), ('float', float), ('bool', bool), ('object', object)]:
        if not hasattr(np, a):
            setattr(np, a, b)

    class _M(types.ModuleType):
        def __getattr__(self, n):
            if n.startswith('__'):
                raise AttributeError(n)
            v = mock.MagicMock(name=self.__name__ + '.' + n)
            setattr(self, n, v)
            return v

    names = ['s2clientprotocol'] + ['s2clientprotocol.%s_pb2' % x for x in (
        'sc2api', 'raw', 'common', 'spatial', 'ui', 'error', 'debug', 'data', 'score', 'query')]
    for name in names:
        if name not in sys.modules:
            sys.modules[name] = _M(name)
            if '.' in name:
                setattr(sys.modules['s2clientprotocol'], name.split('.')[1], sys.modules[name])


def load_reference(spatial=128, enable_baselines=('winloss',), seed=0, use_value_feature=False):
    """Returns (model, cfg, modules) with modules = dict(F=features, ReinforcementLoss=..., SupervisedLoss=...)."""
    assert reference_available(), 'reference tree not mounted'
    install_shims()
    import torch
    from distar.agent.default.model import Model
    from distar.agent.default.lib import features as F
    from distar.agent.default.rl_training.rl_loss import ReinforcementLoss
    from distar.agent.default.rl_training import as_rl_utils
    from distar.ctools.utils import read_config
    cfg = read_config(os.path.join(REFERENCE_ROOT, 'distar/bin/rl_user_config.yaml'))
    cfg.common.type = 'rl'
    cfg.learner.use_value_feature = bool(use_value_feature)
    cfg.learner.player_id = 'MP0'
    sx, sy = (spatial, spatial) if isinstance(spatial, int) else spatial          # (spatial_x, spatial_y), e.g. (160, 152)
    cfg.model.spatial_x, cfg.model.spatial_y = sx, sy
    cfg.model.enable_baselines = list(enable_baselines)
    F.SPATIAL_SIZE[:] = [sy, sx]
    torch.manual_seed(seed)
    model = Model(cfg, use_value_network=True)
    mods = dict(F=F, ReinforcementLoss=ReinforcementLoss, as_rl_utils=as_rl_utils, Model=Model)
    try:
        from distar.agent.default.sl_training.sl_loss import SupervisedLoss
        mods['SupervisedLoss'] = SupervisedLoss
    except Exception as e:  # pragma: no cover
        mods['SupervisedLoss_error'] = repr(e)
    return model, cfg, mods
