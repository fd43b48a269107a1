"""The ``distar/agent/b200`` pipeline: what a DI-star maintainer drops next to ``distar/agent/default``.

DI-star selects an agent pipeline by NAME: ``import_module(pipeline, 'RLLearner' | 'SLLearner' | 'Agent')`` imports
``distar.agent.<pipeline>.{rl_learner, sl_learner, agent}`` (distar/agent/import_helper.py:3-14), chosen by ``learner.agent``
/ ``actor.agents`` in the user yaml (bin/rl_user_config.yaml:27,135-137).  Copy (or symlink) this directory to
``distar/agent/b200`` and set ``learner.agent: 'b200'``; or call ``register()`` below, which mounts it under that module name
without touching the DI-star tree.  Everything here subclasses / re-exports the reference's own classes, swapping only the
members on the hot path: the model, the loss, the optimiser + gradient clip, the DP wrapper.
"""
import importlib
import sys

MODULES = ('agent', 'rl_learner', 'sl_learner', 'model', 'model.model', 'rl_training', 'rl_training.rl_loss', 'sl_training',
           'sl_training.sl_loss')


def register(name: str = 'b200') -> str:
    """Mount this package as ``distar.agent.<name>`` (idempotent).  Needs ``distar`` importable."""
    import distar.agent  # noqa: F401
    base = 'distar.agent.' + name
    if base not in sys.modules:
        pkg = importlib.import_module(__name__)
        sys.modules[base] = pkg
        setattr(sys.modules['distar.agent'], name, pkg)
    for m in MODULES:
        if base + '.' + m not in sys.modules:
            sys.modules[base + '.' + m] = importlib.import_module(__name__ + '.' + m)
    return base
