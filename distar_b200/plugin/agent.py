"""distar/agent/b200/agent.py — the stock ``Agent`` (observation parsing, action decoding, trajectory collection: all CPU /
game-protocol code, out of scope) computing with the B200 model.  ``Agent.__init__`` builds ``Model(cfg)`` from its module
global (distar/agent/default/agent.py:20,105,143); its calls ``compute_logp_action`` / ``compute_teacher_logit``
(agent.py:127,312,503,513,725,737) keep their signatures.  The stock module's source is executed a second time under this
pipeline's name with that one global rebound, so the ``default`` pipeline's own Agent (a league may mix pipelines per player,
bin/rl_user_config.yaml:135-137) keeps the reference model."""
import importlib.util

import distar.agent.default.agent as _default
from distar_b200.model import Model

_spec = importlib.util.spec_from_file_location(__name__ + '_impl', _default.__file__)
_impl = importlib.util.module_from_spec(_spec)
_impl.__package__ = _default.__package__            # the stock module's relative imports (.lib, .model, ...) resolve as before
_spec.loader.exec_module(_impl)
_impl.Model = Model
Agent = _impl.Agent
