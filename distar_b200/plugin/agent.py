"""distar/agent/b200/agent.py — the stock ``Agent`` (observation parsing, action decoding, trajectory collection: all CPU /
game-protocol code, out of scope) computing with the B200 model.  ``Agent.__init__`` builds ``Model(cfg)`` from the module
global (distar/agent/default/agent.py:20,105,143): rebinding that global is the whole integration; its calls
``compute_logp_action`` / ``compute_teacher_logit`` (agent.py:127,312,503,513,725,737) keep their signatures."""
import distar.agent.default.agent as _default
from distar_b200.model import Model

_default.Model = Model
Agent = _default.Agent
