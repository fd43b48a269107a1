"""distar/agent/b200/model/model.py — same constructor, methods and state_dict keys as distar/agent/default/model/model.py."""
from distar_b200.model import Model  # noqa: F401
