"""distar/agent/b200/rl_training/rl_loss.py — drop-in for distar/agent/default/rl_training/rl_loss.py."""
from distar_b200.rl_loss import ReinforcementLoss  # noqa: F401
