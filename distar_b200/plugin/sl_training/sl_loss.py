"""distar/agent/b200/sl_training/sl_loss.py — drop-in for distar/agent/default/sl_training/sl_loss.py."""
from distar_b200.sl_loss import SupervisedLoss  # noqa: F401
