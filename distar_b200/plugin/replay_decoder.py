"""distar/agent/b200/replay_decoder.py — ``ReplayDecoder`` is looked up per pipeline (import_helper.py:3-8, bin/sl_train.py:47-48);
replay parsing is game-protocol CPU code outside the hot path, so the pipeline hands out the reference's own class."""
from distar.agent.default.replay_decoder import ReplayDecoder  # noqa: F401
