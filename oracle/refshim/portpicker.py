"""TEST SHIM (import-time only): rl_learner.py imports portpicker at module level."""


def pick_unused_port():
    return 0
