"""TEST SHIM (import-time only): rl_learner.py / learner_comm.py import Flask at module level."""


class Flask:
    def __init__(self, *a, **k):
        pass

    def route(self, *a, **k):
        return lambda f: f

    def run(self, *a, **k):
        pass


class _Req:
    json = {}


request = _Req()
