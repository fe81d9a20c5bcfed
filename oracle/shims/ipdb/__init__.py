"""Oracle shim: the reference imports ipdb.set_trace at module scope (lib/model/sde.py:6)."""


def set_trace(*a, **k):
    return None
