"""Minimal stand-in for `accelerate` so the READ-ONLY reference at /root/reference imports
in this container (accelerate is not installed and there is no network).  TEST INFRASTRUCTURE
ONLY: used by oracle/ref_shim.py when generating golden fixtures.  Single-device CPU runs never
call these meaningfully."""
from contextlib import contextmanager

__version__ = "0.0.0-stub"


def dispatch_model(model, *a, **k):
    return model


def infer_auto_device_map(*a, **k):
    return {}


@contextmanager
def init_empty_weights(*a, **k):
    yield


from . import utils, hooks, big_modeling  # noqa: E402,F401
