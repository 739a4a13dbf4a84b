"""Import the UNMODIFIED reference (intel/auto-round v0.15.0, /root/reference) in this container.

TEST INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py to produce the fixtures under
tests/golden/ that pin the oracle.  /root/reference does not exist on the GPU box, so nothing
in the `-m gpu` tests, smoke() or bench.py may call this module.

Recipe: `accelerate` is a hard import of the reference but is not installed;
transformers' dependency check must run BEFORE a metadata-less `accelerate` becomes importable.
"""
import os
import sys

REFERENCE_ROOT = "/root/reference"
_STUB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "accelerate_stub")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "auto_round"))


def import_reference():
    if not reference_available():
        raise RuntimeError("reference tree not present (expected only in the build container)")
    import transformers  # noqa: F401
    import transformers.quantizers  # noqa: F401
    import transformers.modeling_utils  # noqa: F401

    if _STUB not in sys.path:
        sys.path.insert(0, _STUB)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import auto_round  # noqa: E402

    return auto_round


class DummyTokenizer:
    """The reference insists on a tokenizer object when `model` is an nn.Module (context/model.py:270)."""

    pad_token_id = None
    eos_token_id = None
    bos_token_id = None

    def save_pretrained(self, *a, **k):
        return None

    def __call__(self, *a, **k):
        raise RuntimeError("dummy tokenizer: datasets must be passed as token tensors")
