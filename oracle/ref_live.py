"""Live importer of the UNMODIFIED reference (test infrastructure — never a product path).

Only usable where ``/root/reference`` is mounted (the build container).  It is used by
``tests/golden/make_golden.py`` to mint the golden vectors and by
``tests/test_oracle_vs_reference.py`` (skipped when the tree is absent) to pin the restatement in
``oracle/ctvq_oracle.py`` against the real thing.  Nothing under ``ct_vae_b200/`` may import this.

The reference's ``models/__init__.py:25`` star-imports ``ct_mcq_vae`` which needs ``torch_geometric``
(``models/ct_mcq_vae.py:2,5``; not installed, no network).  Three empty stub modules in
``sys.modules`` make the package import; the quantiser classes never touch them.
"""
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("CTVQ_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "vq_vae.py"))


def load():
    """Return the reference's ``models`` package (imported from REFERENCE_ROOT, read-only)."""
    if not available():
        raise RuntimeError(f"reference tree not mounted at {REFERENCE_ROOT}")
    for name in ("torch_geometric", "torch_geometric.nn", "torch_geometric.utils"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # SyntaxWarnings from the reference's docstrings
        import models  # noqa: E402  (the reference's package)
    return models
