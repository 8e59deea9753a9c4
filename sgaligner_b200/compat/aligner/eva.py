"""Drop-in shim for ``from aligner.eva import EVA`` (the reference's EVA baseline, src/aligner/eva.py)."""
from sgaligner_b200.eva import *                   # noqa: F401,F403
from sgaligner_b200.eva import __all__             # noqa: F401
