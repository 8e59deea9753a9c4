"""Drop-in shim: put ``sgaligner_b200/compat`` in front of the reference's ``src`` on ``sys.path``
and the reference's own ``from aligner.sg_aligner import *`` / ``from aligner.losses import *``
(src/trainers/trainval_sgaligner.py:11-12, src/inference/sgaligner/*.py) resolve to the B200 path."""
