from sgaligner_b200.sg_aligner import *            # noqa: F401,F403
from sgaligner_b200.sg_aligner import __all__      # noqa: F401
