from sgaligner_b200.losses import *                # noqa: F401,F403
from sgaligner_b200.losses import __all__          # noqa: F401
