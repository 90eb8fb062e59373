from mantis_b200.models.mllava_next import *  # noqa: F401,F403
from mantis_b200.models.mllava_next import __all__  # noqa: F401
