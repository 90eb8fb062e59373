from mantis_b200.models.idefics3 import *  # noqa: F401,F403
from mantis_b200.models.idefics3 import __all__  # noqa: F401
