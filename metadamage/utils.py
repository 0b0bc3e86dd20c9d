from metadamage_b200.utils import *  # noqa: F401,F403
from metadamage_b200.utils import Config, SubstitutionBases, extract_name  # noqa: F401
