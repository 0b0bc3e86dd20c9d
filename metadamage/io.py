from metadamage_b200.io import *  # noqa: F401,F403
