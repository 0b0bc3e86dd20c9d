from metadamage_b200.main import *  # noqa: F401,F403
