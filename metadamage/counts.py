from metadamage_b200.counts import *  # noqa: F401,F403
