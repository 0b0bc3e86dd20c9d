from metadamage_b200.fits import *  # noqa: F401,F403
