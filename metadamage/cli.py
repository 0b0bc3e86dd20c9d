from metadamage_b200.cli import *  # noqa: F401,F403
from metadamage_b200.cli import cli_app, cli_main  # noqa: F401
