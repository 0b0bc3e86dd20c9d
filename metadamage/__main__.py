from metadamage_b200.cli import cli_main

cli_main()
