from .cli import cli_main

cli_main()
