"""Import-compatibility shim: `metadamage.cli.cli_app`, `metadamage.utils.extract_name`, ...
resolve to the B200 build (metadamage_b200), so code and tests written against the reference's
package name keep working."""
from metadamage_b200 import __version__  # noqa: F401
