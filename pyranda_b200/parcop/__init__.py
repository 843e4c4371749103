"""Package layout mirror of the reference: `from pyranda import parcop; parcop.parcop.ddx(val)`
(pyranda/parcop/__init__.py:1 imports the f2py module `parcop` into the package `parcop`)."""
from . import parcop  # noqa: F401
