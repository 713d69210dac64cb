"""Drop-in for ``pysiib``: ``from pysiib import SIIB`` (intel.py:4)."""
from nele_gan_b200.api import SIIB  # noqa: F401
