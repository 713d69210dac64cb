"""Drop-in for ``pystoi``: ``from pystoi.stoi import stoi`` (intel.py:8), ``from pystoi import stoi``."""
from nele_gan_b200.api import stoi  # noqa: F401
