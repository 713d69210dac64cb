"""Drop-in for the reference module ``pyHASPI/pyhaspi2.py`` (hot-path entry points only):
``from pyHASPI.pyhaspi2 import haspi_v2`` (intel.py:7) resolves here when
``nele_gan_b200.api.dropin_path()`` is first on ``sys.path``."""
from nele_gan_b200.api import haspi, haspi_v2, hasqi_v2  # noqa: F401
