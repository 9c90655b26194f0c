"""gfs3d -- host side of the B200-native hot path of GFS-3DSeg_GWs (C ABI in include/gfs3d.h)."""
from . import ops  # noqa: F401
from ._lib import LIB_PATH, lib  # noqa: F401
