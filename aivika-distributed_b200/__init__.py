"""twb200: B200-native Time Warp engine behind the aivika-distributed DIO surface (host-side mirror)."""
from . import _ffi  # noqa: F401
