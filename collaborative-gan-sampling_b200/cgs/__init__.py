"""cgs: ctypes loader for libcgs.so, network specs / weight packing, and the multi-GPU driver."""
from . import lib  # noqa: F401
