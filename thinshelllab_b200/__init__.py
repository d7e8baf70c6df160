"""thinshelllab_b200 -- B200 (sm_100a) implementation of ThinShellLab's differentiable implicit thin-shell step.

Host Python keeps state in torch CUDA tensors and calls hand-written CUDA through the C ABI of
libtsl.so (include/tsl.h).  There is no CPU fallback: importing works anywhere (so the ABI can be checked),
but creating an engine without a GPU or without the built library raises.
"""
from ._lib import lib, LibraryMissing, TslError  # noqa: F401
from .core import ShellEngine, StepStats  # noqa: F401

__version__ = "0.1.0"
