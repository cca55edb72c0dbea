"""B200-native packet-erasure codec (LDPC + RS) behind the libldpc_cuda C ABI.

The package holds the CUDA library sources (csrc/), its build recipe (build.py), the committed
code definitions (codes/*.mat) and the host-side mirror of the reference's host program (codec.py).
Importing the package does not need a GPU; creating a codec does."""
from . import _lib  # noqa: F401
from .build import build  # noqa: F401

__all__ = ["LdpcCodec", "RsCodec", "fill_random", "pack_mask", "unpack_mask", "build"]


def __getattr__(name):
    if name in ("LdpcCodec", "fill_random", "pack_mask", "unpack_mask", "RsCodec"):
        from . import codec
        return getattr(codec, name)
    raise AttributeError(name)
