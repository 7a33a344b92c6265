"""c-kzg-4844_b200: a B200-native KZG (EIP-4844 / EIP-7594) engine behind the c-kzg-4844 API.

The product is `libckzg_b200.so` (csrc/*.cu: sm_100a kernels + C ABI, src/ckzg.c: the frozen C API).
This Python package is the host-side mirror of the reference's Python binding
(bindings/python/ckzg_wrap.c:12-803): same function names and argument order, over ctypes.

The directory name contains '-' and '.', so import it through `__graft_entry__.load_package()` or
`importlib` (module name `ckzg_b200`).  There is no CPU fallback: importing works anywhere, but any
call needs the built library and a CUDA device, and raises otherwise.
"""
from .ckzg_py import *  # noqa: F401,F403
from .ckzg_py import LIB_PATH, SETUP_TXT  # noqa: F401
