"""Shared fixtures for the -m gpu tests: the product library driven through the same ctypes front
end as the reference build (oracle.ref_lib.CKZG works on any library exporting the c-kzg API)."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402
from oracle import ref_lib  # noqa: E402

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def product_lib_path():
    return entry.load_package().LIB_PATH


def product(precompute=0):
    """CKZG front end on libckzg_b200.so -- raises if the library is missing (no fallback)."""
    return ref_lib.CKZG(so_path=product_lib_path(), precompute=precompute)


def reference(precompute=0):
    return ref_lib.CKZG(precompute=precompute)


def synth_blob(b, seed=4844):
    """SURVEY.md §8(d) synthetic input: element (b,i) = SHA256(le64(seed)||le64(b)||le64(i)) mod r."""
    pre = seed.to_bytes(8, "little") + b.to_bytes(8, "little")
    return b"".join((int.from_bytes(hashlib.sha256(pre + i.to_bytes(8, "little")).digest(), "big") % R).to_bytes(32, "big") for i in range(4096))
