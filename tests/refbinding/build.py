"""Build the reference's UNMODIFIED CPython binding (bindings/python/ckzg_wrap.c, compiled in place from
/root/reference -- nothing copied) against OUR headers and OUR library:

    gcc -shared -I include ckzg_wrap.c -lckzg_b200

This is the drop-in claim made concrete: the binding source does not change, only its build line
(INTEGRATION.md).  Output: tests/refbinding/_build/ckzg.<abi>.so (git-ignored; travels to the GPU box).
No-op when /root/reference is absent (GPU box: the prebuilt module is used)."""
import os
import subprocess
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = "/root/reference/bindings/python/ckzg_wrap.c"
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "ckzg" + sysconfig.get_config_var("EXT_SUFFIX"))


def build():
    if not os.path.exists(SRC):
        return OUT if os.path.exists(OUT) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    lib_dir = os.path.join(ROOT, "c-kzg-4844_b200")
    inc = sysconfig.get_paths()["include"]
    cmd = ["gcc", "-shared", "-fPIC", "-O1", "-I", os.path.join(ROOT, "include"), "-I", inc, SRC, "-L", lib_dir, "-lckzg_b200",
           "-Wl,-rpath,$ORIGIN/../../../c-kzg-4844_b200", "-o", OUT]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build())
