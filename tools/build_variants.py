"""A/B libraries of the pipelined cell kernel: compiles cell_kernel.cu with other build-time knobs (HX_PIPE_*) and links it
with the objects of the shipped library into dft_efe_b200/lib/exp/libhxb200_<name>.so.  Select one with
HXB200_LIB=<path> (capi.py); the shipped libhxb200.so is never touched.  Usage:
    python tools/build_variants.py name1:HX_PIPE_REGD=112,HX_PIPE_REGS=56,HX_PIPE_RBF=2 name2:HX_PIPE_NACC=2 ...
"""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dft_efe_b200 import build as b  # noqa: E402


def build_variant(name: str, defines: list[str], source: str = "cell_kernel.cu", alt_path: str | None = None) -> str:
    """alt_path: another revision of `source` (e.g. `git show <rev>:dft_efe_b200/csrc/cell_kernel.cu > /tmp/x.cu`)"""
    b.build()
    exp = os.path.join(b.LIBDIR, "exp")
    os.makedirs(exp, exist_ok=True)
    obj = os.path.join(exp, f"{source[:-3]}_{name}.o")
    cmd = [b._nvcc()] + b.NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-I", b.CSRC, "-c", alt_path or os.path.join(b.CSRC, source), "-o", obj]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    open(obj + ".ptxas.log", "w").write(out.stdout)
    if out.returncode != 0:
        sys.stderr.write(out.stdout)
        raise RuntimeError(f"nvcc failed on variant {name}")
    objs = [os.path.join(b.LIBDIR, "obj", s.replace(".cu", ".o")) for s in b.SOURCES if s != source] + [obj]
    lib = os.path.join(exp, f"libhxb200_{name}.so")
    subprocess.check_call([b._nvcc(), "-shared", "-o", lib] + objs + ["-lcudart", "-ldl"])
    return lib


if __name__ == "__main__":
    procs = []
    for spec in sys.argv[1:]:
        name, _, defs = spec.partition(":")
        alt = None
        if "@" in defs:  # name:DEF=1,DEF2=2@/path/to/other_revision_of_cell_kernel.cu
            defs, _, alt = defs.partition("@")
        print(build_variant(name, [d for d in defs.split(",") if d], alt_path=alt))
