"""Builds an A/B variant of libpda_b200.so with extra -D macros (experiments only: `PDA_B200_LIB=<path>` selects it).

    python tools/build_variant.py <tag> -DFC_W3_REG=64 [...]   ->  probabilistic_domain_adaptation_b200/libpda_b200_<tag>.so
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from probabilistic_domain_adaptation_b200 import build as b  # noqa: E402


def main():
    tag, macros = sys.argv[1], sys.argv[2:]
    out = os.path.join(b.PKG, f"libpda_b200_{tag}.so")
    objs, procs = [], []
    for src in b.SOURCES:
        obj = os.path.join(b.PKG, "build", f"{tag}_" + src.replace(".cu", ".o"))
        os.makedirs(os.path.dirname(obj), exist_ok=True)
        flags = [f for f in b.NVCC_FLAGS if f not in ("-Xptxas", "-v")]
        procs.append(subprocess.Popen(["nvcc", *flags, *macros, "-c", os.path.join(b.CSRC, src), "-o", obj]))
        objs.append(obj)
    for p in procs:
        if p.wait() != 0:
            raise SystemExit("nvcc failed")
    subprocess.check_call(["nvcc", "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    print(out)


if __name__ == "__main__":
    main()
