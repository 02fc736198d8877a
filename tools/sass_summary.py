"""cuobjdump -sass of the built libpda_b200.so -> per-kernel counts of the Blackwell-native mnemonics
(UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA load/store, UTCBAR = tcgen05.commit, SYNCS =
mbarrier) plus a short excerpt around the first tensor-core instruction of each kernel.

    python tools/sass_summary.py > profiles/r02_sass_summary.md
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "probabilistic_domain_adaptation_b200", "libpda_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "FFMA2", "HFMA2", "FFMA",
             "REDG", "RED.", "ATOMG", "LDG", "STG", "LDS", "STS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731
    kernels, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            kernels[cur].append(line.rstrip())
    print("# SASS summary of the shipped `libpda_b200.so` (`cuobjdump -sass`, sm_100a)\n")
    print("Counts of instructions per kernel.  `UTCHMMA` = `tcgen05.mma`, `LDTM`/`STTM` = `tcgen05.ld`/`st`, `UTMALDG`/`UTMASTG`"
          " = TMA load / store, `UTCBAR` = `tcgen05.commit`, `SYNCS` = mbarrier ops; `HMMA` (legacy `mma.sync`) must be 0.\n")
    cols = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "FFMA", "FFMA2", "HFMA2", "LDG", "STG",
            "LDS", "STS", "REDG+ATOMG"]
    print("| kernel | instructions | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    tc = []
    for name, lines in kernels.items():
        text = "\n".join(lines)
        def cnt(mn):
            return len(re.findall(r"\b" + re.escape(mn) + r"[\.\s;]", text))
        row = {c: cnt(c) for c in cols if "+" not in c}
        row["FFMA"] -= row["FFMA2"]
        row["REDG+ATOMG"] = cnt("REDG") + cnt("ATOMG") + cnt("RED")
        short = demangle(name)
        short = re.sub(r"\(.*", "", short).replace("pda::", "").replace("void ", "")
        print(f"| `{short}` | {len(lines)} | " + " | ".join(str(row[c]) for c in cols) + " |")
        if row["UTCHMMA"]:
            tc.append((short, lines))
    print("\n## Excerpts (first `UTCHMMA` of each tensor-core kernel with its neighbours)\n")
    for short, lines in tc:
        i = next(k for k, l in enumerate(lines) if "UTCHMMA" in l)
        print(f"### `{short}`\n\n```")
        for l in lines[max(0, i - 4):i + 5]:
            print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip())
        print("```\n")


if __name__ == "__main__":
    main()
