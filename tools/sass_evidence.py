#!/usr/bin/env python
"""profiles/rNN_sass_tcgen05.md: the Blackwell-specific SASS of the shipped library (no GPU needed).

    python tools/sass_evidence.py [out.md]
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bithtm_b200", "_lib", "libbithtm_b200.so")
PATS = ["UTCIMMA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "SYNCS", "IMMA", "FENCE.VIEW.ASYNC", "UTCATOMSWS"]


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_tcgen05.md")
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    out = ["# SASS evidence -- `cuobjdump -sass bithtm_b200/_lib/libbithtm_b200.so` (sm_100a), round 2", "",
           "Blackwell-specific instructions per kernel (tcgen05.mma = UTCIMMA, tcgen05.ld = LDTM, cp.async.bulk.tensor = "
           "UTMALDG, tcgen05.commit = UTCBAR, mbarrier = SYNCS, tcgen05.alloc = UTCATOMSWS; the legacy mma.sync path = IMMA):", ""]
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        cnt = {p: len(re.findall((r"\bIMMA\." if p == "IMMA" else r"\b" + re.escape(p) + r"\b"), f)) for p in PATS}
        n = len(re.findall(r"/\*[0-9a-f]{4}\*/", f))
        if any(cnt.values()):
            out.append(f"- `{name}` ({n} instructions): " + ", ".join(f"{k} {v}" for k, v in cnt.items() if v))
    out += ["", "## `t5::k_sp_overlap_batched_t5` (bh_sp_overlap_batched_tc5): the tensor-memory / TMA / MMA instructions", "```"]
    for f in funcs[1:]:
        if "k_sp_overlap_batched_t5" in f.split("\n", 1)[0]:
            for ln in f.split("\n"):
                if re.search(r"UTCIMMA|LDTM|UTMALDG|UTCBAR|FENCE.VIEW.ASYNC|UTCATOMSWS|SYNCS.ARRIVE|SYNCS.PHASECHK", ln):
                    out.append(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", ln.strip()))
    out.append("```")
    open(out_path, "w").write("\n".join(out) + "\n")
    print(out_path)


if __name__ == "__main__":
    main()
