#!/usr/bin/env python
"""Fingerprint of the SASS of every kernel in maelstrom_b200/libmaelstrom_b200.so: sha256 over the
instruction stream with encodings dropped and constant-bank offsets of kernel parameters beyond the
first (`c[0x0][0x6xx..0xaxx]`, which move whenever Params grows) masked.  Used to tell whether a
change touched the kernels whose numbers are in profiles/ (the WL = 0 instantiations of k_round).

    python tools/sass_fingerprint.py            # print
    python tools/sass_fingerprint.py --write    # update profiles/sass_fingerprint.json
"""
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "maelstrom_b200", "libmaelstrom_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass_fingerprint.json")


def fingerprint(so=SO):
    txt = subprocess.run(["cuobjdump", "-sass", so], check=True, capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            t = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).strip()
            t = re.sub(r"c\[0x0\]\[0x[6-9a-f][0-9a-f]{2}\]", "c[0x0][PARAM]", t)
            funcs[cur].append(t)
    demangled = {}
    for name, ins in funcs.items():
        pretty = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
        pretty = re.sub(r"\(.*\)$", "", pretty).replace("void ", "")
        demangled[pretty] = {"instructions": len(ins), "sha256": hashlib.sha256("\n".join(ins).encode()).hexdigest()[:16]}
    return demangled


if __name__ == "__main__":
    fp = fingerprint()
    if "--write" in sys.argv:
        with open(OUT, "w") as f:
            json.dump(fp, f, indent=1, sort_keys=True)
            f.write("\n")
    for k in sorted(fp):
        print("%-44s %6d  %s" % (k, fp[k]["instructions"], fp[k]["sha256"]))
