#!/usr/bin/env python3
"""Compact SASS of one kernel of libhsmc_gpu.so: `sass_fn.py <substring of the mangled name> [out]`
(one instruction per line: address, text).  Used to check registers/spills/packed-fp32 use before spending GPU time."""
import re, subprocess, sys, os
lib = os.environ.get("HSMC_GPU_LIB", os.path.join(os.path.dirname(__file__), "..", "hsmc_b200", "csrc", "libhsmc_gpu.so"))
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
for p in re.split(r"\n\s*Function : ", txt)[1:]:
    name = p.split("\n", 1)[0]
    if sys.argv[1] in name:
        out = []
        for ln in p.split("\n"):
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                out.append(f"{m.group(1)} {m.group(2).strip()}")
        dst = sys.argv[2] if len(sys.argv) > 2 else "/dev/stdout"
        open(dst, "w").write(name + "\n" + "\n".join(out) + "\n")
        break
