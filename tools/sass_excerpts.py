"""Writes profiles/r02_sass_excerpts.txt: for every kernel of libmmf_b200.so, the count of the tensor / TMA / TMEM SASS
mnemonics that prove which hardware path it uses, plus a short excerpt around the first tensor instruction."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "multimodalfilter_b200", "libmmf_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "HMMA", "LDSM", "SYNCS", "FFMA2", "FHFMA", "LDGSTS", "REDUX", "SHFL"]
funcs, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        funcs[name] = []
    elif name and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
        funcs[name].append(line.split("/*", 2)[1].split("*/", 1)[1].rsplit("/*", 1)[0].rstrip(" ;") if False else line)
demangle = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
with open(os.path.join(ROOT, "profiles", "r02_sass_excerpts.txt"), "w") as f:
    f.write("SASS evidence from multimodalfilter_b200/libmmf_b200.so (cuobjdump -sass; sm_100a only).\n"
            "Counts of the mnemonics that identify the hardware path; UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,\n"
            "UBLKCP = cp.async.bulk (TMA bulk copy), UTCBAR = tcgen05.commit, HMMA = mma.sync, LDSM = ldmatrix, SYNCS = mbarrier.\n\n")
    for (mangled, lines), pretty in zip(funcs.items(), demangle):
        cnt = collections.Counter()
        for l in lines:
            op = re.search(r"\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", l)
            if op:
                for k in KEYS:
                    if op.group(1).startswith(k):
                        cnt[k] += 1
        f.write(f"{pretty[:150]}\n    {len(lines)} instructions; " + ", ".join(f"{k} {v}" for k, v in cnt.items()) + "\n")
        for key in ("UTCHMMA", "HMMA"):
            idx = [i for i, l in enumerate(lines) if re.search(r"\s" + key, l)]
            if idx:
                lo, hi = max(0, idx[0] - 3), min(len(lines), idx[0] + 9)
                f.write("".join("        " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip() + "\n" for l in lines[lo:hi]))
                break
        f.write("\n")
print("written", len(funcs), "kernels")
