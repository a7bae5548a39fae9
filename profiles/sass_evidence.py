#!/usr/bin/env python3
"""Counts the Blackwell-specific SASS mnemonics per kernel of libicdrag.so (no GPU needed):
   python profiles/sass_evidence.py > profiles/sass_evidence.txt
tcgen05.mma = UTC*MMA, tcgen05.ld / st = LDTM / STTM, TMA = UTMALDG / UTMASTG, tcgen05.commit = UTCBAR, mbarrier = SYNCS;
HMMA is mma.sync: only in skinny_linear_kernel, the few-token (<= 64) latency path of the encoder, where a 128-lane
tcgen05 tile would be 90 % padding; every throughput kernel is tcgen05."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "rag-project-icd10_b200", "csrc", "libicdrag.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = {"UTC*MMA": r"\bUTC[A-Z]*MMA", "LDTM": r"\bLDTM", "STTM": r"\bSTTM", "UTMALDG": r"\bUTMALDG", "UTMASTG": r"\bUTMASTG",
       "UTCBAR": r"\bUTCBAR", "SYNCS": r"\bSYNCS", "HMMA": r"\bHMMA"}
cur, cnt = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
    elif cur:
        for k, p in pat.items():
            if re.search(p, line):
                cnt[cur][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
print("# SASS mnemonics per kernel of libicdrag.so (cuobjdump -sass, sm_100a); see the docstring of profiles/sass_evidence.py")
for (k, c), name in zip(cnt.items(), names):
    name = name.replace("icd::(anonymous namespace)::", "").replace("void ", "")
    name = re.sub(r"\((CUtensorMap_st|const|icd::|int|void|float|long|unsigned|__nv).*", "", name)
    print(f"{name[:64]:64s} " + ("  ".join(f"{a}={c[a]}" for a in pat if c[a]) or "(CUDA-core kernel)"))
print("# HMMA total:", sum(c["HMMA"] for c in cnt.values()))
