"""Count the SASS mnemonics that prove which hardware paths each libgca kernel uses (B200_PROFILING.md: tcgen05.mma ->
UTCHMMA, tcgen05.commit -> UTCBAR, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG / UTMASTG, cp.async.bulk ->
UBLKCP, mbarrier -> SYNCS, mma.sync tf32 -> HMMA.1688.F32.TF32, mma.sync f16 (the scaled 2xFP16 split) -> HMMA.16816.F32,
movmatrix -> MOVM, multi-GPU barrier -> ST/LD .STRONG.SYS).   python profiles/sass_evidence.py > profiles/r2_sass_evidence.md"""
import collections
import re
import subprocess
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from gconv_adapter_b200.build import lib_path  # noqa: E402

lib = sys.argv[1] if len(sys.argv) > 1 else lib_path()
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA|UTCBAR|LDTM|UTMALDG|UTMASTG|UBLKCP|SYNCS|HMMA\.1688\.F32\.TF32|HMMA\.16816\.F32|MOVM|FFMA)\b")
keys = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA.1688.F32.TF32", "HMMA.16816.F32", "MOVM", "FFMA"]
cnt, cur = collections.OrderedDict(), None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        dem = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        dem = dem.replace("void ", "").replace("gca::(anonymous namespace)::", "").replace("gca::<unnamed>::", "").replace("gca::tc::", "tc::")
        dem = (dem.split(">(")[0] + ">") if ">(" in dem else dem.split("(")[0]
        dem = dem.replace("(bool)", "")
        cur = dem
        cnt.setdefault(cur, collections.Counter())
        continue
    if cur:
        m = pat.search(ln)
        if m:
            cnt[cur][m.group(1)] += 1
print("# SASS evidence (cuobjdump -sass of libgca.so): static instruction counts per kernel\n")
print("| kernel | " + " | ".join(keys) + " |")
print("|---|" + "---|" * len(keys))
for k in sorted(cnt):
    if k.startswith("k0_") or "k_" not in k:
        continue
    print("| `" + k + "` | " + " | ".join(str(cnt[k].get(x, 0)) for x in keys) + " |")
