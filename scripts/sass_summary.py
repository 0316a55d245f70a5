"""profiles/sass_r2.txt: per-kernel counts of the Blackwell tensor / TMA instructions in the built library
(cuobjdump -sass ctunet_b200/libctunet_b200.so): UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA tensor load),
UBLKCP (bulk copy), UTCBAR (tcgen05.commit), SYNCS (mbarrier), plus the cubin architectures present."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "ctunet_b200", "libctunet_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
MN = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "ATOMG", "RED"]
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        counts[cur]["_total"] += 1
        if op in MN:
            counts[cur][op] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
out = ["# SASS summary of ctunet_b200/libctunet_b200.so (cuobjdump -sass); architectures: %s" % ", ".join(arch),
       "# kernels with tensor-core / TMA instructions first; %d kernels in total" % len(counts), "",
       "%-110s %7s " % ("kernel", "instrs") + " ".join("%8s" % m for m in MN)]
rows = []
for (name, c), dm in zip(counts.items(), demangle):
    dm = re.sub(r"\(.*", "", dm.replace("void ", "").replace("ctu::", ""))
    rows.append((-(c["UTCHMMA"] + c["UTMALDG"] + c["LDTM"]), dm, c))
for _, dm, c in sorted(rows, key=lambda r: (r[0], r[1])):
    out.append("%-110s %7d " % (dm[:110], c["_total"]) + " ".join("%8d" % c[m] for m in MN))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
out.append("")
out.append("%-110s %7d " % ("TOTAL", tot["_total"]) + " ".join("%8d" % tot[m] for m in MN))
open(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_r2.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:14]))
print(out[-1])
