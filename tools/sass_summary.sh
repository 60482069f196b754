#!/usr/bin/env bash
# Per-kernel counts of the Blackwell-only SASS instructions in the shipped library (no GPU needed):
#   UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / .st (TMEM), UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit.
#   tools/sass_summary.sh > profiles/r2_sass_summary.txt
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
SO="$HERE/csm-train-pytorch_b200/libcsm_b200.so"
echo "# cuobjdump -sass $(basename "$SO")  (nvcc $(nvcc --version | grep release | sed 's/.*release //'), -gencode arch=compute_100a,code=sm_100a)"
echo "# kernel | UTCHMMA | LDTM | STTM | UTMALDG | UTCBAR | HMMA(mma.sync) | MUFU.EX2"
cuobjdump -sass "$SO" | awk '
  /Function :/ { if (name != "") print name, u, l, s, t, c, h, m; name=$3; u=l=s=t=c=h=m=0 }
  /UTCHMMA/ {u++} /LDTM/ {l++} /STTM/ {s++} /UTMALDG/ {t++} /UTCBAR/ {c++} / HMMA/ {h++} /MUFU.EX2/ {m++}
  END { print name, u, l, s, t, c, h, m }' | python3 -c '
import subprocess, sys
rows = [l.split() for l in sys.stdin if l.strip()]
names = subprocess.run(["c++filt", "-p"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
out = []
for n, r in zip(names, rows):
    if int(r[1]) or int(r[4]) or int(r[6]):
        n = n.replace("(anonymous namespace)::", "").replace("void ", "").replace("csm::", "")
        out.append(" | ".join([n[:110]] + r[1:]))
print("\n".join(sorted(out)))'
