"""Per-kernel SASS evidence of the Blackwell-native paths (run here, no GPU needed):
    python profiles/sass_summary.py > profiles/sass_r2.md
Counts, for every kernel of libsgqn_b200.so, the SASS mnemonics that prove tcgen05 (UTC*MMA), tensor-memory loads / stores
(LDTM / STTM), TMA (UTMALDG / UTMASTG / UBLKCP) and -- as the legacy tensor path that should NOT appear -- HMMA."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "sgqn-carla_b200", "libsgqn_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
kern, rows = None, collections.OrderedDict()
pat = {"UTC*MMA": r"\bUTC[A-Z]*MMA\b", "LDTM": r"\bLDTM\b", "STTM": r"\bSTTM\b", "UTMALDG": r"\bUTMALDG\b", "UTMASTG": r"\bUTMASTG\b",
       "UBLKCP": r"\bUBLKCP\b", "SYNCS (mbarrier)": r"\bSYNCS\b", "HMMA (legacy)": r"\bHMMA\b", "instructions": r"^\s+/\*[0-9a-f]{4}\*/"}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern).split("(")[0]
        rows[kern] = collections.Counter()
        continue
    if kern:
        for k, p in pat.items():
            if re.search(p, line):
                rows[kern][k] += 1
print("# SASS evidence, round 2 build (`cuobjdump -sass sgqn-carla_b200/libsgqn_b200.so`, sm_100a)\n")
print("Kernels that issue tcgen05 MMAs (accumulators in TMEM, operands staged by TMA):\n")
cols = list(pat)
print("| kernel | " + " | ".join(cols) + " |")
print("|---|" + "---:|" * len(cols))
tc = [(k, c) for k, c in rows.items() if c["UTC*MMA"]]
for k, c in sorted(tc):
    print(f"| `{k}` | " + " | ".join(str(c[x]) for x in cols) + " |")
print(f"\n{len(tc)} tensor-core kernels; {len(rows) - len(tc)} CUDA-core kernels (gather, saliency select, losses, LayerNorm, Adam, RNG, "
      "SIMT GEMM for the 102-wide / 1-wide layers).  No kernel contains HMMA (mma.sync)." if not any(c["HMMA (legacy)"] for c in rows.values())
      else "\nWARNING: HMMA present")
print("\nAll kernels:\n")
for k, c in sorted(rows.items()):
    print(f"- `{k}`: {c['instructions']} instructions")
