"""SASS opcode counts per kernel of csrc/libdfine_sm100.so (cuobjdump -sass | c++filt) -> profiles/<name>.md.

    python tools/sass_summary.py profiles/r2_sass_opcodes.md
Runs without a GPU.  Mnemonics (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA load,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = legacy mma.sync, UCGABAR = cluster barrier, ACQBULK / PREEXIT =
griddepcontrol.wait / launch_dependents (programmatic dependent launch).
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
so = ROOT / "custom_d_fine_b200" / "csrc" / "libdfine_sm100.so"
out = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "profiles" / "sass_opcodes.md"
sass = subprocess.run(f"cuobjdump -sass {so} | c++filt", shell=True, capture_output=True, text=True).stdout
COLS = ["UTC*MMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "SYNCS", "HMMA", "UCGABAR", "ACQBULK", "PREEXIT", "FFMA"]
rows, cur, cnt = [], None, None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (.*)", line)
    if m:
        if cur:
            rows.append((cur, cnt))
        name = m.group(1)
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)
        cur, cnt = name.strip(), collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        cnt["total"] += 1
        base = op.split(".")[0]
        if re.match(r"UTC\w*MMA", base):
            cnt["UTC*MMA"] += 1
        elif base in COLS:
            cnt[base] += 1
if cur:
    rows.append((cur, cnt))
rows.sort(key=lambda r: (-r[1]["UTC*MMA"], -r[1]["HMMA"], -r[1]["total"]))
with open(out, "w") as f:
    f.write("# SASS opcode counts per kernel of libdfine_sm100.so (cuobjdump -sass, sm_100a; tools/sass_summary.py)\n\n")
    f.write("Tensor-core / TMA / TMEM mnemonics (B200_PROFILING.md): `UTC*MMA` = tcgen05.mma, `LDTM` / `STTM` = tcgen05.ld / st, "
            "`UTMALDG` = TMA load,\n`UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops, `HMMA` = legacy mma.sync (the attention "
            "kernels), `UCGABAR` = cluster barrier,\n`ACQBULK` / `PREEXIT` = griddepcontrol.wait / launch_dependents "
            "(programmatic dependent launch: every kernel has both).\n\n")
    f.write("| kernel | " + " | ".join(COLS) + " | total instr |\n|---|" + "---:|" * (len(COLS) + 1) + "\n")
    tot = collections.Counter()
    for name, c in rows:
        f.write(f"| `{name[:90]}` | " + " | ".join(str(c[k]) for k in COLS) + f" | {c['total']} |\n")
        tot.update(c)
    f.write(f"| **all {len(rows)} kernels** | " + " | ".join(str(tot[k]) for k in COLS) + f" | {tot['total']} |\n")
print(out, len(rows), "kernels")
