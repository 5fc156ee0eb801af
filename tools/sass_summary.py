"""Per-kernel counts of the SASS mnemonics that prove which hardware paths the library uses (tcgen05 MMA, tensor-memory
loads, TMA, bulk copies, programmatic dependent launch), from `cuobjdump -sass` of the in-tree .so.  No GPU needed.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "egohmr_b200", "lib", "libegohmr_b200.so")
PATTERNS = [("UTCHMMA", "tcgen05.mma kind::f16"), ("UTCHMMA.2CTA", "... cta_group::2"), ("LDTM", "tcgen05.ld (TMEM -> registers)"),
            ("UTMALDG", "cp.async.bulk.tensor (TMA load)"), ("UTMALDG.4D", "... 4-D boxes (implicit-GEMM convolutions)"),
            ("UBLKCP", "cp.async.bulk (1-D bulk copy)"), ("UTCBAR", "tcgen05.commit -> mbarrier"), ("SYNCS", "mbarrier ops"),
            ("ACQBULK", "griddepcontrol.wait (PDL)"), ("PREEXIT", "griddepcontrol.launch_dependents (PDL)"),
            ("HMMA", "mma.sync (legacy tensor path; expected 0)"), ("FFMA", "fp32 FMA")]

out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
kern = None
counts = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("void ", "")
        name = re.sub(r"\(.*", "", name)
        kern = name
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[T\d]+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for pat, _ in PATTERNS:
            if op == pat or op.startswith(pat + ".") or (pat.count(".") and op.startswith(pat)):
                counts[kern][pat] += 1

print(f"# {os.path.relpath(SO, ROOT)}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)")
for pat, what in PATTERNS:
    print(f"#   {pat:14s} {what}")
cols = [p for p, _ in PATTERNS]
print(f"{'kernel':58s} " + " ".join(f"{c[:9]:>9s}" for c in cols))
tot = collections.Counter()
for k, c in counts.items():
    if not any(c[p] for p in cols if p != "FFMA"):
        if c["FFMA"] == 0:
            continue
    print(f"{k[:58]:58s} " + " ".join(f"{c[p]:9d}" for p in cols))
    tot.update(c)
print(f"{'TOTAL':58s} " + " ".join(f"{tot[p]:9d}" for p in cols))
