"""SASS opcode summary per kernel of libpggan_b200.so (cuobjdump -sass): proof that the tensor-core kernels are
tcgen05 (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA loads, SYNCS = mbarrier) and contain no legacy HMMA.
    python tools/sass_summary.py > profiles/r2_sass_opcodes.md"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "pggan_b200/lib/libpggan_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.splitlines()
blocks = re.split(r"\s+Function : \S+\n", txt)[1:]
ops_of_interest = ["UTCHMMA", "UTCQMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", "HMMA", "IMMA", "F2FP.SATFINITE", "ELECT", "RED", "ATOMG", "LDG", "STG", "LDS", "STS", "SHFL", "ACQBULK", "PREEXIT"]
rows = collections.OrderedDict()
for name, body in zip(names, blocks):
    short = re.sub(r"^void ", "", name)
    short = re.sub(r"\(.*", "", short).replace("pg::", "")
    base = re.sub(r"<.*", "", short)
    cnt = collections.Counter()
    n = 0
    for line in body.splitlines():
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        n += 1
        op = m.group(1)
        for o in ops_of_interest:
            if op.startswith(o):
                cnt[o] += 1
    r = rows.setdefault(base, {"variants": 0, "instr": 0, "cnt": collections.Counter()})
    r["variants"] += 1
    r["instr"] += n
    r["cnt"].update(cnt)
print("# r2 — SASS opcode counts per kernel (all template instantiations summed), `cuobjdump -sass libpggan_b200.so`\n")
print("UTCHMMA = tcgen05.mma (kind::f16 / tf32), LDTM = tcgen05.ld (TMEM -> registers), UTMALDG = cp.async.bulk.tensor (TMA) loads,")
print("SYNCS = mbarrier ops, ELECT = elect.sync, F2FP.SATFINITE = saturating fp32 -> half conversion, ACQBULK / PREEXIT = griddepcontrol.wait / .launch_dependents (programmatic dependent launch).  HMMA/IMMA (legacy mma.sync) must be 0.\n")
cols = ["UTCHMMA", "UTMALDG", "LDTM", "SYNCS", "ELECT", "HMMA", "IMMA", "F2FP.SATFINITE", "RED", "ATOMG", "SHFL", "ACQBULK", "PREEXIT"]
print("| kernel | variants | SASS instr | " + " | ".join(cols) + " |")
print("|---|---|---|" + "---|" * len(cols))
tot = collections.Counter()
for k, r in sorted(rows.items(), key=lambda kv: -kv[1]["cnt"]["UTCHMMA"]):
    print(f"| `{k}` | {r['variants']} | {r['instr']} | " + " | ".join(str(r["cnt"][c]) for c in cols) + " |")
    tot.update(r["cnt"])
print("| **total** | | | " + " | ".join(str(tot[c]) for c in cols) + " |")
