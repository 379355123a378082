#!/usr/bin/env python
"""Per-role stall breakdown of the pipelined cell kernel from an ncu source page (ncu -i rep --page source --csv):
the kernel's SASS is split at its USETMAXREG instructions (producer warps / scatter warps / DMMA warps), out-of-line
wait loops are attributed to the role whose code jumps to them.  Usage: python tools/ncu_roles.py src.csv [n_hot]"""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 12
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
ins = rows[2:]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = int(ins[0][col["Address"]], 16)
recs = []
for r in ins:
    a = int(r[col["Address"]], 16) - base
    recs.append(dict(a=a, s=r[col["Source"]].strip(), n=int(r[col["# Samples"]]), ex=int(r[col["Instructions Executed"]]),
                     st={k: int(r[col[k]]) for k in stalls if int(r[col[k]])}))
# role boundaries
marks = [(x["a"], x["s"]) for x in recs if "USETMAXREG" in x["s"]]
print("register re-splits:", [(hex(a), s) for a, s in marks])
bounds = [m[0] for m in marks]
names = ["prologue", "producers(A,gather,publisher)", "scatter", "dmma"]
def role_of(a):
    k = sum(1 for b in bounds if a >= b)
    return names[min(k, len(names) - 1)]
# out-of-line tails: after the last EXIT of the dmma section jumps go back; attribute by branch target
last = bounds[-1]
# find the end of the DMMA role's main body = the first 'EXIT' after the last mark
end_dmma = None
for x in recs:
    if x["a"] > last and x["s"].startswith("EXIT"):
        end_dmma = x["a"]; break
tail_owner = {}
if end_dmma:
    # every tail block ends with 'BRA 0x....' back into its owner
    blk = []
    for x in recs:
        if x["a"] <= end_dmma: continue
        blk.append(x)
        m = re.match(r"BRA (0x[0-9a-f]+)", x["s"])
        if m:
            tgt = int(m.group(1), 16) - base
            for y in blk: tail_owner[y["a"]] = role_of(tgt) if tgt <= end_dmma else None
            blk = []
tot = {}
for x in recs:
    r = tail_owner.get(x["a"]) or role_of(x["a"])
    x["role"] = r
    t = tot.setdefault(r, dict(n=0, st={}))
    t["n"] += x["n"]
    for k, v in x["st"].items(): t["st"][k] = t["st"].get(k, 0) + v
alln = sum(t["n"] for t in tot.values())
for r, t in tot.items():
    print(f"\n== {r}: {t['n']} samples ({100*t['n']/alln:.1f} %)")
    for k, v in sorted(t["st"].items(), key=lambda kv: -kv[1])[:8]:
        print(f"   {k:26s} {v:7d} {100*v/max(t['n'],1):5.1f} %")
    hot = sorted([x for x in recs if x["role"] == r], key=lambda x: -x["n"])[:nhot]
    for x in hot:
        print(f"   {x['n']:6d} {x['ex']:9d} {hex(x['a']):>7s}  {x['s'][:60]:60s} {x['st']}")
