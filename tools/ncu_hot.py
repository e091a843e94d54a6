"""Top stall-sample SASS lines of an `ncu --page source --csv` export (gzip ok).
usage: python tools/ncu_hot.py file.csv[.gz] [N]"""
import csv, gzip, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = list(csv.reader(f))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
col = {n: i for i, n in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) >= len(hdr) - 2]
def num(r, n):
    try: return float(r[col[n]])
    except Exception: return 0.0
tot = sum(num(r, "# Samples") for r in body)
print(f"{path}: {len(body)} SASS lines, {tot:.0f} samples")
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
idx = sorted(range(len(body)), key=lambda i: -num(body[i], "# Samples"))[:top]
for i in sorted(idx):
    r = body[i]
    s = num(r, "# Samples")
    why = sorted(((num(r, n), n[6:]) for n in stalls), reverse=True)[:3]
    print(f"{i:5d} {100*s/tot:5.1f}%  exec={r[col['Instructions Executed']]:>9s}  {r[col['Source']].strip()[:70]:70s} " +
          " ".join(f"{n}:{v:.0f}" for v, n in why if v > 0))
