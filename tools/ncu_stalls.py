"""aggregate the SASS-level source page of an ncu report (ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > f.csv):
warp-stall samples by reason, by opcode, and the hottest instructions."""
import collections
import csv
import sys


def analyze(path, ntop=25):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    idx = {name: i for i, name in enumerate(hdr)}
    stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    tot, byop, top, total = collections.Counter(), collections.Counter(), [], 0
    for r in rows[h + 1:]:
        if len(r) < len(hdr) or r[0] == "Address":
            continue
        try:
            s = int(r[idx["# Samples"]] or 0)
        except ValueError:
            continue
        total += s
        op = r[idx["Source"]].split()
        opn = (op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "?")).split(".")[0]
        byop[opn] += s
        for c in stall_cols:
            tot[c] += int(r[idx[c]] or 0)
        top.append((s, r[idx["Source"]].strip(), r[idx["Instructions Executed"]]))
    print("total samples", total)
    print("stalls:", [(k, v) for k, v in tot.most_common(10)])
    print("by opcode:", byop.most_common(14))
    top.sort(reverse=True)
    for s, src, n in top[:ntop]:
        print(f"   {s:7d} {n:>9s}  {src[:110]}")


if __name__ == "__main__":
    analyze(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
