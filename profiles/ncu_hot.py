"""Per-instruction stall summary of an `ncu --page source --csv` export: hottest SASS lines with their
dominant stall reasons.  usage: ncu_hot.py src.csv [top] [lo_line hi_line]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[idx["# Samples"]]) for r in data)
    print("total samples", tot)
    agg = {h: sum(int(r[idx[h]] or 0) for r in data) for h in stall}
    print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    if len(sys.argv) > 4:
        lo, hi = int(sys.argv[3]), int(sys.argv[4])
        for k, r in enumerate(data[lo:hi]):
            s = int(r[idx["# Samples"]])
            reasons = sorted(((int(r[idx[h]] or 0), h[6:]) for h in stall), reverse=True)[:2]
            print(f"{lo + k:5d} {s:6d} {r[idx['Instructions Executed']]:>7s}  {r[1].strip()[:70]:70s} {reasons}")
        return
    order = sorted(range(len(data)), key=lambda k: -int(data[k][idx["# Samples"]]))[:top]
    for k in order:
        r = data[k]
        reasons = sorted(((int(r[idx[h]] or 0), h[6:]) for h in stall), reverse=True)[:3]
        print(f"{k:5d} {r[idx['# Samples']]:>6s} {r[idx['Instructions Executed']]:>7s}  {r[1].strip()[:70]:70s} {reasons}")


if __name__ == "__main__":
    main()
