"""Top SASS lines by warp-stall samples from `ncu -i X.ncu-rep --page source --csv` (stdin), grouped per kernel.
    ncu -i gpurun_out/x.ncu-rep --page source --csv | python scripts/ncu_src_top.py [N]
"""
import csv, sys
n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
i = 0
kern = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        kern += 1
        name = rows[i][1][:90]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr):
                body.append(rows[j])
            j += 1
        si = hdr.index("# Samples")
        src = hdr.index("Source")
        stall_cols = [k for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[si] or 0) for r in body)
        print(f"== kernel {kern}: {name}  total samples {tot}")
        agg = {}
        for k in stall_cols:
            agg[hdr[k]] = sum(int(r[k] or 0) for r in body)
        print("   stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
        for r in sorted(body, key=lambda r: -int(r[si] or 0))[:n]:
            st = {hdr[k][6:]: int(r[k] or 0) for k in stall_cols if int(r[k] or 0) > 0}
            top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
            print(f"   {int(r[si]):6d} {100.0 * int(r[si]) / max(tot, 1):5.1f}%  {r[src][:70]:70s} {top}")
        i = j
    else:
        i += 1
