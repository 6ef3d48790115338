"""Top SASS instructions by warp-stall samples for every kernel of an Nsight Compute report
(`ncu -i REPORT --page source --csv --print-source sass`), as committed under profiles/.

    python tools/ncu_hotspots.py gpurun_out/r1_cic32.ncu-rep [N] > profiles/..._hotspots.txt
"""
import csv
import io
import subprocess
import sys


def main(path, top=14):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    kernels, cur, hdr = [], None, None
    for r in rows:
        if len(r) >= 2 and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kernels.append(cur)
            hdr = None
        elif r and r[0] == "Address":
            hdr = r
        elif cur is not None and hdr is not None and len(r) == len(hdr):
            cur["rows"].append(dict(zip(hdr, r)))
    seen = set()
    for k in kernels:
        if k["name"] in seen:      # the source page lists a kernel once per view
            continue
        seen.add(k["name"])
        tot = sum(int(x["Warp Stall Sampling (All Samples)"] or 0) for x in k["rows"])
        print("=" * 110)
        print(k["name"][:108])
        print("instructions: %d, warp-stall samples: %d" % (len(k["rows"]), tot))
        stall_cols = [c for c in k["rows"][0] if c.startswith("stall_")] if k["rows"] else []
        agg = {c: sum(int(x[c] or 0) for x in k["rows"]) for c in stall_cols}
        print("samples by reason: " + ", ".join("%s %.0f%%" % (c[6:], 100.0 * v / max(tot, 1))
                                                for c, v in sorted(agg.items(), key=lambda t: -t[1])[:6]))
        print("%7s %6s  %-14s %s" % ("samples", "share", "top reason", "instruction"))
        for x in sorted(k["rows"], key=lambda x: -int(x["Warp Stall Sampling (All Samples)"] or 0))[:top]:
            s = int(x["Warp Stall Sampling (All Samples)"] or 0)
            why = max(stall_cols, key=lambda c: int(x[c] or 0))[6:] if stall_cols else ""
            print("%7d %5.1f%%  %-14s %s" % (s, 100.0 * s / max(tot, 1), why, x["Source"].strip()[:80]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 14)
