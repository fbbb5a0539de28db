"""Per-instruction view of an .ncu-rep (source page, SASS): total warp instructions executed, the opcode mix weighted
by execution count, and the hottest stall sites.  Usage: python tools/ncu_sass_hot.py rep.ncu-rep [top]"""
import csv, io, subprocess, sys, collections, json

def main(rep, top=25, out=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ix = {h: n for n, h in enumerate(hdr)}
    total = 0; mix = collections.Counter(); stalls = []
    samples_total = 0
    for r in rows[2:]:
        if len(r) < len(hdr): continue
        n = int(r[ix["Instructions Executed"]] or 0)
        s = int(r[ix["Warp Stall Sampling (All Samples)"]] or 0)
        src = r[ix["Source"]].strip()
        op = src.split()[1] if src.startswith("@") else src.split()[0]
        op = op.split(".")[0].rstrip(";")
        mix[op] += n; total += n; samples_total += s
        stalls.append((s, n, src))
    res = {"source": rep, "warp_instructions_executed": total, "opcode_mix": mix.most_common(40),
           "stall_samples_total": samples_total,
           "hottest": [{"samples": s, "executed": n, "sass": src} for s, n, src in sorted(stalls, reverse=True)[:top]]}
    if out:
        json.dump(res, open(out, "w"), indent=1)
    print("warp instructions executed:", total)
    print("mix:", ", ".join(f"{k} {v/total*100:.1f}%" for k, v in mix.most_common(24)))
    for s, n, src in sorted(stalls, reverse=True)[:top]:
        print(f"{s:7d} {s/max(samples_total,1)*100:5.1f}%  x{n:9d}  {src}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25, sys.argv[3] if len(sys.argv) > 3 else None)
