"""Summarise an .ncu-rep (read here, on the CPU box) into profiles/<name>.summary.json (+ .csv of the
raw page for the key metrics).  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name"""
import csv, io, json, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "sm__cycles_elapsed.avg", "gpc__cycles_elapsed.max",
]

def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        e = {"kernel": d.get("Kernel Name"), "metrics": {}}
        for k in KEYS:
            if k in d:
                try: v = float(d[k].replace(",", ""))
                except ValueError: v = d[k]
                e["metrics"][k] = {"value": v, "unit": units[hdr.index(k)]}
        m = e["metrics"]
        def val(k, scale=1.0):
            return m[k]["value"] * scale if k in m else None
        unit_scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        rd = val("dram__bytes_read.sum", unit_scale.get(m.get("dram__bytes_read.sum", {}).get("unit", "byte"), 1.0))
        wr = val("dram__bytes_write.sum", unit_scale.get(m.get("dram__bytes_write.sum", {}).get("unit", "byte"), 1.0))
        if rd is not None and wr is not None:
            e["dram_bytes_per_launch"] = rd + wr
            t = m["gpu__time_duration.sum"]
            tus = t["value"] * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(t["unit"], 1.0)
            e["duration_us"] = tus
            e["dram_gbs_under_ncu"] = (rd + wr) / tus / 1e3
        res.append(e)
    json.dump({"source": rep, "launches": res}, open(out + ".summary.json", "w"), indent=1)
    print(json.dumps(res, indent=1)[:3000])

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
