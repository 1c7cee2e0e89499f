"""Summarise an ncu report (--set full) into the per-kernel CSV kept under profiles/ and the DRAM-traffic file bench.py reads.
usage: ncu_summary.py report.ncu-rep out.csv [traffic.json frames samples]
  out.csv: one row per captured launch: time, DRAM bytes read / written, executed warp instructions, issue-slot and warp
           activity, shared-memory pipe, L2 hit rate, occupancy limits
  traffic.json (optional): dram__bytes_read.sum + dram__bytes_write.sum per launch, normalised per frame (per sample for
           k_front) with the frame / sample count of the captured push; phases A and B of k_link_walk are added up"""
import csv
import io
import json
import subprocess
import sys

COLS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units, body = rows[0], rows[1], rows[2:]
    idx = {n: i for i, n in enumerate(head)}
    name_i = idx["Kernel Name"]
    have = [c for c in COLS if c in idx]
    table = []
    for r in body:
        rec = {"kernel": r[name_i]}
        for c in have:
            v = r[idx[c]].replace(",", "")
            try:
                f = float(v)
            except ValueError:
                f = float("nan")
            u = units[idx[c]]
            if c.startswith("dram__bytes"):
                f *= UNIT_SCALE.get(u, 1.0)            # bytes
            elif c == "gpu__time_duration.sum":
                f *= UNIT_SCALE.get(u, 1.0)            # microseconds
            rec[c] = f
        table.append(rec)
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["ID", "Kernel Name"] + have)
        w.writerow(["", ""] + ["byte" if c.startswith("dram__bytes") else "us" if c == "gpu__time_duration.sum" else units[idx[c]] for c in have])
        for i, rec in enumerate(table):
            w.writerow([i, rec["kernel"]] + ["%.6f" % rec[c] for c in have])
    if len(sys.argv) >= 6:
        frames, samples = float(sys.argv[4]), float(sys.argv[5])
        per = {}
        for rec in table:
            k = rec["kernel"].split("(")[0].split("<")[0].replace("void ", "").strip()
            b = rec.get("dram__bytes_read.sum", 0.0) + rec.get("dram__bytes_write.sum", 0.0)
            per.setdefault(k, []).append(b)
        traffic = {}
        for k, v in per.items():
            # launches of one push: the two phases of the walker add up, a kernel captured in several pushes is averaged per push
            pushes = max(1, len(per.get("k_front", [1])))
            tot = sum(v) / pushes
            traffic[k] = {"dram_bytes_per_sample": tot / samples} if k == "k_front" else {"dram_bytes_per_frame": tot / frames}
        json.dump({"_note": "ncu --set full --clock-control none over bench.py --workload c2 --frames-per-step %d (%d samples per push); "
                            "dram__bytes_read.sum + dram__bytes_write.sum per push, normalised per frame / per sample; k_link_walk: phase A + phase B"
                            % (frames, samples), "c2": traffic, "c5": traffic}, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
