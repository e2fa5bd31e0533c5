"""Summarise an .ncu-rep (read here, no GPU needed) into a small text table for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [launches.csv] > profiles/rNN_ncu_summary.txt"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max", "gpc__cycles_elapsed.avg.per_second",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full --clock-control none summary of", rep)
    print("# (per-launch values; cold-cache, serialised replay: compare shares and ratios, not absolute times)")
    for r in rows[2:]:
        print("\n== {}  grid {} block {}".format(r[idx["Kernel Name"]].split("(")[0], r[idx.get("Grid Size", 0)], r[idx.get("Block Size", 0)]))
        for w in WANT:
            if w in idx:
                print("   {:92s} {:>16s} {}".format(w, r[idx[w]], units[idx[w]]))
    if len(sys.argv) > 2:
        print("\n# launch list (gpu__time_duration.sum, ncu --metrics pass) from", sys.argv[2])
        rows = list(csv.reader(open(sys.argv[2])))
        hdr = None
        agg = defaultdict(list)
        for r in rows:
            if r and r[0] == "ID":
                hdr = r
                continue
            if hdr and len(r) == len(hdr):
                d = dict(zip(hdr, r))
                agg[d["Kernel Name"].split("(")[0]].append(float(d["Metric Value"]) / 1e3)
        tot = sum(sum(v) for v in agg.values())
        for k, v in agg.items():
            print("   {:40s} launches {:3d}  avg {:9.1f} us  share of step {:5.1f} %".format(k, len(v), sum(v) / len(v), 100 * sum(v) / tot))


if __name__ == "__main__":
    main()
