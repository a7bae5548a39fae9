#!/usr/bin/env python3
"""Reads .ncu-rep captures (brought back in gpurun_out/) with `ncu -i ... --page raw --csv` and
writes the handful of metrics DESIGN.md / bench.py quote into a small tracked text file.

  python profiles/summarise.py gpurun_out/r01_scan_tc_b1024.ncu-rep [...] > profiles/r01_scan_summary.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.max.per_second",
    "launch__grid_size",
    "launch__block_size",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum",
]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        yield {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def short(name):
    return name.split(".", 2)[-1] if name.startswith(("TPC.", "SM_C.", "SM_A.", "SM_B.")) else name


def main():
    for rep in sys.argv[1:]:
        print(f"== {rep}")
        for r in rows_of(rep):
            kname = r.get("Kernel Name", ("?", ""))[0]
            print(f"-- kernel {kname[:90]}  grid {r.get('Grid Size', ('', ''))[0]} block {r.get('Block Size', ('', ''))[0]}")
            keyed = {short(k): v for k, v in r.items()}
            for w in WANT:
                if w in keyed:
                    v, u = keyed[w]
                    print(f"   {w:88s} {v:>18s} {u}")
        print()


if __name__ == "__main__":
    main()
