#!/usr/bin/env python3
"""DRAM bytes per launch from `ncu --set full` captures of profiles/ncu_scan.py (one .ncu-rep per corpus size, two scan_tc
launches -- sampling pre-pass, main scan -- per batch size in BATCHES order) -> entries of profiles/ncu_traffic.json.
   python profiles/traffic_from_ncu.py r02h gpurun_out/r02h_scan_100000000.ncu-rep:100000000:1024,256,128 ..."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    for vals in rows[2:]:
        yield dict(zip(hdr, vals))


def num(x):
    return float(x.replace(",", ""))


def main():
    tag = sys.argv[1]
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    table = json.load(open(path))
    for spec in sys.argv[2:]:
        rep, nrows, batches = spec.split(":")
        launches = [r for r in rows_of(rep) if "scan_tc" in r.get("Kernel Name", "")]
        bl = [int(b) for b in batches.split(",")]
        assert len(launches) == 2 * len(bl), (rep, len(launches))
        for i, B in enumerate(bl):
            pre, main_ = launches[2 * i], launches[2 * i + 1]
            def bytes_of(r):
                rd = [k for k in r if k.endswith("dram__bytes_read.sum")][0]
                wr = [k for k in r if k.endswith("dram__bytes_write.sum")][0]
                return num(r[rd]), num(r[wr])
            # units: the raw page reports bytes scaled (Gbyte/Mbyte) per column; read the unit row through a second pass
            table[f"scan_tc_kernel:{nrows}:{B}"] = {"launch": main_.get("Kernel Name", "")[:60], "raw": {"main": bytes_of(main_), "pre_pass": bytes_of(pre)},
                                                   "capture": f"profiles/{tag}_ncu_summary.txt ({os.path.basename(rep)}, launch {2 * i + 1})"}
    json.dump(table, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
