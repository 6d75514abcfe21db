#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into the small JSON summaries kept under profiles/:
   python tools/ncu_summarize.py gpurun_out/r2_full_k_cross.ncu-rep profiles/r2_ncu_full_k_cross_n16385.json"""
import csv
import io
import json
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "launch__shared_mem_per_block_dynamic"]


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    recs = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[hdr.index("Kernel Name")]
        name = re.sub(r"^void |pmg::<unnamed>::|\(.*$", "", name)
        name = re.sub(r"^.*::(?=k_)", "", name)  # whatever is left of an anonymous-namespace prefix
        rec = {"kernel": name}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                rec[k] = ("%s %s" % (r[i], units[i])).strip()
        recs.append(rec)
    json.dump(recs, open(out, "w"), indent=1)
    for rec in recs:
        print(rec["kernel"][:70], rec.get("gpu__time_duration.sum"), rec.get("dram__bytes_read.sum"), rec.get("dram__bytes_write.sum"),
              rec.get("launch__registers_per_thread"))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
