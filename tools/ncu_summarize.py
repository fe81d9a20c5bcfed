"""Turns `ncu --set full` reports into the small text summaries committed under profiles/ (the .ncu-rep files themselves are
scratch: tens of MB).  Usage:  python tools/ncu_summarize.py OUT_PREFIX report1.ncu-rep [report2.ncu-rep ...]
Writes OUT_PREFIX_summary.txt (key metrics per captured launch) and OUT_PREFIX_traffic.json ({kernel: DRAM bytes per launch},
read by bench.py for `roofline.traffic`)."""
import csv
import io
import json
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    prefix, reps = sys.argv[1], sys.argv[2:]
    lines, traffic = [], {}
    for rep in reps:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            lines.append(f"# {rep}: no launches captured")
            continue
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for vals in rows[2:]:
            name = vals[col["Kernel Name"]]
            short = re.sub(r"\(.*", "", name).replace("void ", "").strip()
            lines.append(f"## {short}   [{rep.split('/')[-1]}]  grid {vals[col['Grid Size']] if 'Grid Size' in col else ''}")
            for k in KEYS:
                if k in col:
                    lines.append(f"{k} [{units[col[k]]}] = {vals[col[k]]}")
            try:
                rd = float(vals[col["dram__bytes_read.sum"]].replace(",", "")) * TO_BYTES[units[col["dram__bytes_read.sum"]]]
                wr = float(vals[col["dram__bytes_write.sum"]].replace(",", "")) * TO_BYTES[units[col["dram__bytes_write.sum"]]]
                key = re.sub(r"<.*", "", short).replace("vpho::", "")
                key = {"mano_forward_kernel": "mano_forward", "k_mano_tc": "mano_forward"}.get(key, key)
                traffic[key] = max(traffic.get(key, 0), int(rd + wr))
            except Exception:
                pass
            lines.append("")
    open(prefix + "_summary.txt", "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(prefix + "_traffic.json", "w"), indent=1)
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
