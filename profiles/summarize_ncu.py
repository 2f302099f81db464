"""Turn an `ncu --set full` report into the small CSV that is committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof_r1d.ncu-rep profiles/r1d_full_summary.csv

Runs here (no GPU needed): `ncu -i <rep> --page raw --csv` and keeps, per profiled launch, the metrics the roofline
statements in DESIGN.md / bench.py rest on.  `traffic_MB` = dram__bytes_read.sum + dram__bytes_write.sum (per launch).
"""
import csv
import io
import subprocess
import sys

KEEP = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("sm__cycles_elapsed.max", "sm_cycles"),
]

UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {name: i for i, name in enumerate(head)}
    kname = col["Kernel Name"]
    fields = ["id", "kernel"] + [short for _, short in KEEP] + ["traffic_MB", "time_us"]
    with open(out, "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(fields)
        for r in body:
            vals = {}
            for metric, short in KEEP:
                if metric in col and r[col[metric]] != "":
                    v = float(r[col[metric]].replace(",", ""))
                    v *= UNIT_SCALE.get(units[col[metric]], 1.0)
                    vals[short] = v
            traffic = (vals.get("dram_read", 0.0) + vals.get("dram_write", 0.0)) / 1e6
            name = r[kname].split("(")[0].replace("void ", "")
            wr.writerow([r[col["ID"]], name] + [("%.6g" % vals[s]) if s in vals else "" for _, s in KEEP]
                        + ["%.3f" % traffic, "%.2f" % vals.get("time", float("nan"))])
    print("wrote", out, len(body), "launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
