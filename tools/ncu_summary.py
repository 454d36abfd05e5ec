"""Compact views of ncu output for profiles/ (the .ncu-rep files themselves stay in gpurun_out/).

  python tools/ncu_summary.py launches <ncu --csv log> <out.csv>     one row per launch: short kernel name,
                                                                    grid, block, us, DRAM read / write bytes
  python tools/ncu_summary.py report <file.ncu-rep> <out.csv>        key metrics of the first kernel in a
                                                                    --set full capture
"""
import collections
import csv
import re
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def short(name):
    m = re.search(r"(\w+_kernel)(<[^>]*>)?", name)
    if m and ("idiff" in name or m.group(1).split("_kernel")[0] in (
            "gemm2", "gemm", "attention2", "attention", "gn_stats", "gn_apply", "gn_fused", "row_stats", "layernorm40", "layernorm_generic",
            "scaleu_coef", "scaleu_reduce", "scaleu_apply", "upsample2x", "im2col_s2", "silu_f16", "fourier_embed",
            "timestep_embedding", "plms_update", "latent_mean", "nchw_f32_to_nhwc_f16", "nhwc_f16_to_nchw_f32")):
        return m.group(1) + (m.group(2) or "")
    m = re.search(r"at::native::(\w+)", name)
    return "torch:" + (m.group(1) if m else name[:40])


def launches(src, dst):
    with open(src) as f:
        rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
    per = collections.OrderedDict()
    for r in rows:
        d = per.setdefault(r["ID"], {"kernel": short(r["Kernel Name"]), "grid": r["Grid Size"], "block": r["Block Size"]})
        v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
        d[r["Metric Name"]] = v
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "time_us", "dram_read_bytes", "dram_write_bytes"])
        for i, d in per.items():
            w.writerow([i, d["kernel"], d["grid"], d["block"], round(d.get("gpu__time_duration.sum", 0), 2),
                        int(d.get("dram__bytes_read.sum", 0)), int(d.get("dram__bytes_write.sum", 0))])
    tot = sum(d.get("gpu__time_duration.sum", 0) for d in per.values())
    print(f"{len(per)} launches, {tot / 1e3:.2f} ms (cold caches, serialised)")


KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
    "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
    "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_selected",
    "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
    "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_sample_count",
]


def report(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for k in ("Kernel Name", "Grid Size", "Block Size"):
            w.writerow([k, "", vals[col[k]]])
        for k in KEYS:
            if k in col:
                w.writerow([k, units[col[k]], vals[col[k]]])
    print("wrote", dst)


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3])
