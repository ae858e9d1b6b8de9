"""Summarise an .ncu-rep (one kernel) into a small JSON for profiles/: selected raw metrics, warp-stall shares,
and per-gradient-evaluation DRAM traffic.   python tools/ncu_summary.py REPORT.ncu-rep OUT.json GRAD_EVALS ["what was captured"]"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__block_size", "launch__grid_size", "launch__cluster_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "smsp__pcsamp_sample_buffer_full", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep, out, evals = sys.argv[1], sys.argv[2], float(sys.argv[3])
    # optional 4th argument: what was captured (default: the bench.py capture of the headline kernel)
    how = sys.argv[4] if len(sys.argv) > 4 else (
        "ncu --set full --clock-control none --import-source on, one launch of the persistent kernel "
        "(bench.py --steps 3 --warmup 3 --no-cpu --power-iters 2: the captured launch is the 3-iteration "
        "timed solve = 4 gradient evaluations); timings under ncu are not bench values")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: {"unit": u, "value": v} for h, u, v in zip(hdr, units, vals)}
    res = {k: d[k] for k in KEEP if k in d}
    stalls = {}
    for k, v in d.items():
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
            try:
                stalls[k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(v["value"])
            except ValueError:
                pass
    tot = sum(stalls.values()) or 1.0
    res["_warp_stall_share"] = {k: round(v / tot, 4) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]) if v / tot >= 0.005}

    def to_bytes(m):
        if m not in d:
            return None
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[d[m]["unit"]]
        return float(d[m]["value"].replace(",", "")) * scale
    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    res["_derived"] = {"gradient_evaluations_in_launch": evals, "dram_bytes_read": rd, "dram_bytes_write": wr,
                       "dram_bytes_per_gradient_eval": (rd + wr) / evals,
                       "how": how}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
