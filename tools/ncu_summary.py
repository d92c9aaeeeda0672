"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into the JSON kept under profiles/.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.json [kernel-row-index]"""
import csv, io, json, subprocess, sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_dshared_op_st.sum", "sass__inst_executed_local_loads", "smsp__inst_executed.sum",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    row = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2 + row]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") or h in ("Kernel Name", "Grid Size", "Block Size"):
            d[h] = {"value": v, "unit": u}
    json.dump(d, open(out, "w"), indent=1, sort_keys=True)
    print("wrote", out, len(d), "metrics; kernel:", d.get("Kernel Name", {}).get("value"))


if __name__ == "__main__":
    main()
