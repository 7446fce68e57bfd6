"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of metrics
DESIGN.md argues from.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xxx.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "kernel time"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per warp instruction (of 32)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "CBU (branch) pipe %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit %"),
    ("l1tex__t_bytes.sum", "L1 bytes"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / SMSP"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (L1/L2/HBM)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (smem/local)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: lg throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: mio throttle"),
    ("smsp__sass_average_branch_targets_threads_uniform.pct", "uniform branch targets %"),
    ("sass__inst_executed_local_loads", "local loads (warp instr)"),
    ("sass__inst_executed_local_stores", "local stores (warp instr)"),
    ("sass__inst_executed_shared_loads", "shared loads (warp instr)"),
]


def traffic(rep, out_json):
    """profiles/traffic.json: DRAM bytes per launch (read + write, mean over the captured launches)."""
    import json
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, all_rows = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    # the frame launches of the step; the marches of the beam start (beam_start_kernel, one per frame) are listed beside them
    data = [r for r in all_rows if "beam_start" not in r[name_i]]
    march = [r for r in all_rows if "beam_start" in r[name_i]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def dram(r):
        b = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(k)
            b += float(r[i]) * scale[units[i]]
        return b
    tot = [dram(r) for r in data]
    inst = [int(float(r[hdr.index("smsp__inst_executed.sum")])) for r in data] if "smsp__inst_executed.sum" in hdr else []

    def col(key, conv=float, rows_=None):
        return [conv(r[hdr.index(key)]) for r in (data if rows_ is None else rows_)] if key in hdr else []
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from octree_ray_tracing_b200.build import kernel_source_hash
    json.dump({"dram_bytes_per_launch": int(sum(tot) / len(tot)), "per_launch": [int(x) for x in tot],
               "warp_instructions_per_launch": inst,
               "march_warp_instructions_per_launch": [int(x) for x in col("smsp__inst_executed.sum", float, march)],
               "march_kernel_time_us": col("gpu__time_duration.sum", float, march),
               "l2_sectors_per_launch": [int(x) for x in col("lts__t_sectors.sum")],
               "l1_hit_pct": col("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": col("lts__t_sector_hit_rate.pct"),
               "active_threads_per_warp_instruction": col("smsp__thread_inst_executed_per_inst_executed.ratio"),
               "issue_active_pct": col("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "kernel_time_us": col("gpu__time_duration.sum"),
               "source": rep, "kernel_source_sha16": kernel_source_hash(),
               "note": "written by tools/ncu_summary.py --traffic from an `ncu --set full` capture of the frame launches (poses A, B, C; with the beam start also "
                       "their three beam_start_kernel marches) of one bench step; bench.py uses it only while kernel_source_sha16 equals the hash of the sources it runs",
               "kernels": [r[name_i].split("(")[0] for r in data]}, open(out_json, "w"), indent=1)


def main():
    if len(sys.argv) > 3 and sys.argv[2] == "--traffic":
        traffic(sys.argv[1], sys.argv[3])
        return
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    print(f"# ncu summary of `{rep}`\n")
    print("launches: " + "; ".join(f"#{i} {r[name_i].split('(')[0]}" for i, r in enumerate(data)) + "\n")
    print("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
    print("|---|---|" + "---|" * len(data))
    for k, label in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"| {label} (`{k}`) | {units[i]} | " + " | ".join(r[i] for r in data) + " |")


if __name__ == "__main__":
    main()
