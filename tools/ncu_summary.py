"""Compact, committable summary of an .ncu-rep (read on the CPU box): one block per profiled launch with the
metrics the roofline needs (duration, DRAM bytes, pipe utilisation, occupancy, registers, stall mix).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/prof_r1.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}")
    for r in rows[2:]:
        print(f"\n== launch id {r[idx['ID']]}: {r[idx['Kernel Name']][:110]}")
        print(f"   grid {r[idx.get('Grid Size', 0)]} block {r[idx.get('Block Size', 0)]}")
        for k in KEYS:
            if k in idx and r[idx[k]] != "":
                print(f"   {k:72s} {r[idx[k]]:>18s} {units[idx[k]]}")
        stalls = [(h, float(r[i] or 0)) for h, i in idx.items()
                  if h.startswith("smsp__average_warp_latency_issue_stalled") or h.startswith("smsp__average_warps_issue_stalled")]
        stalls = [(h, v) for h, v in stalls if v > 0 and "not_issued" not in h]
        stalls.sort(key=lambda kv: -kv[1])
        for h, v in stalls[:8]:
            print(f"   stall {h.split('issue_stalled_')[-1][:60]:66s} {v:18.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
