"""Summarise an .ncu-rep (raw page CSV) into the few numbers the roofline needs."""
import csv
import re
import subprocess
import sys

KEYS = [
    r"^Kernel Name$", r"^Grid Size$", r"^Block Size$", r"gpu__time_duration.sum$", r"dram__bytes_read.sum$",
    r"dram__bytes_write.sum$", r"launch__registers_per_thread$", r"launch__occupancy_limit_registers$",
    r"launch__occupancy_limit_shared_mem$", r"launch__occupancy_limit_warps$", r"launch__waves_per_multiprocessor$",
    r"sm__warps_active.avg.pct_of_peak_sustained_active$", r"sm__throughput.avg.pct_of_peak_sustained_elapsed$",
    r"gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed$", r"sm__inst_executed_pipe_fp64.*pct",
    r"sm__pipe_fp64_cycles_active.*pct", r"^smsp__inst_executed.sum$", r"op_dfma_pred_on.sum$", r"op_dmul_pred_on.sum$",
    r"op_dadd_pred_on.sum$", r"mem_local_op_ld.sum$", r"mem_local_op_st.sum$", r"local_load|local_store",
    r"thread_inst_executed_per_inst_executed.ratio$", r"launch__shared_mem_per_block_dynamic$",
    r"^sm__cycles_elapsed.avg$", r"^lts__t_bytes.sum$", r"issue_stalled.*_per_warp_active.pct$",
    r"smsp__issue_active.avg.pct", r"smsp__cycles_active.avg$", r"l1tex__t_sector_hit_rate.pct", r"lts__t_sector_hit_rate.pct",
    r"sm__inst_executed_pipe_(lsu|alu|fma|fp64|xu).*sum$", r"smsp__inst_executed_op_shared", r"shared.*bank_conflict",
]


def main(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=" * 100)
        for i, h in enumerate(hdr):
            if any(re.search(k, h) for k in KEYS):
                print(f"{h[:95]:95s} {r[i]:>20s} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
