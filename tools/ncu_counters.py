"""ncu CSV (per-kernel metric rows) -> profiles/ncu_counters.json, the committed counters bench.py's roofline reads.

    ncu --metrics <list below> --clock-control none -k regex:'kino_|reduce_f' -s 9 -c 3 --csv \\
        --log-file gpurun_out/counters.csv python tools/time_kino.py
    python tools/ncu_counters.py gpurun_out/counters.csv profiles/r02/ncu_counters_vNN.csv

(-s 9: the three launches captured are one full-mask evaluation after time_kino's warm-up.)  The CSV itself is
copied next to the JSON so that every number in the JSON can be re-derived from profiles/ alone."""
import csv
import json
import os
import shutil
import sys

METRICS = ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,"
           "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,"
           "gpu__time_duration.sum,sm__cycles_elapsed.avg,dram__bytes_read.sum,dram__bytes_write.sum,"
           "sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,"
           "smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,"
           "launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__occupancy_limit_registers,"
           "launch__occupancy_limit_shared_mem")
KEYS = {"smsp__sass_thread_inst_executed_op_dfma_pred_on.sum": "dfma", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum": "dadd",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum": "dmul", "smsp__inst_executed.sum": "inst_executed",
        "smsp__thread_inst_executed.sum": "thread_inst_executed", "gpu__time_duration.sum": "time_ns",
        "sm__cycles_elapsed.avg": "sm_cycles", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "sm__inst_executed_pipe_fp64.sum": "inst_fp64",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "launch__registers_per_thread": "registers", "launch__shared_mem_per_block_dynamic": "smem_per_block",
        "launch__occupancy_limit_registers": "occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem": "occupancy_limit_shared_mem"}
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0, "second": 1e9}


def main(src, copy_to=None, knot_evals=30720):
    rows = list(csv.reader(open(src)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    kernels = {}
    for r in rows[hdr + 1:]:
        if len(r) < len(H):
            continue
        rec = dict(zip(H, r))
        name = rec["Kernel Name"].split("(")[0].replace("void ", "").replace("hb::", "").strip()
        key = KEYS.get(rec["Metric Name"])
        if key is None:
            continue
        val = float(rec["Metric Value"].replace(",", "")) * SCALE.get(rec["Metric Unit"], 1.0)
        kernels.setdefault(name, {})[key] = val
    for k in kernels.values():
        if {"dfma", "dadd", "dmul"} <= set(k):
            k["flops"] = 2 * k["dfma"] + k["dadd"] + k["dmul"]
            k["flops_per_knot_eval"] = k["flops"] / knot_evals
    out = {"knot_evals_per_launch": knot_evals, "kernels": kernels,
           "source": f"ncu per-kernel counters of one full-mask hb_eval (B=1024, N=30) under tools/time_kino.py; CSV: "
                     f"{copy_to or src}"}
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_counters.json")
    json.dump(out, open(dst, "w"), indent=1)
    if copy_to:
        shutil.copy(src, copy_to)
    for n, k in kernels.items():
        print(n, {a: (round(b, 1) if isinstance(b, float) else b) for a, b in k.items() if a in ("time_ns", "flops_per_knot_eval", "fp64_pipe_active_pct", "issue_active_pct", "registers")})


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
