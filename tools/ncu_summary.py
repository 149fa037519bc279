"""Key numbers of an ncu report (raw page): python tools/ncu_summary.py X.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__cycles_elapsed.max", "smsp__sass_inst_executed_op_shared_ld.sum",
        "smsp__sass_inst_executed_op_shared_st.sum", "launch__grid_size", "launch__block_size"]
for r in rows[2:]:
    print("---", r[hdr.index("Kernel Name")][:90])
    for w in want:
        if w in hdr:
            print(f"  {w:72s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}")
    st = {h: float(r[i]) for i, h in enumerate(hdr)
          if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and r[i]}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:9]
    print("  stalls/issue:", ", ".join(f"{k.split('stalled_')[1].split('_per_')[0]}={v:.2f}" for k, v in top))
