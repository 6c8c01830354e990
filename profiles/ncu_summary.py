import csv, os, collections, sys
rep=sys.argv[1]; kern=sys.argv[2] if len(sys.argv)>2 else None
import subprocess
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
idx={h:i for i,h in enumerate(hdr)}
want=["Kernel Name","gpu__time_duration.sum","smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","l1tex__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed","l1tex__t_sector_hit_rate.pct","dram__bytes_read.sum","dram__bytes_write.sum","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio","smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio","smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio","smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio","smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio"]
for r in rows[2:]:
    if kern and kern not in r[idx["Kernel Name"]]: continue
    for w in want:
        if w in idx: print("  %-95s %s %s"%(w,r[idx[w]][:60],units[idx[w]]))
    print()
