"""Summarise an .ncu-rep (raw page) into the handful of metrics we track.  python tools/ncu_summary.py rep [out.md]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [
 ('Kernel Name','kernel'),('gpu__time_duration.sum','time'),('launch__grid_size','grid'),('launch__block_size','block'),
 ('launch__registers_per_thread','regs'),('launch__occupancy_limit_registers','occ_lim_regs'),('launch__occupancy_limit_shared_mem','occ_lim_smem'),
 ('sm__warps_active.avg.pct_of_peak_sustained_active','achieved_occ_pct'),
 ('smsp__inst_executed.sum','warp_inst'),('smsp__issue_active.avg.pct_of_peak_sustained_active','issue_active_pct'),
 ('sm__inst_executed_pipe_fma.sum','pipe_fma_inst'),('sm__inst_executed_pipe_alu.sum','pipe_alu_inst'),('sm__inst_executed_pipe_fp64.sum','pipe_fp64_inst'),('sm__inst_executed_pipe_lsu.sum','pipe_lsu_inst'),('sm__inst_executed_pipe_xu.sum','pipe_xu_inst'),
 ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','fma_pipe_pct'),
 ('dram__bytes_read.sum','dram_rd'),('dram__bytes_write.sum','dram_wr'),('dram__throughput.avg.pct_of_peak_sustained_elapsed','dram_pct'),
 ('lts__t_sector_hit_rate.pct','l2_hit_pct'),('lts__t_bytes.sum','l2_bytes'),('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smem_bank_conflicts'),
 ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smem_wavefronts'),
 ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','st_long_sb'),
 ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','st_short_sb'),
 ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','st_barrier'),
 ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','st_mio'),
 ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','st_lg'),
 ('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','st_no_inst'),
 ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','st_math'),
 ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','st_wait'),
 ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','st_not_sel'),
 ('smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','st_dispatch'),
 ('smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','st_branch'),
 ('smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','st_membar'),
 ('smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio','st_sleep'),
]
out = []
for name, short in want:
    if name in hdr:
        i = hdr.index(name)
        vals = [r[i] for r in data]
        out.append(f"| {short} [{units[i]}] | " + " | ".join(v[:42] for v in vals) + " |")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
