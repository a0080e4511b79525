"""Summarise an .ncu-rep (raw page) into the handful of counters we track under profiles/."""
import csv, subprocess, sys
KEEP = ['gpu__time_duration.sum','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem',
 'launch__grid_size','launch__block_size','sm__warps_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum',
 'lts__t_sector_hit_rate.pct','lts__t_bytes.sum','l1tex__t_sector_hit_rate.pct','sass__inst_executed_local_loads','sass__inst_executed_local_stores',
 'sass__inst_executed_shared_loads','sass__inst_executed_shared_stores','sass__inst_executed_global_loads','smsp__inst_executed.sum',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed_pipe_fp64.sum','sm__inst_executed_pipe_fp64.sum',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__cycles_elapsed.max']
def main(path, out=None):
    raw = subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units, vals = r[0], r[1], r[2]
    lines=[]
    for h,u,v in zip(hdr,units,vals):
        if h in KEEP or ('issue_stalled' in h and h.endswith('per_issue_active.ratio')) or h=='Kernel Name':
            lines.append(f"{h:95s} {u:16s} {v}")
    txt="\n".join(lines)
    if out: open(out,'w').write(f"# {path}\n"+txt+"\n")
    print(txt)
main(*sys.argv[1:])
