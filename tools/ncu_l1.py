"""L1TEX data-pipe budget of a kernel in an .ncu-rep: wavefronts per unit of work (global / shared ld / st / other)."""
import csv
import subprocess
import sys

rep, units = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    f = lambda k: float(d.get(k, "0").replace(",", "") or 0)
    sh, ld, st = f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), f("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum"), f("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum")
    tot = f("l1tex__data_pipe_lsu_wavefronts.sum")
    if tot == 0:
        tot = f("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg") * 148
    print("kernel", d.get("Kernel Name", "?")[:60], "time_ms", f("gpu__time_duration.sum"))
    print(f"  per unit: lsu wavefronts {tot / units:.0f}  global {(tot - sh) / units:.0f}  shared {sh / units:.0f} (ld {ld / units:.0f} st {st / units:.0f} other {(sh - ld - st) / units:.0f})"
          f"  tc {f('l1tex__data_pipe_tc_wavefronts_mem_shared.sum') / units:.0f}")
    print(f"  conflicts ld {f('l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum') / units:.0f} st {f('l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum') / units:.0f}"
          f"  lsu pipe busy {f('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):.1f}%  L1 hit {f('l1tex__t_sector_hit_rate.pct'):.1f}%"
          f"  issue {f('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f}%  cycles/unit {f('sm__cycles_elapsed.max') * 148 / units:.0f}"
          f"  inst/unit {f('smsp__inst_executed.sum') / units:.0f}  dram GB {(f('dram__bytes_read.sum') + f('dram__bytes_write.sum')):.2f}")
