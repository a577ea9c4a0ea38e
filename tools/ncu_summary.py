#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): key raw metrics + top stall sites from the source page.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.txt]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__inst_issued.avg.pct_of_peak_sustained_active', 'sm__inst_issued.avg.per_cycle_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_barriers', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum', 'lts__t_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
keep += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled')]
# the pipe the FFT kernel computes on (FP64), the pipes around it, and the L2 side of the key stream
keep += [h for h in hdr if h.startswith(('sm__pipe_fp64', 'sm__inst_executed_pipe_fp64', 'smsp__inst_executed_pipe_fp64', 'sm__inst_executed_pipe_fma', 'sm__inst_executed_pipe_alu',
                                         'sm__inst_executed_pipe_uniform', 'sm__inst_executed_pipe_tensor', 'lts__t_bytes', 'lts__t_sectors_op_read.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
                                         'smsp__inst_executed_op_shared', 'sm__sass_inst_executed_op_shared'))]
print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?", file=out)
for h, u, v in zip(hdr, units, vals):
    if h in keep:
        print(f"{h} [{u}] = {v}", file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(h2)}
def g(r, h):
    try: return float(r[ix[h]])
    except Exception: return 0.0
tot = sum(g(r, '# Samples') for r in data)
print(f"\ninstructions in kernel: {len(data)}   samples: {int(tot)}", file=out)
for key in ['stall_long_sb', 'stall_no_inst', 'stall_wait', 'stall_math', 'stall_barrier', 'stall_short_sb', 'stall_mio', 'stall_dispatch', 'stall_lg', 'stall_not_selected', 'stall_selected']:
    s = sum(g(r, key) for r in data)
    print(f"== {key}: {int(s)} samples ({100*s/max(tot,1):.1f}%)", file=out)
    for r in sorted(data, key=lambda r: -g(r, key))[:6]:
        print(f"   {int(g(r,key)):7d}  {r[ix['Address']][-5:]}  {r[ix['Source']].strip()[:80]}", file=out)
print("\n== shared-memory bank conflicts (excess wavefronts) by instruction", file=out)
for r in sorted(data, key=lambda r: -g(r, 'L1 Wavefronts Shared Excessive'))[:10]:
    if g(r, 'L1 Wavefronts Shared Excessive') > 0:
        print(f"   excess {int(g(r,'L1 Wavefronts Shared Excessive')):10d} of {int(g(r,'L1 Wavefronts Shared')):10d}  {r[ix['Source']].strip()[:70]}", file=out)
