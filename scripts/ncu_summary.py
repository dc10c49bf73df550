"""Summarise an ncu report (details page + instruction mix) into text for profiles/."""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(det)))
h = rows[0]
keep = ('Duration', 'Registers Per Thread', 'Theoretical Occupancy', 'Achieved Occupancy',
        'Executed Ipc Active', 'Issue Slots Busy', 'Compute (SM) Throughput', 'Memory Throughput',
        'DRAM Throughput', 'L1/TEX Hit Rate', 'L2 Hit Rate', 'Eligible Warps Per Scheduler',
        'Issued Warp Per Scheduler', 'No Eligible', 'Avg. Active Threads Per Warp',
        'Warp Cycles Per Issued Instruction', 'Block Limit Registers', 'Grid Size', 'Block Size',
        'Dynamic Shared Memory Per Block', 'Local Memory Spilling Requests')
for r in rows[1:]:
    d = dict(zip(h, r))
    if d.get('Metric Name') in keep:
        print('launch %s | %-40s | %s %s' % (d.get('ID'), d.get('Metric Name'), d.get('Metric Value'), d.get('Metric Unit')))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'smsp__inst_executed.sum',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']
for li, r in enumerate(rows[2:]):
    for i, hh in enumerate(hdr):
        if hh in want:
            print('launch %d | %-85s | %s %s' % (li, hh, r[i], units[i]))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
# first kernel only
hdr = rows[1]
ia, isrc = hdr.index('Instructions Executed'), hdr.index('Source')
tot = 0
byop = collections.Counter()
nstatic = 0
for r in rows[2:]:
    if len(r) <= ia or r[0] == 'Kernel Name':
        break
    try:
        n = int(r[ia])
    except ValueError:
        continue
    nstatic += 1
    s = re.sub(r'^@!?U?P\d+\s+', '', r[isrc].strip())
    op = (s.split()[0].rstrip(';') if s else '?').split('.')[0]
    byop[op] += n
    tot += n
print('launch 0 | static SASS instructions %d, executed warp instructions %d' % (nstatic, tot))
fp64 = sum(byop[o] for o in ('DFMA', 'DMUL', 'DADD', 'DSETP', 'DMNMX'))
print('launch 0 | FP64-pipe share of executed instructions %.1f%%' % (100.0 * fp64 / max(tot, 1)))
for op, n in byop.most_common(16):
    print('launch 0 | op %-8s %6.2f%%' % (op, 100.0 * n / tot))
