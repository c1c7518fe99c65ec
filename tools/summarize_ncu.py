"""Turns an .ncu-rep capture of k_render into the committed summaries under profiles/.

  python tools/summarize_ncu.py gpurun_out/render.ncu-rep r01 4096
"""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'launch__registers_per_thread', 'launch__grid_size',
    'launch__block_size', 'launch__shared_mem_per_block_dynamic',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
]
UNIT = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}


def main(rep, tag, num_envs):
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  hdr, units, vals = rows[0], rows[1], rows[2]
  out = os.path.join(root, 'profiles', f'render_{tag}_ncu_summary.csv')
  rec = {}
  with open(out, 'w') as f:
    f.write('metric,unit,value\n')
    f.write(f'kernel,,{vals[hdr.index("Kernel Name")]}\n')
    for k in KEEP:
      if k in hdr:
        i = hdr.index(k)
        f.write(f'{k},{units[i]},{vals[i]}\n')
        rec[k] = (units[i], float(vals[i]))
  rd = rec['dram__bytes_read.sum']; wr = rec['dram__bytes_write.sum']
  traffic = rd[1] * UNIT[rd[0]] + wr[1] * UNIT[wr[0]]
  with open(os.path.join(root, 'profiles', 'render_traffic.json'), 'w') as f:
    json.dump({'source': os.path.basename(out), 'num_envs': int(num_envs), 'dram_bytes_per_launch': traffic,
               'dram_bytes_per_env': traffic / int(num_envs),
               'note': 'dram__bytes_read.sum + dram__bytes_write.sum of one k_render launch (ncu --set full)'}, f, indent=1)
  print(out, traffic)


if __name__ == '__main__':
  main(*sys.argv[1:4])
