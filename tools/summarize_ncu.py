"""Turns an .ncu-rep capture of one kernel into the committed summaries under profiles/.

  python tools/summarize_ncu.py gpurun_out/render_r02.ncu-rep render_r02 [clean_up 7 4096]
With the substrate / players / envs given (a k_render capture), also records the DRAM traffic per env in
profiles/render_traffic.json together with the hash of the kernel source it was measured on (bench.py reports
`roofline.traffic` only while that hash still matches).
"""
import csv
import hashlib
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
    'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
    'launch__occupancy_limit_registers',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
]
UNIT = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}


def source_hash(root):
  h = hashlib.sha1()
  for name in ('render.cuh', 'common.cuh'):
    with open(os.path.join(root, 'meltingpot_b200', 'csrc', name), 'rb') as f:
      h.update(f.read())
  return h.hexdigest()


def main(rep, tag, substrate=None, players=None, num_envs=None):
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  hdr, units, vals = rows[0], rows[1], rows[2]
  out = os.path.join(root, 'profiles', f'{tag}_ncu_summary.csv')
  rec = {}
  with open(out, 'w') as f:
    f.write('metric,unit,value\n')
    f.write(f'kernel,,{vals[hdr.index("Kernel Name")]}\n')
    for k in KEEP:
      if k in hdr:
        i = hdr.index(k)
        f.write(f'{k},{units[i]},{vals[i]}\n')
        rec[k] = (units[i], float(vals[i]))
  print(out)
  if substrate:
    rd = rec['dram__bytes_read.sum']; wr = rec['dram__bytes_write.sum']
    traffic = rd[1] * UNIT[rd[0]] + wr[1] * UNIT[wr[0]]
    path = os.path.join(root, 'profiles', 'render_traffic.json')
    doc = {'note': 'dram__bytes_read.sum + dram__bytes_write.sum of one k_render launch (ncu --set full), per env; valid for the '
                   'kernel source whose sha1 (render.cuh + common.cuh) is recorded', 'captures': {}}
    if os.path.exists(path):
      old = json.load(open(path))
      if 'captures' in old:
        doc = old
    doc['captures'][f'{substrate}__{players}p'] = {
        'source': os.path.basename(out), 'num_envs': int(num_envs), 'dram_bytes_per_launch': traffic,
        'dram_bytes_per_env': traffic / int(num_envs), 'source_sha1': source_hash(root)}
    with open(path, 'w') as f:
      json.dump(doc, f, indent=1)
    print(path, traffic)


if __name__ == '__main__':
  main(*sys.argv[1:6])
