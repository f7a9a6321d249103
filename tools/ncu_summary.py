"""Compact text summary of an .ncu-rep (one launch): `name [unit] = value` lines for the metrics the roofline
discussion uses.  python tools/ncu_summary.py report.ncu-rep "header comment" > profiles/rNN_....txt"""
import csv, subprocess, sys

KEEP = ('Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active', 'sm__cycles_active.avg',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput', 'lts__t_sector_hit_rate', 'lts__t_bytes.sum',
        'lts__throughput', 'l1tex__data_pipe_lsu_wavefronts_mem_shared', 'launch__registers_per_thread', 'launch__block',
        'launch__grid_size', 'launch__shared_mem', 'launch__cluster', 'launch__occupancy_limit', 'sm__inst_executed.avg.per_cycle',
        'sm__throughput', 'smsp__warp_issue_stalled', 'sm__warps_active', 'smsp__cycles_active.avg', 'lts__t_sectors_srcunit_tex',
        'sm__inst_executed_pipe_uniform', 'smsp__inst_executed.sum', 'xbar', 'lts__t_sectors.sum', 'lts__t_sectors_op', 'lts__t_bytes')


def main():
  rep = sys.argv[1]
  for line in sys.argv[2:]:
    print('# ' + line)
  out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(out.splitlines()))
  hdr, units = rows[0], rows[1]
  for r in rows[2:]:
    for h, u, v in zip(hdr, units, r):
      if any(k in h for k in KEEP):
        print('%s [%s] = %s' % (h, u, v))


if __name__ == '__main__':
  main()
