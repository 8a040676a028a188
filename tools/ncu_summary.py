"""Extract the judged metrics from .ncu-rep captures into small text files under profiles/.
    python tools/ncu_summary.py gpurun_out/prof_split_c3.ncu-rep profiles/r01_split_c3.txt
"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'l1tex__t_bytes.sum', 'lts__t_bytes.sum']


def main(rep, out):
    raw = open(rep).read() if rep.endswith('.csv') else subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, 'w') as fh:
        fh.write(f'# ncu --set full --clock-control none, source: {rep}\n')
        for vals in rows[2:]:
            name = vals[hdr.index('Kernel Name')]
            fh.write(f'\nkernel: {name}\n')
            for i, h in enumerate(hdr):
                if h in WANT:
                    fh.write(f'  {h:80s} {vals[i]:>18s} {units[i]}\n')
    print(open(out).read())


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
