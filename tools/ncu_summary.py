"""Summarise an .ncu-rep (read here with `ncu -i`): per-kernel key metrics, optionally the hottest source lines.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--source N]"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct', 'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct',
        'launch__waves_per_multiprocessor', 'launch__grid_size', 'launch__block_size']


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('----', r[idx['Kernel Name']][:90])
        for w in WANT:
            if w in idx:
                print(f"  {w:66s} {r[idx[w]]:>22s} {units[idx[w]]}")
    if '--source' in sys.argv:
        n = int(sys.argv[sys.argv.index('--source') + 1])
        src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda'],
                             capture_output=True, text=True).stdout
        print(src[:200])


if __name__ == '__main__':
    main()
