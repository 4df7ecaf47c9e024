"""Per-kernel summary of an `ncu --set full` report: python tools/summarize_ncu.py gpurun_out/prof_kernels.ncu-rep [title]"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index('Kernel Name')
    print(f'# {title}')
    for r in rows[2:]:
        name = r[kn].split('(')[0]
        print(f'\n== {name}')
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f'   {k:88s} {r[i]:>16s} {units[i]}')
        # warp stall reasons (cycles a warp waits per issued instruction), largest first
        stalls = []
        for i, k in enumerate(hdr):
            if 'issue_stalled' in k and k.endswith('_per_warp_active.pct') and 'not_issued' not in k:
                try:
                    stalls.append((float(r[i].replace(',', '')), k))
                except ValueError:
                    pass
        for v, k in sorted(stalls, reverse=True)[:8]:
            print(f'   stall {k[len("smsp__average_warps_issue_stalled_"):-len("_per_warp_active.pct")]:80s} {v:>16.2f} %')


if __name__ == '__main__':
    main()
