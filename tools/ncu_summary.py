"""Text summary of one `ncu --set full --import-source on` capture of resolve_kernel, as committed under
profiles/: key metrics, then stall samples and executed warp instructions per CUDA source line.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>_ncu_summary.txt
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
           'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'launch__grid_size',
           'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
           'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
           'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
           'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']


def ncu(rep, *args):
    return subprocess.run(['ncu', '-i', rep] + list(args), capture_output=True, text=True).stdout


def main(rep):
    raw = list(csv.reader(io.StringIO(ncu(rep, '--page', 'raw', '--csv'))))
    hdr = next(r for r in raw if 'Kernel Name' in r)
    units = raw[raw.index(hdr) + 1]
    row = raw[raw.index(hdr) + 2]
    print('kernel', row[hdr.index('Kernel Name')])
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            print(m, row[i], units[i])
    rows = list(csv.reader(io.StringIO(ncu(rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'))))
    inst, smp, txt = collections.Counter(), collections.Counter(), {}
    cur = None
    for r in rows:
        if not r or r[0] in ('File Path', 'Function Name', 'Line No'):
            continue
        if r[0]:
            try:
                cur = int(r[0]); txt[cur] = r[1]
            except ValueError:
                cur = None
        else:
            try:
                inst[cur] += int(r[7]); smp[cur] += int(r[4])
            except (ValueError, IndexError):
                pass
    ti, ts = sum(inst.values()), sum(smp.values())
    print('\n== per CUDA source line: stall samples / executed warp instructions (all passes of the capture) ==')
    print('samples', ts, 'inst', ti)
    for l, c in smp.most_common(30):
        print('%4d smp %4.1f%% inst %4.1f%%  L%d: %s' % (c, 100.0 * c / max(ts, 1), 100.0 * inst[l] / max(ti, 1), l, (txt.get(l) or '')[:110]))
    print('\n== lines by executed instructions ==')
    for l, c in inst.most_common(25):
        print('%5.1f%% inst %4d smp  L%d: %s' % (100.0 * c / max(ti, 1), smp[l], l, (txt.get(l) or '')[:110]))


if __name__ == '__main__':
    main(sys.argv[1])
