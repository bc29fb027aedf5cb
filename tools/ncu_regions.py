"""Executed warp instructions and stall samples of one `ncu --set full --import-source on` capture, per source FILE and
FUNCTION (the enclosing definition of each line), then the hottest lines with their file.
    python tools/ncu_regions.py gpurun_out/prof.ncu-rep [n_lines]"""
import collections, csv, io, re, subprocess, sys

def main(rep, ntop=25):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    rows = csv.reader(io.StringIO(out))
    fname = None; hdr = None; cur = None; func = None
    inst = collections.Counter(); smp = collections.Counter(); line_i = collections.Counter(); line_s = collections.Counter(); txt = {}
    start = re.compile(r'^(template|__device__|__global__|struct |static |inline |BB_HD|extern|int |void |const )')
    for r in rows:
        if not r:
            continue
        if r[0] in ('File Name', 'File Path'):
            fname = r[1].split('/')[-1]; func = '(top)'; continue
        if r[0] == 'Line No':
            hdr = r; ci = hdr.index('Instructions Executed'); cs = hdr.index('# Samples'); continue
        if hdr is None:
            continue
        if r[0]:
            src = r[1]
            if start.match(src) and ('(' in src or src.startswith('struct')) and not src.rstrip().endswith(';'):
                m = re.search(r'([A-Za-z_0-9]+)\s*\(', re.sub(r'__launch_bounds__\([^)]*\)|__align__\(\d+\)', '', src)) if not src.startswith('struct') else re.search(r'struct\s+(\w+)', src)
                func = m.group(1) if m else src[:40]
            try:
                cur = (fname, int(r[0])); txt[cur] = src[:110]
            except ValueError:
                cur = None
            continue
        if cur is None:
            continue
        try:
            i = int(r[ci]); s = int(r[cs])
        except (ValueError, IndexError):
            continue
        inst[(fname, func)] += i; smp[(fname, func)] += s; line_i[cur] += i; line_s[cur] += s
    ti = sum(inst.values()) or 1; ts = sum(smp.values()) or 1
    print('total warp instructions %d, stall samples %d' % (ti, ts))
    print('== by function: %inst  %samples')
    for k, v in inst.most_common(30):
        print('%5.1f%% %5.1f%%  %s:%s' % (100.0 * v / ti, 100.0 * smp[k] / ts, k[0], k[1]))
    print('== hottest lines by samples')
    for k, v in line_s.most_common(ntop):
        print('%5.1f%% smp %5.1f%% inst  %s:%d  %s' % (100.0 * v / ts, 100.0 * line_i[k] / ti, k[0], k[1], txt.get(k, '')))

if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
