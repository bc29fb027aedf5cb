"""Stall reasons of one `ncu --set full --import-source on` capture: totals per reason, and the top source lines with
their reason split.   python tools/ncu_stalls.py gpurun_out/prof.ncu-rep [n_lines]"""
import collections, csv, io, subprocess, sys

def main(rep, ntop=14):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if r and r[0] == 'Line No')
    cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = collections.Counter(); per = collections.defaultdict(collections.Counter); txt = {}
    cur = None
    for r in rows:
        if not r or r[0] in ('File Path', 'Function Name', 'Line No'):
            continue
        if r[0]:
            try:
                cur = (int(r[0]), r[1][:100])
            except ValueError:
                cur = None
            continue
        for i in cols:
            try:
                v = int(r[i])
            except (ValueError, IndexError):
                continue
            if v:
                tot[hdr[i]] += v; per[cur][hdr[i]] += v
    s = sum(tot.values())
    print('stall samples by reason:', ', '.join('%s %.1f%%' % (k[6:], 100.0 * v / s) for k, v in tot.most_common()))
    for cur, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values()))[:ntop]:
        n = sum(c.values())
        print('%5.1f%%  L%s: %s\n        %s' % (100.0 * n / s, cur[0] if cur else '?', cur[1] if cur else '', ', '.join('%s %d' % (k[6:], v) for k, v in c.most_common(4))))

if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 14)
