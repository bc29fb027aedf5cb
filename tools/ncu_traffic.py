"""DRAM bytes per launch of one `ncu --set full` capture -> profiles/traffic_<workload>.json (read by bench.py for
`roofline.traffic`; carries a hash of binder_b200/csrc so that a capture of other sources is ignored).
    python tools/ncu_traffic.py gpurun_out/prof_config3.ncu-rep config3"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(rep, workload):
    import bench
    raw = list(csv.reader(io.StringIO(subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout)))
    hdr = next(r for r in raw if 'Kernel Name' in r)
    units, row = raw[raw.index(hdr) + 1], raw[raw.index(hdr) + 2]

    def val(name):
        i = hdr.index(name)
        v = float(row[i].replace(',', ''))
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[units[i]]
    rd, wr = val('dram__bytes_read.sum'), val('dram__bytes_write.sum')
    out = {'workload': workload, 'kernel': row[hdr.index('Kernel Name')], 'dram_bytes_read': rd, 'dram_bytes_write': wr,
           'dram_bytes_per_launch': rd + wr, 'gpu_time_us': float(row[hdr.index('gpu__time_duration.sum')].replace(',', '')),
           'capture': os.path.basename(rep), 'source_hash': bench.source_hash(),
           'note': 'one launch under ncu --set full (cold caches, serialised); a lone launch may leave part of its answers in L2 (dram write < bytes written)'}
    p = os.path.join(ROOT, 'profiles', 'traffic_%s.json' % workload)
    json.dump(out, open(p, 'w'), indent=1)
    print(p, out)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
