#!/usr/bin/env python
"""bench.py — DNS queries/sec of the batched resolve path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload configX] [--mode auto|shard|replicas|nccl]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path (parse -> zone lookup -> answer bytes) over one batch.

  N=1   default workload = BASELINE.json configs[2] ("config3", the largest single-GPU configuration):
        10M-record zone, 262,144-query batches, 50 % SRV / 50 % A on service names.  `--workload config2`
        (1M zone, 65,536 A lookups) and `config5` (90 % misses) run the other single-GPU shapes.
  N>1   default workload = configs[3] ("config4"): the same 10M zone hash-sharded over the ranks, a global batch
        of 1,048,576 queries split evenly over the ranks' ingress, 60 % A / 20 % SRV / 20 % AAAA (bench_multi.py).

  value      kernel path, batches resident in HBM, 4 independent batches in flight (one stream each), timed
             between two CUDA events on the stream the side streams fork from and join into.  Steady state:
             after >= 50 ms of device warm-up, R = 15 regions of K steps each are timed and the MEDIAN region
             is reported, so the number does not depend on K.  The steps cycle over a ring of distinct batches
             whose footprint (inputs + outputs) exceeds L2 — no L2 flush needed.
  e2e        the same metric through bb_resolve_submit/_wait (the C ABI a host calls) with pinned HOST buffers:
             H2D of packets+offsets and the answers landing in host memory inside the timed region.
  roofline   SURVEY.md §8(d) algorithmic HBM bytes of one launch (computed from the actual batch and its actual
             answers) / that kernel's launch duration (CUDA events; one launch at a time) vs MEASURED_PEAKS.json.
  cpu_baseline  the CPU oracle (a C++ port of lib/server.js + lib/zk.js — the Node.js reference cannot run in
             this image) on the box's host cores, bounded sample; plus the same word-wise algorithm over the
             same table image on the CPU ("same_table") when that harness is built.

--impl reference times the CPU port alone (it is the only reference arm that exists here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'dns_queries_per_sec'
UNIT = 'queries/s'
SEED = 0xB1DDE5
DEFAULTS = {   # workload: (zone records, batch, response bytes reserved per query)
    'config2': (1000000, 65536, 96),
    'config3': (10000000, 262144, 512),
    'config4': (10000000, 1048576, 320),
    'config5': (10000000, 262144, 96),
    'config5_rd0': (10000000, 262144, 96),
}
L2_BYTES = 126e6
REGIONS = 15


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version there) must not
    add to it: fd 1 is pointed at stderr for the whole run and only emit() writes to the real one."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, dev=0):
        self.rows, self.proc, self.dev = [], None, dev

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.dev), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == 'Active' for r in self.rows)]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p)).get('hbm_gbs', 6650.0), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def source_hash():
    """Identifies the kernel sources a profile was captured from (profiles/*traffic*.json carry it)."""
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(ROOT, 'binder_b200', 'csrc')
    for f in sorted(os.listdir(d)):
        h.update(open(os.path.join(d, f), 'rb').read())
    return h.hexdigest()[:16]


def measured_traffic(workload):
    """DRAM bytes per launch from an `ncu --set full` capture of THIS source and THIS workload
    (tools/ncu_summary.py writes profiles/traffic_<workload>.json with the source hash), else None."""
    p = os.path.join(ROOT, 'profiles', 'traffic_%s.json' % workload)
    if os.path.exists(p):
        t = json.load(open(p))
        if t.get('source_hash') == source_hash():
            return t.get('dram_bytes_per_launch')
    return None


class OracleLoader(object):
    """Loads the CPU oracle (test infrastructure, oracle/) in a background thread while the GPU side
    sets up: a 10M-record zone takes ~100 s to load into its JSON DOM."""

    def __init__(self, zone, recursion):
        self.zone, self.recursion, self.orc, self.err, self.secs = zone, recursion, None, None, 0.0
        self.same, self.same_err, self.same_secs = None, None, 0.0
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def _run(self):
        try:
            sys.path.insert(0, os.path.join(ROOT, 'oracle'))
            from oracle_lib import Oracle
            t0 = time.time()
            self.orc = Oracle(self.zone.dns_domain, self.zone.datacenter, self.recursion, snapshot=self.zone.jsonl)
            self.secs = time.time() - t0
        except Exception as ex:          # reported by get()
            self.err = ex
            return
        try:                             # the second CPU arm: the same table image and algorithm on the host (oracle/same_table.py)
            import same_table
            t0 = time.time()
            self.same = same_table.SameTable(self.zone.dns_domain, self.zone.jsonl, self.recursion)
            self.same_secs = time.time() - t0
        except Exception as ex:
            self.same_err = repr(ex)

    def get(self):
        self.th.join()
        if self.err:
            raise self.err
        return self.orc


def cpu_port(orc, data, off, budget_s=12.0):
    """Oracle (C++ port of the reference path) on the host cores, bounded sample."""
    cores = os.cpu_count() or 1
    n = len(off) - 1
    ns = min(n, 65536)                       # single-thread leg on a slice: it is ~1 M q/s
    t1 = orc.timed_resolve(data[:int(off[ns]) + 16], off[:ns + 1], nthreads=1, repeat=2)
    est = min(t1) * n / ns / max(cores * 0.5, 1)
    reps = max(2, min(40, int(budget_s / max(est, 1e-3))))
    tn = orc.timed_resolve(data, off, nthreads=cores, repeat=reps)
    return {'value': n / min(tn), 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': '%d x one %d-query batch of the same workload, best call; all %d host threads '
                      '(single thread: %.0f queries/s on %d queries)' % (reps, n, cores, ns / min(t1), ns),
            'single_thread_value': ns / min(t1),
            'note': 'literal port: JSON-DOM walk + std::unordered_map<std::string> per query (oracle/oracle.cpp, -O2)'}


def cpu_same_table(st, data, off, want_bytes, want_miss, per_q, seed=0, budget_s=8.0, against='the GPU result'):
    """SURVEY.md 8(d) "same open-addressed layout so the comparison isolates the processor": the word-wise device code
    compiled for the host (-O3) over the same zone image, all host threads; bounded sample."""
    cores = os.cpu_count() or 1
    n = len(off) - 1
    ns = min(n, 65536)
    s1, _, _ = st.timed_resolve(data[:int(off[ns]) + 16], off[:ns + 1], nthreads=1, repeat=1, resp_cap=per_q)
    est = s1 * n / ns / max(cores * 0.5, 1)
    reps = max(2, min(30, int(budget_s / max(est, 1e-3))))
    sn, tb, tm = st.timed_resolve(data, off, seed=seed, nthreads=cores, repeat=reps, resp_cap=per_q)
    return {'value': n / sn, 'unit': UNIT, 'cores': cores, 'kind': 'port (same table image, same word-wise algorithm: the device source compiled for the host)',
            'sample': '%d x one %d-query batch of the same workload, best pass; all %d host threads (single thread: %.0f queries/s on %d queries)'
                      % (reps, n, cores, ns / s1, ns),
            'single_thread_value': ns / s1,
            'parity': 'response bytes %d and misses %d %s %s' % (tb, tm, 'equal' if (tb, tm) == (want_bytes, want_miss) else 'DIFFER from', against)}


def workload_setup(args):
    from binder_b200 import synth
    desc, service_frac, mix, miss_frac, recursion = synth.WORKLOADS[args.workload]
    zr, b, per_q = DEFAULTS[args.workload]
    zone_records = args.zone_records or zr
    batch = args.batch or b
    if (zone_records, batch) != (zr, b):
        desc += ' [run at zone_records=%d, batch=%d]' % (zone_records, batch)
    t0 = time.time()
    zone = synth.gen_zone(zone_records, service_frac=service_frac)
    log('[bench] generated %d-record zone (%d hosts, %d services) in %.1fs' % (zone.n_records, zone.n_hosts, zone.n_services, time.time() - t0))
    return zone, desc, mix, miss_frac, recursion, batch, per_q


def run_reference(args):
    """--impl reference: the CPU port of the reference path (Node.js cannot run here), all host threads,
    on the b200 arm's workload; each step = one bounded sample (up to 262,144 queries: enough to amortise the thread start-up
    of a step over 128+ threads) of that workload.  The line also carries the second CPU arm (same table image and algorithm
    on the host) under cpu_baseline.same_table; `value` stays the literal port's."""
    from binder_b200 import synth
    zone, desc, mix, miss_frac, recursion, batch, _ = workload_setup(args)
    if args.gpus > 1:
        batch = max(batch // args.gpus, 1)
    sample = min(batch, 262144)
    data, off, _ = synth.gen_batch(zone, sample, 1000, mix, miss_frac, rd=synth.WORKLOAD_RD.get(args.workload, True))
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    from oracle_lib import Oracle
    t0 = time.time()
    orc = Oracle(zone.dns_domain, zone.datacenter, recursion, snapshot=zone.jsonl)
    log('[reference] oracle loaded in %.1fs' % (time.time() - t0))
    cores = os.cpu_count() or 1
    orc.timed_resolve(data, off, nthreads=cores, repeat=max(args.warmup, 1))
    ts = orc.timed_resolve(data, off, nthreads=cores, repeat=args.steps)
    total = sum(ts)
    v = sample * args.steps / total
    same = None
    try:
        import same_table
        t0 = time.time()
        st = same_table.SameTable(zone.dns_domain, zone.jsonl, recursion)
        o = orc.resolve_batch(data, off, seed=SEED, nthreads=cores)
        same = cpu_same_table(st, data, off, len(o[0]), len(o[4]), 1232, seed=SEED, budget_s=5.0, against="the literal port's")
        same['zone_build_s'] = time.time() - t0
    except Exception as ex:
        same = {'unavailable': repr(ex)}
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps, 'higher_is_better': True,
            'scaling': 'strong' if args.workload == 'config4' else 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
            'config': {'workload': desc, 'zone_records': zone.n_records, 'batch': batch,
                       'note': 'CPU restatement of lib/server.js + lib/zk.js + mname codec (oracle/oracle.cpp); '
                               'the Node.js reference is not runnable in this image; each step resolves a '
                               '%d-query sample of the workload on all host threads' % sample},
            'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': '%d steps x one %d-query sample of the workload, all %d host threads' % (args.steps, sample, cores),
                             'same_table': same},
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    emit(line)


class KernelPath(object):
    """Device-resident timing of one workload on one engine: ring of distinct batches, IN_FLIGHT streams,
    median-of-regions steady state, per-launch duration, SURVEY.md §8(d) bytes of a real launch."""
    IN_FLIGHT = 4

    def __init__(self, eng, zone, workload, B, per_q, dev, seed0=1000):
        import torch
        from binder_b200 import synth
        self.torch, self.eng, self.zone, self.B, self.dev, self.wl = torch, eng, zone, B, dev, workload
        _, _, mix, miss_frac, _ = synth.WORKLOADS[workload]
        self.out_cap = B * per_q
        self.ring_n = int(min(24, max(3, np.ceil(1.5 * L2_BYTES / (B * (56 + per_q * 0.6))))))
        self.ring = [synth.gen_batch(zone, B, seed0 + r, mix, miss_frac, rd=synth.WORKLOAD_RD.get(workload, True)) for r in range(self.ring_n)]
        self.d = []
        for data, off, _ in self.ring:
            self.d.append(dict(
                pk=torch.from_numpy(data).to(dev), off=torch.from_numpy(off.view(np.int32)).to(dev),
                out=torch.empty(self.out_cap, dtype=torch.uint8, device=dev), oo=torch.empty(B + 1, dtype=torch.int32, device=dev),
                st=torch.empty(B, dtype=torch.uint8, device=dev), ms=torch.empty(B, dtype=torch.int32, device=dev),
                ol=torch.empty(B, dtype=torch.int16, device=dev), tot=torch.zeros(4, dtype=torch.int32, device=dev)))
        self.stream = torch.cuda.current_stream()
        self.side = [torch.cuda.Stream(device=dev) for _ in range(self.IN_FLIGHT)]

    def step_on(self, k, cs):
        b = self.d[k % self.ring_n]
        self.eng.resolve_device(b['pk'].data_ptr(), b['off'].data_ptr(), self.B, SEED, 0, b['out'].data_ptr(), self.out_cap,
                                b['oo'].data_ptr(), b['ol'].data_ptr(), b['st'].data_ptr(), b['ms'].data_ptr(), b['tot'].data_ptr(), cs)

    def timed_concurrent(self, k0, nsteps, depth=None):
        """nsteps steps spread round-robin over `depth` side streams (= batches in flight), timed on the stream they fork
        from and join into."""
        torch, stream = self.torch, self.stream
        side = self.side[:depth or self.IN_FLIGHT]
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(stream)
        for st in side:
            st.wait_event(ea)
        for k in range(nsteps):
            self.step_on(k0 + k, side[k % len(side)].cuda_stream)
        for st in side:
            ev = torch.cuda.Event()
            ev.record(st)
            stream.wait_event(ev)
        eb.record(stream)
        torch.cuda.synchronize()
        return ea.elapsed_time(eb)

    def measure(self, K, warmup):
        """-> dict(value, ms_per_step, regions, warm_ms, serial_ms, graph_ms, kern_ms)"""
        torch, stream = self.torch, self.stream
        for k in range(warmup):
            self.step_on(k, stream.cuda_stream)
        torch.cuda.synchronize()
        warm_ms, kk = 0.0, 0
        while warm_ms < 50.0:                    # >= 50 ms of device work before anything is timed
            warm_ms += self.timed_concurrent(kk, max(K, self.IN_FLIGHT * 2))
            kk += K
        regions = [self.timed_concurrent(kk + r * K, K) for r in range(REGIONS)]
        ms_region = float(np.median(regions))
        depth = self.IN_FLIGHT
        # A kernel that already fills the GPU for several waves (a 262,144-query batch of long answers) gains nothing from a
        # second batch in flight and can lose to the interleaving: the same regions with ONE batch in flight, and the
        # better pipelining depth is the one reported (both are recorded)
        regions1 = [self.timed_concurrent(kk + (REGIONS + r) * K, K, depth=1) for r in range(REGIONS)]
        self.depth_ms = {self.IN_FLIGHT: ms_region / K, 1: float(np.median(regions1)) / K}
        if float(np.median(regions1)) < ms_region:
            regions, ms_region, depth = regions1, float(np.median(regions1)), 1
        self.depth = depth
        # the same K steps strictly one after another on one stream: per-launch duration for the roofline
        serial = []
        for r in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for k in range(K):
                self.step_on(r * K + k, stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            serial.append(e0.elapsed_time(e1) / K)
        serial_ms = float(np.median(serial))
        # the same loop replayed as one CUDA graph: removes host launch gaps, so it is the kernel's own
        # average duration (what the roofline uses)
        graph_ms = None
        try:
            gsteps = min(max(K, 20), 200)
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            self.step_on(0, side.cuda_stream)    # the engine allocates this stream's scratch outside the capture
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=side):
                cs = torch.cuda.current_stream().cuda_stream
                for k in range(gsteps):
                    self.step_on(k, cs)
            g.replay()
            torch.cuda.synchronize()
            reps = []
            for _ in range(5):
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                g.replay()
                g1.record()
                torch.cuda.synchronize()
                reps.append(g0.elapsed_time(g1) / gsteps)
            graph_ms = float(np.median(reps))
        except Exception as ex:      # capture is an optimisation of the measurement, not a requirement
            log('[bench] CUDA-graph replay unavailable: %r' % (ex,))
        ms_per_step = ms_region / K
        return dict(value=self.B / (ms_per_step * 1e-3), ms_per_step=ms_per_step, regions=regions, warm_ms=warm_ms,
                    serial_ms=serial_ms, graph_ms=graph_ms, kern_ms=min(serial_ms, graph_ms) if graph_ms else serial_ms)

    def result0(self):
        """Batch 0 resolved once more -> host copies of its results (out_off, out_len, status, totals, out)."""
        self.step_on(0, self.stream.cuda_stream)
        self.torch.cuda.synchronize()
        b0 = self.d[0]
        oo = b0['oo'].cpu().numpy().view(np.uint32)
        ol = b0['ol'].cpu().numpy().view(np.uint16)
        tot = b0['tot'].cpu().numpy()
        assert tot[0] == oo[self.B], 'totals disagree with the offset array'
        assert int(ol.astype(np.int64).sum()) == int(tot[0]), 'lengths disagree with the total'
        return dict(oo=oo, ol=ol, st=b0['st'].cpu().numpy(), tot=tot, out=b0['out'].cpu().numpy(),
                    miss=np.sort(b0['ms'].cpu().numpy().view(np.uint32)[:int(tot[1])]))

    def roofline(self, m, r0):
        from binder_b200 import synth
        rd_b, wr_b = synth.algorithmic_bytes(self.zone, self.ring[0][1], self.ring[0][2], r0['ol'])
        peak, peak_src = measured_peaks()
        kern_ms = m['kern_ms']
        achieved = (rd_b + wr_b) / (kern_ms * 1e-3) / 1e9
        return {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': measured_traffic(self.wl), 'peak_source': peak_src, 'kernel': 'bbk::resolve_kernel',
                'kernel_ms': kern_ms,
                'kernel_timing': 'one launch at a time, CUDA events: ' + ('CUDA-graph replay' if m['graph_ms'] and m['graph_ms'] <= m['serial_ms'] else 'stream loop'),
                'algorithmic_bytes_per_launch': rd_b + wr_b, 'read_bytes': rd_b, 'write_bytes': wr_b,
                'bytes_per_query': (rd_b + wr_b) / self.B,
                'read_only_frac': rd_b / (kern_ms * 1e-3) / 1e9 / peak,
                'in_flight_frac': (rd_b + wr_b) / (m['ms_per_step'] * 1e-3) / 1e9 / peak}

    def check_oracle(self, orc, r0):
        from binder_b200.engine import repack
        o = orc.resolve_batch(self.ring[0][0], self.ring[0][1], seed=SEED, nthreads=os.cpu_count() or 1)
        got, goff = repack(r0['out'], r0['oo'], r0['ol'])
        assert np.array_equal(got, o[0]) and np.array_equal(goff, o[1]), 'GPU answers differ from the CPU oracle (%s)' % self.wl
        assert np.array_equal(r0['st'], o[3]), 'GPU statuses differ from the CPU oracle (%s)' % self.wl
        assert np.array_equal(r0['miss'], o[4]), 'miss list differs from the CPU oracle (%s)' % self.wl
        return 'bit-exact vs oracle: all %d responses of a timed batch (%d bytes), statuses and miss list' % (self.B, len(got))

    def free(self):
        self.d = None
        self.torch.cuda.empty_cache()


def run_single(args, local_rank):
    import ctypes
    import torch
    from binder_b200 import synth
    from binder_b200.engine import Engine, repack
    from binder_b200 import build as bbuild
    bbuild.build()
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device — binder_b200 has no CPU path to benchmark')
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    wl = args.workload
    zone, desc, mix, miss_frac, recursion, B, per_q = workload_setup(args)
    loader = None if args.no_cpu else OracleLoader(zone, recursion)
    # secondary workloads measured on the same engine (kernel path + roofline only): same-zone ones only
    if args.also is None:
        args.also = {'config3': 'config4', 'config5': 'config5_rd0'}.get(wl, 'none')
    also = [w for w in (args.also.split(',') if args.also and args.also != 'none' else [])
            if w in synth.WORKLOADS and w != wl and synth.WORKLOADS[w][1] == synth.WORKLOADS[wl][1] and synth.WORKLOADS[w][4] == recursion
            and not args.zone_records and not args.batch]
    maxB = max([B] + [DEFAULTS[w][1] for w in also])

    t0 = time.time()
    eng = Engine(zone.dns_domain, zone.datacenter, recursion=recursion, device=local_rank, max_batch=maxB,
                 max_batch_bytes=maxB * 64, ordered=args.ordered)
    zstat = eng.load_snapshot(zone.jsonl)
    log('[bench] zone image %.0f MB built+uploaded in %.1fs' % (zstat['image_bytes'] / 1e6, time.time() - t0))

    K = args.steps
    kp = KernelPath(eng, zone, wl, B, per_q, dev)
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    m = kp.measure(K, args.warmup)
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0
    r0 = kp.result0()
    roofline = kp.roofline(m, r0)
    ring, ring_n, out_cap, oo, ol, st0, tot = kp.ring, kp.ring_n, kp.out_cap, r0['oo'], r0['ol'], r0['st'], r0['tot']
    value, ms_per_step, regions, warm_ms = m['value'], m['ms_per_step'], m['regions'], m['warm_ms']

    # ---- e2e through the C ABI with host buffers -------------------------------------------------
    e2e = None
    e2e_launches = 0
    if not args.no_e2e:
        from binder_b200._lib import lib, check
        L = lib()
        nslots = L.bb_engine_slots(eng._h)
        hb = []
        for r in range(nslots * 2):
            data, off, _ = ring[r % ring_n]
            sizes = dict(pk=data.size, off=(B + 1) * 4, out=out_cap, oo=(B + 1) * 4, ol=B * 2, st=B, ms=B * 4)
            ptr = {k: L.bb_host_alloc(v) for k, v in sizes.items()}
            ctypes.memmove(ptr['pk'], data.ctypes.data, data.size)
            ctypes.memmove(ptr['off'], off.ctypes.data, (B + 1) * 4)
            hb.append(dict(ptr=ptr, nm=ctypes.c_uint32(0), in_bytes=int(off[B]) + (B + 1) * 4))

        def submit(slot, h):
            check(L.bb_resolve_submit(eng._h, slot, h['ptr']['pk'], h['ptr']['off'], B, SEED, 0, h['ptr']['out'],
                                      out_cap, h['ptr']['oo'], h['ptr']['ol'], h['ptr']['st'], h['ptr']['ms'], ctypes.byref(h['nm'])))

        def e2e_run(nst):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            inflight = [False] * nslots
            for k in range(nst):
                slot = k % nslots
                if inflight[slot]:
                    check(L.bb_resolve_wait(eng._h, slot))
                submit(slot, hb[k % len(hb)])
                inflight[slot] = True
            for slot in range(nslots):
                if inflight[slot]:
                    check(L.bb_resolve_wait(eng._h, slot))
            torch.cuda.synchronize()
            return time.perf_counter() - t0
        e2e_run(max(args.warmup, nslots * 2))
        per = e2e_run(nslots * 2) / (nslots * 2)
        ksteps = int(min(max(K, nslots * 4, 0.5 / max(per, 1e-6)), 4000))      # >= 0.5 s of wall clock
        l0 = eng.launch_count()
        dts = [e2e_run(ksteps) for _ in range(3)]
        dt = float(np.median(dts))
        e2e_launches = (eng.launch_count() - l0) // 3
        h = hb[0]
        oo_h = np.ctypeslib.as_array(ctypes.cast(h['ptr']['oo'], ctypes.POINTER(ctypes.c_uint32)), shape=(B + 1,))
        d2h = int(oo_h[B]) + (B + 1) * 4 + B * 2 + B + 16
        e2e = {'value': B * ksteps / dt, 'unit': UNIT, 'h2d_bytes_per_step': h['in_bytes'], 'd2h_bytes_per_step': d2h,
               'steps': ksteps, 'in_flight': nslots, 'timing': 'wall clock bracketed by device synchronize, median of 3 runs',
               'h2d_gbs': h['in_bytes'] * ksteps / dt / 1e9, 'd2h_gbs': d2h * ksteps / dt / 1e9,
               'api': 'bb_resolve_submit/bb_resolve_wait, pinned host buffers (answers written by the kernel straight into them)'}
        out_h = np.ctypeslib.as_array(ctypes.cast(h['ptr']['out'], ctypes.POINTER(ctypes.c_uint8)), shape=(int(oo_h[B]),))
        ol_h = np.ctypeslib.as_array(ctypes.cast(h['ptr']['ol'], ctypes.POINTER(ctypes.c_uint16)), shape=(B,))
        # slot 0's host result must equal the device-resident result of the same batch
        dev_packed = repack(r0['out'], oo, ol)[0]
        assert np.array_equal(repack(out_h, oo_h, ol_h)[0], dev_packed), 'e2e result differs from kernel-path result'
        for hbuf in hb:
            for pv in hbuf['ptr'].values():
                L.bb_host_free(pv)

    # ---- secondary workloads on the same zone / engine (kernel path + roofline; parity when the oracle is there) ----
    secondary = {}
    sec_kp = []
    for w in also:
        zr2, b2, pq2 = DEFAULTS[w]
        k2 = KernelPath(eng, zone, w, b2, pq2, dev, seed0=3000)
        m2 = k2.measure(K, args.warmup)
        r2 = k2.result0()
        rf = k2.roofline(m2, r2)
        secondary[w] = {'workload': synth.WORKLOADS[w][0] + ' [single GPU]', 'batch': b2, 'value': m2['value'], 'ms_per_step': m2['ms_per_step'],
                        'kernel_ms': m2['kern_ms'], 'roofline_frac': rf['frac'], 'in_flight_frac': rf['in_flight_frac'], 'batches_in_flight': k2.depth,
                        'bytes_per_query': rf['bytes_per_query'], 'answered': int((r2['st'] == 0).sum()), 'misses': int(r2['tot'][1])}
        launches += 0
        sec_kp.append((w, k2, r2))

    # ---- CPU baseline + bit-exact check of the timed workload -------------------------------------
    cpu, parity = None, 'not checked (--no-cpu)'
    if loader is not None:
        orc = loader.get()
        log('[cpu] oracle loaded the %d-record zone in %.1fs (background)' % (zone.n_records, loader.secs))
        parity = kp.check_oracle(orc, r0)
        for w, k2, r2 in sec_kp:
            secondary[w]['parity'] = k2.check_oracle(orc, r2)
        cpu = cpu_port(orc, ring[0][0], ring[0][1])
        cpu['oracle_load_s'] = loader.secs
        if loader.same is not None:
            try:
                cpu['same_table'] = cpu_same_table(loader.same, ring[0][0], ring[0][1], int(r0['tot'][0]), int(r0['tot'][1]), per_q, seed=SEED + 0)
                cpu['same_table']['zone_build_s'] = loader.same_secs
            except Exception as ex:
                cpu['same_table'] = {'unavailable': repr(ex)}
        else:
            cpu['same_table'] = {'unavailable': loader.same_err}

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': 1, 'steps': K, 'warmup': args.warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong' if wl == 'config4' else 'weak', 'vs_baseline': None,
            'dtype': 'u8', 'data': 'synthetic',
            'config': {'workload': desc, 'zone_records': zone.n_records, 'batch': B, 'table_mb': zstat['image_bytes'] / 1e6,
                       'l2_policy': 'inputs larger than L2: ring of %d distinct batches (%.0f MB in+out) over a %.0f MB table'
                                    % (ring_n, ring_n * (ring[0][0].size + int(tot[0]) + 11 * B) / 1e6, zstat['image_bytes'] / 1e6),
                       'parallelism': 'single GPU', 'batches_in_flight': kp.depth,
                       'ms_per_step_by_batches_in_flight': {str(k): v for k, v in kp.depth_ms.items()},
                       'timing': 'median of %d regions of %d steps after %.0f ms of device warm-up (regions ms: min %.3f max %.3f)'
                                 % (REGIONS, K, warm_ms, min(regions), max(regions)),
                       'serial_ms_per_step': m['serial_ms'], 'graph_replay_ms_per_step': m['graph_ms'],
                       'output_packing': 'query order (look-back)' if args.ordered else 'arrival (one atomic claim per tile)',
                       'answered': int((st0 == 0).sum()), 'misses': int(tot[1]), 'response_bytes': int(tot[0]),
                       'parity': parity, 'also_measured': secondary},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches + e2e_launches),
            'roofline': roofline, 'cpu_baseline': cpu}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default=None, choices=['config2', 'config3', 'config4', 'config5', 'config5_rd0'],
                    help='default: config3 on one GPU (largest single-GPU configuration), config4 on N>1')
    ap.add_argument('--zone-records', type=int, default=0, help='override the workload\'s zone size (recorded in config)')
    ap.add_argument('--batch', type=int, default=0, help='override the workload\'s batch (global batch for N>1)')
    ap.add_argument('--mode', default='auto', choices=['auto', 'shard', 'replicas', 'nccl'],
                    help='N>1: auto (default) = replicas while the zone image fits one GPU several times over (measured faster at every '
                         'N: DESIGN.md section 8), else shard; shard = hash-sharded zone, route+push over NVLink peer memory; replicas = '
                         'every GPU holds the full zone, no exchange; nccl = sharded zone, routed records exchanged with an NCCL all-to-all')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg and the oracle parity check (profiling runs)')
    ap.add_argument('--no-e2e', action='store_true', help='skip the e2e leg (profiling runs)')
    ap.add_argument('--also', default=None, help='N=1: comma list of further workloads on the same zone measured (kernel path + '
                    'roofline) and reported under config.also_measured; "none" to skip')
    ap.add_argument('--ordered', action='store_true', help='query-order packing (look-back) instead of arrival packing')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    args.steps = max(args.steps, 1)
    claim_stdout()

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.workload is None:
        args.workload = 'config4' if max(world, args.gpus) > 1 else 'config3'

    if args.impl == 'reference':
        if rank == 0:
            run_reference(args)
        return
    if world > 1:
        import bench_multi
        return bench_multi.main(args, rank, world, local_rank)
    run_single(args, local_rank)


if __name__ == '__main__':
    main()
