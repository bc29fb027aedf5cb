#!/usr/bin/env python
"""bench.py — DNS queries/sec of the batched resolve path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path (parse -> zone lookup -> answer bytes) over one batch.
N=1 workload = BASELINE.json configs[1]: 1M-record zone, 65,536 A-record lookups per batch.

  value      kernel path, batches resident in HBM, K steps with 4 independent batches in flight
             (one stream each), timed between two CUDA events on the main stream that the
             streams fork from and join into; the steps cycle over a ring of distinct batches
             whose total footprint (inputs + outputs) exceeds L2 — no L2 flush needed.
  e2e        the same metric through bb_resolve_submit/_wait (the C ABI a host calls) with
             pinned HOST buffers: H2D of packets+offsets and D2H of answers inside the timed
             region, 4 batches in flight.
  roofline   algorithmic HBM bytes per launch / measured kernel time vs MEASURED_PEAKS.json.
  cpu_baseline  the CPU oracle (a C++ port of lib/server.js + lib/zk.js — the Node.js
             reference cannot run in this image) on the box's host cores, bounded sample.

--impl reference times that CPU port alone (it is the only reference arm that exists here).
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
WORKLOAD = 'config2: 1M-record zone, 65536-query A-record batches (100% hit, RD=1, no OPT)'
ZONE_RECORDS = 1000000
BATCH = 65536
RING = 24            # distinct batches cycled through: 24 x (3.0 MB in + 4.7 MB out) = 185 MB > 126 MB L2


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


def algorithmic_bytes(off, out_off, key_len_plus_payload):
    """B(q) = len(query)+4 + probe(q) + len(response)+8, summed over the batch."""
    n = len(off) - 1
    read = int(off[n]) + 4 * n + int(key_len_plus_payload)
    write = int(out_off[n]) + 8 * n
    return read, write


def cpu_port(zone, data, off, budget_s=12.0, recursion=False):
    """Oracle (C++ port of the reference path) on the host cores, bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    from oracle_lib import Oracle
    t0 = time.time()
    orc = Oracle(zone.dns_domain, zone.datacenter, recursion, snapshot=zone.jsonl)
    log('[cpu] oracle loaded %d-record zone in %.1fs' % (zone.n_records, time.time() - t0))
    cores = os.cpu_count() or 1
    n = len(off) - 1
    t1 = orc.timed_resolve(data, off, nthreads=1, repeat=2)
    reps = max(2, min(40, int(budget_s / max(min(t1) / max(cores * 0.5, 1), 1e-3))))
    tn = orc.timed_resolve(data, off, nthreads=cores, repeat=reps)
    return orc, {'value': n / min(tn), 'unit': UNIT, 'cores': cores, 'kind': 'port',
                 'sample': '%d x one %d-query batch of the same workload, best call; all %d host threads '
                           '(single thread: %.0f queries/s)' % (reps, n, cores, n / min(t1)),
                 'single_thread_value': n / min(t1)}


def run_reference(args, zone):
    """--impl reference: the CPU port of the reference path (Node.js cannot run here)."""
    from binder_b200 import synth
    data, off = synth.batch_host_a_fast(zone, BATCH, seed=1000)
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    from oracle_lib import Oracle
    orc = Oracle(zone.dns_domain, zone.datacenter, False, snapshot=zone.jsonl)
    cores = os.cpu_count() or 1
    orc.timed_resolve(data, off, nthreads=cores, repeat=max(args.warmup, 1))
    ts = orc.timed_resolve(data, off, nthreads=cores, repeat=args.steps)
    total = sum(ts)
    v = BATCH * args.steps / total
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'zone_records': zone.n_records, 'batch': BATCH,
                       'note': 'CPU restatement of lib/server.js + lib/zk.js + mname codec (oracle/oracle.cpp); '
                               'the Node.js reference is not runnable in this image'},
            'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': '%d steps x one %d-query batch, all %d host threads' % (args.steps, BATCH, cores)},
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20000)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--zone-records', type=int, default=ZONE_RECORDS)
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg (profiling runs)')
    ap.add_argument('--no-e2e', action='store_true', help='skip the e2e leg (profiling runs)')
    ap.add_argument('--workload', default='config2', choices=['config2', 'config3', 'config5'],
                    help='config2 (default, the headline), config3 (services: 50%% SRV / 50%% service-A), config5 (90%% misses, recursion split)')
    ap.add_argument('--ordered', action='store_true', help='query-order packing (look-back) instead of arrival packing')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    claim_stdout()

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    from binder_b200 import synth
    if args.impl == 'reference':
        if rank == 0:
            run_reference(args, synth.gen_zone(args.zone_records))
        return

    if world > 1:
        import bench_multi
        return bench_multi.main(args, rank, world, local_rank)

    import torch
    from binder_b200.engine import Engine, repack
    from binder_b200 import build as bbuild
    bbuild.build()
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device — binder_b200 has no CPU path to benchmark')
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)

    t0 = time.time()
    wl = args.workload
    zone = synth.gen_zone(args.zone_records, service_frac=0.15 if wl == 'config3' else 0.0)
    eng = Engine(zone.dns_domain, zone.datacenter, recursion=(wl == 'config5'), device=local_rank, max_batch=args.batch,
                 max_batch_bytes=args.batch * 64, ordered=args.ordered)
    zstat = eng.load_snapshot(zone.jsonl)
    log('[bench] zone: %d records, table %.0f MB, built+uploaded in %.1fs' % (zone.n_records, zstat['image_bytes'] / 1e6, time.time() - t0))

    B = args.batch
    if wl == 'config2':
        ring = [synth.batch_host_a_fast(zone, B, seed=1000 + r) for r in range(RING)]
        out_cap = B * 96
    elif wl == 'config5':
        ring = [synth.batch_host_a_fast(zone, B, seed=1000 + r, miss_frac=0.9) for r in range(RING)]
        out_cap = B * 96
    else:
        ring = [synth.pack_batch(synth.batch_service(zone, B, seed=1000 + r)) for r in range(8)] * 3
        out_cap = B * 400
    d = []
    for data, off in ring:
        d.append(dict(
            pk=torch.from_numpy(data).to(dev), off=torch.from_numpy(off.view(np.int32)).to(dev),
            out=torch.empty(out_cap, dtype=torch.uint8, device=dev), oo=torch.empty(B + 1, dtype=torch.int32, device=dev),
            st=torch.empty(B, dtype=torch.uint8, device=dev), ms=torch.empty(B, dtype=torch.int32, device=dev),
            ol=torch.empty(B, dtype=torch.int16, device=dev),
            tot=torch.zeros(4, dtype=torch.int32, device=dev)))
    stream = torch.cuda.current_stream()

    def step(k):
        b = d[k % RING]
        eng.resolve_device(b['pk'].data_ptr(), b['off'].data_ptr(), B, 0xB1DDE5, 0, b['out'].data_ptr(), out_cap,
                           b['oo'].data_ptr(), b['ol'].data_ptr(), b['st'].data_ptr(), b['ms'].data_ptr(), b['tot'].data_ptr(),
                           stream.cuda_stream)

    # ---- kernel path, device-resident -----------------------------------------------------------
    # IN_FLIGHT batches at a time, one stream each (independent batches, as a server runs them and
    # as the e2e path below does); timed on the device: the streams fork from / join into the main
    # stream between two CUDA events.
    IN_FLIGHT = 4
    side_streams = [torch.cuda.Stream(device=dev) for _ in range(IN_FLIGHT)]

    def step_on(k, st):
        b = d[k % RING]
        eng.resolve_device(b['pk'].data_ptr(), b['off'].data_ptr(), B, 0xB1DDE5, 0, b['out'].data_ptr(), out_cap,
                           b['oo'].data_ptr(), b['ol'].data_ptr(), b['st'].data_ptr(), b['ms'].data_ptr(), b['tot'].data_ptr(),
                           st.cuda_stream)

    def timed_concurrent(k0, nsteps):
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(stream)
        for st in side_streams:
            st.wait_event(ea)
        for k in range(nsteps):
            step_on(k0 + k, side_streams[k % IN_FLIGHT])
        for st in side_streams:
            ev = torch.cuda.Event()
            ev.record(st)
            stream.wait_event(ev)
        eb.record(stream)
        torch.cuda.synchronize()
        return ea.elapsed_time(eb)

    for k in range(args.warmup):
        step(k)
    timed_concurrent(0, max(args.warmup, IN_FLIGHT))
    torch.cuda.synchronize()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_conc = timed_concurrent(args.warmup, args.steps)
    # the same K steps strictly one after another on one stream: per-launch duration for the roofline
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for k in range(args.steps):
        step(args.warmup + k)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # a second timing of the same loop replayed as one CUDA graph: removes host launch gaps, so it
    # is the kernel's own average duration (what the roofline uses)
    graph_ms = None
    try:
        gsteps = min(args.steps, 500)                # launches per graph; replayed to cover args.steps
        reps = max(1, args.steps // gsteps)
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        step_on(0, side)                 # the engine allocates this stream's scratch outside the capture
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=side):
            cs = torch.cuda.current_stream().cuda_stream
            for k in range(gsteps):
                b = d[(args.warmup + k) % RING]
                eng.resolve_device(b['pk'].data_ptr(), b['off'].data_ptr(), B, 0xB1DDE5, 0, b['out'].data_ptr(), out_cap,
                                   b['oo'].data_ptr(), b['ol'].data_ptr(), b['st'].data_ptr(), b['ms'].data_ptr(), b['tot'].data_ptr(), cs)
        g.replay()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(reps):
            g.replay()
        g1.record()
        torch.cuda.synchronize()
        graph_ms = g0.elapsed_time(g1) * args.steps / (gsteps * reps)
    except Exception as ex:      # capture is an optimisation of the measurement, not a requirement
        log('[bench] CUDA-graph replay unavailable: %r' % (ex,))
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0
    ms_per_step = ms_conc / args.steps
    value = B / (ms_per_step * 1e-3)
    serial_ms = ms / args.steps
    kern_ms = min(serial_ms, graph_ms / args.steps) if graph_ms else serial_ms

    # correctness of what was just timed + algorithmic bytes of one launch
    b = d[(args.warmup + args.steps - 1) % RING]
    oo = b['oo'].cpu().numpy().view(np.uint32)
    tot = b['tot'].cpu().numpy()
    assert tot[0] == oo[B], 'totals disagree with the offset array'
    if wl == 'config2':
        assert (b['st'].cpu().numpy() == 0).all(), 'timed batch was not fully answered'
    key_payload = B * (30 + 1 + 8)            # 30-char key + length byte + (addr, ttl) per hit (config 2; other workloads: same convention, approximate)
    rd_b, wr_b = algorithmic_bytes(ring[0][1], oo, key_payload)
    peak, peak_src = measured_peaks()
    achieved = (rd_b + wr_b) / (kern_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get('dram_bytes_per_launch')
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'peak_source': peak_src, 'kernel': 'bbk::resolve_kernel',
                'kernel_ms': kern_ms, 'kernel_timing': 'one launch at a time: ' + ('CUDA-graph replay of the K launches' if graph_ms and graph_ms / args.steps <= serial_ms else 'stream loop'),
                'algorithmic_bytes_per_launch': rd_b + wr_b, 'read_bytes': rd_b, 'write_bytes': wr_b,
                'read_only_frac': rd_b / (kern_ms * 1e-3) / 1e9 / peak}

    # ---- e2e through the C ABI with host buffers -------------------------------------------------
    e2e = None
    e2e_launches = 0
    if not args.no_e2e:
        import ctypes
        from binder_b200._lib import lib, check
        L = lib()
        nslots = L.bb_engine_slots(eng._h)
        hb = []
        for r in range(nslots * 2):
            data, off = ring[r % RING]
            sizes = dict(pk=data.size, off=(B + 1) * 4, out=out_cap, oo=(B + 1) * 4, ol=B * 2, st=B, ms=B * 4)
            ptr = {k: L.bb_host_alloc(v) for k, v in sizes.items()}
            ctypes.memmove(ptr['pk'], data.ctypes.data, data.size)
            ctypes.memmove(ptr['off'], off.ctypes.data, (B + 1) * 4)
            hb.append(dict(ptr=ptr, nm=ctypes.c_uint32(0), in_bytes=int(off[B]) + (B + 1) * 4))

        def submit(slot, h):
            check(L.bb_resolve_submit(eng._h, slot, h['ptr']['pk'], h['ptr']['off'], B, 0xB1DDE5, 0, h['ptr']['out'],
                                      out_cap, h['ptr']['oo'], h['ptr']['ol'], h['ptr']['st'], h['ptr']['ms'], ctypes.byref(h['nm'])))
        ksteps = min(max(args.steps, nslots * 4), 4000)
        for phase in ('warm', 'timed'):
            nst = max(args.warmup, nslots) if phase == 'warm' else ksteps
            torch.cuda.synchronize()
            l0 = eng.launch_count()
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
            dt = time.perf_counter() - t0
            e2e_launches = eng.launch_count() - l0
        h = hb[0]
        oo_h = np.ctypeslib.as_array(ctypes.cast(h['ptr']['oo'], ctypes.POINTER(ctypes.c_uint32)), shape=(B + 1,))
        d2h = int(oo_h[B]) + (B + 1) * 4 + B * 2 + B + 16
        e2e = {'value': B * ksteps / dt, 'unit': UNIT, 'h2d_bytes_per_step': h['in_bytes'], 'd2h_bytes_per_step': d2h,
               'steps': ksteps, 'in_flight': nslots, 'timing': 'wall clock bracketed by device synchronize',
               'api': 'bb_resolve_submit/bb_resolve_wait, pinned host buffers'}
        out_h = np.ctypeslib.as_array(ctypes.cast(h['ptr']['out'], ctypes.POINTER(ctypes.c_uint8)), shape=(int(oo_h[B]),))
        ol_h = np.ctypeslib.as_array(ctypes.cast(h['ptr']['ol'], ctypes.POINTER(ctypes.c_uint16)), shape=(B,))
        # slot 0's host result must equal the device-resident result of the same batch
        step(0)
        torch.cuda.synchronize()
        dev_packed = repack(d[0]['out'].cpu().numpy(), d[0]['oo'].cpu().numpy().view(np.uint32),
                            d[0]['ol'].cpu().numpy().view(np.uint16))[0]
        assert np.array_equal(repack(out_h, oo_h, ol_h)[0], dev_packed), 'e2e result differs from kernel-path result'

    # ---- CPU baseline + bit-exact spot check of the timed workload --------------------------------
    cpu = None
    if not args.no_cpu:
        orc, cpu = cpu_port(zone, ring[0][0], ring[0][1], recursion=(wl == 'config5'))
        o = orc.resolve_batch(ring[0][0], ring[0][1], seed=0xB1DDE5)
        step(0)
        torch.cuda.synchronize()
        got, goff = repack(d[0]['out'].cpu().numpy(), d[0]['oo'].cpu().numpy().view(np.uint32),
                           d[0]['ol'].cpu().numpy().view(np.uint16))
        assert np.array_equal(got, o[0]) and np.array_equal(goff, o[1]), 'GPU answers differ from the CPU oracle'

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': 1, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'u8', 'data': 'synthetic',
            'config': {'workload': WORKLOAD if wl == 'config2' else wl + ' (SURVEY.md section 8d)', 'zone_records': zone.n_records, 'batch': B, 'table_mb': zstat['image_bytes'] / 1e6,
                       'l2_policy': 'inputs larger than L2: ring of %d distinct batches (%.0f MB in+out) over a %.0f MB table'
                                    % (RING, RING * (ring[0][0].size + out_cap * 2 / 3 + 8 * B) / 1e6, zstat['image_bytes'] / 1e6),
                       'parallelism': 'single GPU', 'batches_in_flight': IN_FLIGHT, 'serial_ms_per_step': serial_ms, 'output_packing': 'query order (look-back)' if args.ordered else 'arrival (one atomic claim per 128-query tile)', 'graph_replay_ms_per_step': graph_ms / args.steps if graph_ms else None},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches + e2e_launches),
            'roofline': roofline, 'cpu_baseline': cpu}
    emit(line)


if __name__ == '__main__':
    main()
