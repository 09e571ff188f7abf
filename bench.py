#!/usr/bin/env python
"""bench.py -- headline benchmark of the FastS residual-and-update hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload tgv256|plate|multiblock|c5]

A "step" is one full time iteration of the hot path (all sub-iterations: 3 RK stages, or ss_iteration Newton
sub-iterations incl. LU sweeps, plus match-join transfers and BC fill) over one synthetic case.  Default workload at
N=1 = BASELINE.json configs[1]: laminar Taylor-Green vortex 256^3, SENSOR scheme, explicit RK3; at N>1 (launched by
torchrun, one rank per GPU) the same case with one 256^3 block per GPU joined periodically (weak scaling), ghost cells of
the cut faces exchanged by NCCL point-to-point every RK stage.
`value` is Mcell.it/s = interior cells x time iterations / second with the state resident in HBM; `e2e` is the same
metric through the reference-facing C-ABI call with HOST buffers (fb_compute_host: H2D of the input time levels and
D2H of the new one inside the timed region).  `--impl reference` times the CPU oracle (the reference's Fortran cannot
be built here: no Fortran compiler, no Cassiopee) with all host threads on a bounded sample of the same workload.
"""
import argparse, ctypes, json, os, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy  # noqa: E402


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f: return float(json.load(f)['hbm_gbs']), 'measured'
    return 6650., 'fallback'


def measured_traffic(workload, kernel):
    """DRAM bytes per launch of `kernel` on `workload` from the committed ncu --set full capture (profiles/traffic.json,
    written by tools/ncu_traffic.py from dram__bytes_read.sum + dram__bytes_write.sum); None when no capture exists."""
    p = os.path.join(ROOT, 'profiles', 'traffic.json')
    if not os.path.exists(p): return None
    with open(p) as f: t = json.load(f)
    e = t.get(workload, {}).get(kernel)
    return None if e is None else e['bytes_per_launch']


def make_case(workload, n=None, rank=0, world=1):
    from fast_b200 import synth
    if workload == 'tgv256':
        n = n or 256
        return synth.case_tgv(n=n), None, 'laminar TGV %d^3, SENSOR o3, explicit RK3 (BASELINE configs[1])' % n
    if workload == 'plate':
        n = n or (512, 256, 64)
        t, j = synth.case_flatplate(n=n, modulo_verif=1000)
        return t, j, 'RANS-SA flat plate %dx%dx%d (3dhomo), AUSM o3, implicit LU-SGS 5 sub-iterations, local dt (BASELINE configs[2])' % n
    if workload == 'multiblock':
        n = n or 128
        t, j = synth.case_multiblock(nblk=(2, 2, 2), n=n, modulo_verif=1000)
        return t, j, '8 x %d^3 blocks, match joins, Roe o3, implicit LU-SGS 3 sub-iterations, lu_match (BASELINE configs[3])' % n
    if workload == 'c5':     # one GPU's share of BASELINE configs[4]: a 2x2x2 brick of 256x256x64 blocks
        from fast_b200 import dist as fdist
        t, table, peers, desc, _ = fdist.local_case('c5', 0, 1, n=n)
        return t, table, desc
    raise ValueError(workload)


def _desc_only(workload):
    d = {'tgv256': 'laminar TGV 256^3, SENSOR o3, explicit RK3 (BASELINE configs[1])',
         'plate': 'RANS-SA flat plate 512x256x64 (3dhomo), AUSM o3, implicit LU-SGS 5 sub-iterations, local dt (BASELINE configs[2])',
         'multiblock': '8 x 128^3 blocks, match joins, Roe o3, implicit LU-SGS 3 sub-iterations, lu_match (BASELINE configs[3])',
         'c5': 'RANS-SA flat plate in blocks of 256x256x64 (3dhomo), 8 blocks per GPU, implicit LU-SGS 5 sub-iterations (BASELINE configs[4])'}
    return None, None, d[workload]
make_case.__wrapped_desc__ = _desc_only


def algorithmic_bytes(pi):
    """doubles moved once per irreducible phase, per cell-update (sub-iteration)."""
    from fast_b200 import VSHARE as V
    neq = int(pi[V.NEQ]); typ = int(pi[V.ITYPZONE]); iflow = int(pi[V.IFLOW])
    m = 10 if typ == 0 else 0
    a = (1 if iflow >= 2 else 0) + (1 if iflow == 3 else 0)
    if int(pi[V.ITYPCP]) == 2:
        c = 0 if typ == 2 else 1
        return 8. * ((11. / 3.) * neq + m + 0 * a + c), 8. * ((11. / 3.) * neq + m + c)   # (whole sub-iteration, rhs kernel)
    return 8. * (13 * neq + 1 + a + 3 * m), 8. * (4 * neq + 1 + a + m)   # rhs: 3 neq read + neq write + coe1 + a + m


class Clocks(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True); self.stop = False; self.samples = []; self.reasons = set(); self.maxmhz = None
    def run(self):
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        while not self.stop:
            try:
                o = subprocess.run(['nvidia-smi', '--query-gpu=' + q, '--format=csv,noheader,nounits', '-i', os.environ.get('LOCAL_RANK', '0')],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(',')
                self.samples.append(float(o[0])); self.maxmhz = float(o[1])
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), o[2:6]):
                    if 'Active' in v and 'Not' not in v: self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.2)
    def summary(self):
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.maxmhz, 'reasons': sorted(self.reasons)}


def cpu_oracle_rate(workload, seconds=15., threads=None):
    """Mcell.it/s of the CPU oracle (OpenMP: k slabs for the rhs, one zone per thread for the LU) on a bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    from helpers import oracle_from_tree, V
    sample_n = {'tgv256': 96, 'plate': (128, 64, 16), 'multiblock': 32, 'c5': (32, 32, 8)}[workload]
    t, j, desc = make_case(workload, n=sample_n)
    o = oracle_from_tree(t, joins=j, fast=True)
    # all host cores this process may use: torchrun exports OMP_NUM_THREADS=1 to its ranks, which would silently turn the
    # reference arm at N > 1 into a one-thread run (the oracle passes the count in a num_threads clause, so it is honoured)
    nthreads = threads or max(o.L.fo_max_threads(), len(os.sched_getaffinity(0)))
    nss = int(o.zones[0].pi[V.NSSITER])
    cells = sum(int(z.pi[V.IJKV]) * int(z.pi[V.IJKV + 1]) * int(z.pi[V.IJKV + 2]) for z in o.zones)
    o.compute(1, 1, nss, 1000, nthreads=nthreads)                     # warm-up / first touch
    t0 = time.perf_counter(); o.compute(2, 1, nss, 1000, nthreads=nthreads); dt1 = time.perf_counter() - t0
    nit = max(1, min(200, int(seconds / max(dt1, 1e-6))))
    t0 = time.perf_counter()
    for it in range(nit): o.compute(3 + it, 1, nss, 1000, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return cells * nit / dt / 1e6, nthreads, '%s for %d iterations (%.1f s)' % (desc, nit, dt), nss


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0: return
    rates = []
    for s in range(args.warmup + args.steps):
        r, nth, sample, nss = cpu_oracle_rate(args.workload, seconds=max(2., 60. / (args.warmup + args.steps)))
        if s >= args.warmup: rates.append(r)
    v = float(numpy.mean(rates))
    _, _, full_desc = make_case.__wrapped_desc__(args.workload)
    out = {'impl': 'reference', 'metric': 'FP64 cell-updates/sec', 'value': v, 'unit': 'Mcell.it/s', 'n_gpus': args.gpus,
           'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': None, 'higher_is_better': True, 'scaling': 'weak',
           'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
           'config': {'workload': full_desc, 'sample': sample,
                      'note': 'CPU oracle (port of the reference Fortran; the reference itself cannot be built here: no Fortran compiler, no '
                              'Cassiopee), all host threads, bounded sample of the workload (same scheme, smaller mesh)'},
           'cpu_baseline': {'value': v, 'unit': 'Mcell.it/s', 'cores': nth, 'kind': 'port', 'sample': sample},
           'e2e': {'value': v, 'unit': 'Mcell.it/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out))


def run_gpu(args):
    import torch
    from fast_b200 import _lib, VSHARE as V
    from fast_b200.FastS import PyTree as FastS
    from fast_b200.FastC import PyTree as FastC
    from fast_b200.FastC.PyTree import Internal
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    os.environ['FAST_B200_DEVICE'] = str(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        from fast_b200 import dist as fdist
        t, table, peers, desc, cells_total = fdist.local_case(args.workload, rank, world)
        t, metrics, hx = fdist.warmup_rank(t, table, peers, local)
    else:
        t, joins, desc = make_case(args.workload)
        (t, tc, metrics) = FastS.warmup(t, joins)
        hx, cells_total = None, None
    zones = Internal.getZones(t)
    L = _lib.lib(); s = FastC.HOOK['solver']
    pi0 = Internal.getNodeFromName1(Internal.getNodeFromName1(zones[0], '.Solver#ownData'), 'Parameter_int')[1]
    nss = int(pi0[V.NSSITER]); implicit = int(pi0[V.ITYPCP]) <= 1
    nsub = nss - 1 if implicit else nss
    cells = 0
    for z in zones:
        pi = Internal.getNodeFromName1(Internal.getNodeFromName1(z, '.Solver#ownData'), 'Parameter_int')[1]
        cells += int(pi[V.IJKV]) * int(pi[V.IJKV + 1]) * int(pi[V.IJKV + 2])
    if cells_total is None: cells_total = cells
    stream = torch.cuda.ExternalStream(L.fb_stream(s))

    def fence():
        torch.cuda.synchronize(); _lib.check(L.fb_sync(s))
        if dist is not None: dist.barrier(); torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None: return x
        tt = torch.tensor([x], dtype=torch.float64, device='cuda'); dist.all_reduce(tt, op=dist.ReduceOp.MAX); return float(tt.item())

    nit = 1
    for w in range(args.warmup):
        FastS._compute(t, metrics, nit); nit += 1
    fence()
    _lib.check(L.fb_timing_enable(s, 1)); _lib.check(L.fb_timing_reset(s))
    clk = Clocks(); clk.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    fence(); e0.record(stream)
    for k in range(args.steps):
        FastS._compute(t, metrics, nit); nit += 1
    e1.record(stream); fence()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clk.stop = True; clk.join(timeout=2)
    tm = {}
    launches = ctypes.c_longlong(0)
    for name in ('rhs', 'lu', 'bc', 'transfer', 'coe', 'all'):
        v = ctypes.c_double(0); _lib.check(L.fb_timing_get(s, name.encode(), ctypes.byref(v), ctypes.byref(launches)))
        tm[name] = v.value
    _lib.check(L.fb_timing_enable(s, 0))
    nlaunch = int(launches.value)
    if dist is not None:
        tt = torch.tensor([nlaunch], dtype=torch.int64, device='cuda'); dist.all_reduce(tt); nlaunch = int(tt.item())
    value = cells_total * args.steps / (ms * 1e-3) / 1e6
    # roofline of the dominant kernel: algorithmic bytes per launch / mean launch duration (CUDA events on the solver's stream)
    b_sub, b_rhs = algorithmic_bytes(pi0)
    peak, pk = peaks()
    cz = cells / len(zones)
    cnt = {}
    for name in ('rhs_tma', 'rhs_l1', 'lu_batch'):
        v = ctypes.c_longlong(0); _lib.check(L.fb_counter(name.encode(), ctypes.byref(v))); cnt[name] = int(v.value)
    if tm['lu'] > tm['rhs']:
        kern = 'k_lu_tile'
        nl, bpl = 2 * args.steps * nsub * len(zones), (b_sub - b_rhs) / 2.   # lower + upper launch
        if cnt['lu_batch']:      # one lower + one upper launch sweep ALL zones of the rank (k_lu_tile_b)
            kern, nl, cz = 'k_lu_tile_b', 2 * args.steps * nsub, cells
        avg_ms = tm['lu'] / nl
    else:
        kern = 'k_rhs3' if cnt['rhs_tma'] >= cnt['rhs_l1'] else 'k_rhs2'
        nl, bpl = args.steps * nsub * len(zones), b_rhs
        avg_ms = tm['rhs'] / nl
    ach = cz * bpl / (avg_ms * 1e-3) / 1e9
    roof = {'bound': 'hbm', 'kernel': kern, 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
            'traffic': measured_traffic(args.workload, kern) if world == 1 else None,
            'peak_source': pk + ' (MEASURED_PEAKS.json hbm_gbs)' if pk == 'measured' else pk, 'bytes_per_cell_per_launch': bpl,
            'avg_launch_ms': avg_ms,
            'whole_subiteration': {'bytes_per_cell_update': b_sub, 'GBs': cells_total * b_sub * nsub * args.steps / (ms * 1e-3) / 1e9,
                                   'frac': cells_total * b_sub * nsub * args.steps / (ms * 1e-3) / 1e9 / (peak * world)},
            'step_share': {k: (v / tm['all'] if tm['all'] else None) for k, v in tm.items()}}

    # end to end through the C-ABI with host buffers (pinned): H2D inputs + D2H result every step
    e2e = None
    nz = len(zones)
    if args.no_e2e:      # auxiliary multi-GPU lines of the large cases (pinned copies of three time levels per rank)
        return finish(args, rank, world, dist, desc, cells_total, nsub, ms, value, roof, None, nlaunch, clk, hx, implicit, pi0, V)
    hb = [[torch.empty(FastC.block_of(z, suf).size, dtype=torch.float64, pin_memory=True) for z in zones] for suf in ('', '_M1', '_P1')]
    FastS.sync_tree(t)
    for li, suf in enumerate(('', '_M1', '_P1')):
        for zi, z in enumerate(zones): hb[li][zi].numpy()[:] = FastC.block_of(z, suf)
    PtrArr = ctypes.c_void_p * nz
    def ptrs(li): return PtrArr(*[hb[li][zi].data_ptr() for zi in range(nz)])
    fi = ctypes.c_int(1)
    e2e_steps = max(2, min(args.steps, 5))
    order = [0, 1, 2]
    def host_step(n):
        _lib.check(L.fb_compute_host(s, n, 1, ctypes.byref(fi), ptrs(order[0]), ptrs(order[1]), ptrs(order[2])))
        order[:] = [order[2], order[0], order[1]]      # switchPointers__ on the host arrays
    host_step(nit); nit += 1
    fence(); e0.record(stream)
    for k in range(e2e_steps):
        host_step(nit); nit += 1
    e1.record(stream); fence()
    ms2 = max_over_ranks(e0.elapsed_time(e1))
    state_bytes = sum(x.numel() * 8 for x in hb[0])
    h2d = state_bytes * (2 if implicit else 1) * world; d2h = state_bytes * world
    e2e = {'value': cells_total * e2e_steps / (ms2 * 1e-3) / 1e6, 'unit': 'Mcell.it/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
           'steps': e2e_steps, 'ms_per_step': ms2 / e2e_steps}
    a = hb[order[0]][0].numpy()
    assert numpy.isfinite(a).all(), "non-finite state after the benchmark"

    return finish(args, rank, world, dist, desc, cells_total, nsub, ms, value, roof, e2e, nlaunch, clk, hx, implicit, pi0, V)


def finish(args, rank, world, dist, desc, cells_total, nsub, ms, value, roof, e2e, nlaunch, clk, hx, implicit, pi0, V):
    cpu = None
    if not args.no_cpu and world == 1:
        r, nth, sample, _ = cpu_oracle_rate(args.workload, seconds=12.)
        cpu = {'value': r, 'unit': 'Mcell.it/s', 'cores': nth, 'kind': 'port', 'sample': sample}
    if rank != 0:
        if dist is not None: dist.destroy_process_group()
        return
    cfg = {'workload': desc, 'cells': cells_total, 'sub_iterations_per_step': nsub, 'l2': 'working set >> 126 MB L2 (no flush needed)',
           'cell_updates_per_s': cells_total * nsub * args.steps / (ms * 1e-3)}
    if hx is not None:
        cfg['halo'] = {'transport': 'NCCL isend/irecv per neighbour on the exchange stream' + (', after the upper sweep' if implicit else ', overlapped with the interior tiles of the stage'),
                       'bytes_per_exchange_per_rank': hx.bytes_per_exchange,
                       'exchanges_per_step': nsub * (2 if implicit and int(pi0[V.LU_MATCH]) else 1)}
    out = {'metric': 'FP64 cell-updates/sec', 'value': value, 'unit': 'Mcell.it/s', 'n_gpus': world, 'steps': args.steps,
           'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
           'dtype': 'f64', 'data': 'synthetic', 'config': cfg,
           'roofline': roof, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': nlaunch, 'clocks': clk.summary()}
    print(json.dumps(out))
    if dist is not None: dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1); ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3); ap.add_argument('--impl', default='fast_b200')
    ap.add_argument('--workload', default=None); ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    args = ap.parse_args()
    if args.workload is None: args.workload = 'tgv256'
    if args.impl == 'reference': run_reference(args)
    else: run_gpu(args)


if __name__ == '__main__':
    main()
