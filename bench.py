#!/usr/bin/env python
"""bench.py -- headline benchmark of the FastS residual-and-update hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c5|plate|multiblock|tgv256] [--no-secondary]

A "step" is one full time iteration of the hot path (all sub-iterations: ss_iteration Newton sub-iterations incl. the LU
sweeps, or 3 RK stages, plus match-join transfers and BC fill) over one synthetic case.

Default workload = the north-star chain, BASELINE.json configs[4]: the RANS-SA flat plate in blocks of 256x256x64 (AUSM o3,
implicit LU-SGS, 5 Newton sub-iterations, local time step), 8 blocks (a 2x2x2 brick, 33.5 M cells) per GPU -- at N=8
(torchrun, one rank per GPU) the 1024x1024x256 plate of 268 M cells in 64 blocks; weak scaling, ghost cells of the cut faces
exchanged between the ranks every sub-iteration.  At N=1 the line also carries, under config.secondary, the other BASELINE
configurations that fit one GPU -- configs[1] TGV 256^3 (explicit RK3), configs[2] SA plate 512x256x64, configs[3] 8x128^3
Roe lu_match -- each with its own roofline and e2e.
`value` is Mcell.it/s = interior cells x time iterations / second with the state resident in HBM; `e2e` is the same metric
through the reference-facing C-ABI call with HOST buffers (fb_compute_host: H2D of the input time levels and D2H of the new
one inside the timed region).  `--impl reference` times the CPU oracle (the reference's Fortran cannot be built here: no
Fortran compiler, no Cassiopee) with all host threads on the SAME mesh and configuration.
"""
import argparse, ctypes, json, os, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy  # noqa: E402

WORKLOADS = ('c5', 'plate', 'multiblock', 'tgv256')
# cpu_baseline leg of the GPU arm: a bounded sample (about 10-30 s of CPU work) of the same workload
CPU_SAMPLE = {'tgv256': 128, 'plate': (256, 128, 32), 'multiblock': 64, 'c5': (128, 128, 32)}


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


def make_case(workload, n=None):
    """(tree, join table, description) of the single-process case."""
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


def algorithmic_bytes(pi):
    """doubles moved once per irreducible phase, per cell-update (sub-iteration).
    Returns (whole sub-iteration, rhs kernel, one LU sweep) in bytes per cell."""
    from fast_b200 import VSHARE as V
    neq = int(pi[V.NEQ]); typ = int(pi[V.ITYPZONE]); iflow = int(pi[V.IFLOW])
    m = 10 if typ == 0 else 0
    a = (1 if iflow >= 2 else 0) + (1 if iflow == 3 else 0)
    if int(pi[V.ITYPCP]) == 2:
        c = 0 if typ == 2 else 1
        b = 8. * ((11. / 3.) * neq + m + c)
        return b, b, 0.
    sub = 8. * (13 * neq + 1 + a + 3 * m); rhs = 8. * (4 * neq + 1 + a + m)   # rhs: 3 neq read + neq write + coe1 + a + m
    return sub, rhs, (sub - rhs) / 2.


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


# ------------------------------------------------------------------------------------------------------------ CPU arm
def cpu_oracle_rate(workload, n=None, budget_s=15., max_iters=200, threads=None):
    """Mcell.it/s of the CPU oracle (OpenMP: k slabs for the rhs; the zones' LU sweeps spread over the threads, in teams of
    several threads per zone when there are more threads than zones) on the case
    make_case(workload, n): one warm-up iteration (first touch), then whole iterations until `budget_s` is spent (at least 2)."""
    sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    from helpers import oracle_from_tree, V
    t, j, desc = make_case(workload, n=n)
    o = oracle_from_tree(t, joins=j, fast=True)
    del t
    # all host cores this process may use: torchrun exports OMP_NUM_THREADS=1 to its ranks, which would silently turn the
    # reference arm at N > 1 into a one-thread run (the oracle passes the count in a num_threads clause, so it is honoured)
    nthreads = threads or max(o.L.fo_max_threads(), len(os.sched_getaffinity(0)))
    nss = int(o.zones[0].pi[V.NSSITER])
    cells = sum(int(z.pi[V.IJKV]) * int(z.pi[V.IJKV + 1]) * int(z.pi[V.IJKV + 2]) for z in o.zones)
    o.compute(1, 1, nss, 1000, nthreads=nthreads)                     # warm-up / first touch
    # threads left over after one per zone can join the zones' LU sweeps (teams, same bits: fo_invlu_team); one iteration each
    # way decides (a team of spinning threads loses where the process has fewer cores than its affinity mask says)
    team, nz = 0, len(o.zones)
    if int(o.zones[0].pi[V.ITYPCP]) != 2 and nthreads // min(nz, nthreads) >= 2:
        tt = []
        for mode in (0, 1):
            o.L.fo_set_lu_team(mode); t1 = time.perf_counter(); o.compute(2, 1, nss, 1000, nthreads=nthreads); tt.append(time.perf_counter() - t1)
        team = 1 if tt[1] < tt[0] else 0
    o.L.fo_set_lu_team(team)
    nit, t0 = 0, time.perf_counter()
    while nit < max_iters and (nit < 2 or time.perf_counter() - t0 < budget_s):
        o.compute(3 + nit, 1, nss, 1000, nthreads=nthreads); nit += 1
    dt = time.perf_counter() - t0
    o.L.fo_set_lu_team(0)
    nper = nthreads // min(nz, nthreads)
    lu_note = '' if int(o.zones[0].pi[V.ITYPCP]) == 2 else ('; LU sweeps: %d zones, %s' % (nz, 'one per thread' if not team else
              'teams of %d threads per zone ((j slab, i chunk) wavefront, bit-identical to the serial sweep; faster than one zone per thread in a trial iteration)' % nper))
    return cells * nit / dt / 1e6, nthreads, '%s, %d iterations in %.1f s on %d threads%s' % (desc, nit, dt, nthreads, lu_note), nit, dt, desc


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0: return
    # the SAME mesh and configuration as the GPU arm's N=1 workload; the number of timed iterations is bounded so that the arm
    # ends within a few minutes (one c5 iteration is ~20 s of 16-thread CPU work): at most --steps, at least 2
    v, nth, sample, nit, dt, full_desc = cpu_oracle_rate(args.workload, n=args.n, budget_s=args.ref_budget, max_iters=max(2, args.steps))
    out = {'impl': 'reference', 'metric': 'FP64 cell-updates/sec', 'value': v, 'unit': 'Mcell.it/s', 'n_gpus': args.gpus,
           'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / nit * 1e3, 'higher_is_better': True, 'scaling': 'weak',
           'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
           'config': {'workload': full_desc, 'same_mesh_as_gpu_arm': args.n is None, 'steps_timed': nit, 'warmup_done': 1,
                      'note': 'CPU oracle (port of the reference Fortran; the reference itself cannot be built here: no Fortran compiler, no '
                              'Cassiopee), all host threads, the full mesh of the GPU arm; timed iterations bounded by --ref-budget seconds'},
           'cpu_baseline': {'value': v, 'unit': 'Mcell.it/s', 'cores': nth, 'kind': 'port', 'sample': sample},
           'e2e': {'value': v, 'unit': 'Mcell.it/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------------------ GPU arm
def measure(workload, steps, warmup, ctx, do_e2e=True, min_timed_s=0.):
    """One workload on this rank's GPU: warm-up, `steps` timed iterations (device-resident), roofline of the dominant kernel,
    e2e through fb_compute_host.  Returns the result dictionary; the solver is destroyed on the way out."""
    import torch
    from fast_b200 import _lib, VSHARE as V
    from fast_b200.FastS import PyTree as FastS
    from fast_b200.FastC import PyTree as FastC
    from fast_b200.FastC.PyTree import Internal
    rank, world, local, dist = ctx['rank'], ctx['world'], ctx['local'], ctx['dist']
    if world > 1:
        from fast_b200 import dist as fdist
        t, table, peers, desc, cells_total = fdist.local_case(workload, rank, world)
        t, metrics, hx = fdist.warmup_rank(t, table, peers, local)
    else:
        t, joins, desc = make_case(workload)
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
    v0 = ctypes.c_longlong(0); _lib.check(L.fb_counter(b'lu_batch', ctypes.byref(v0)))
    lu_batch0 = int(v0.value)          # process-wide counters: launches of the multi-zone / thread-per-cell LU kernels before this workload
    _lib.check(L.fb_counter(b'lu_cell', ctypes.byref(v0))); lu_cell0 = int(v0.value)

    def fence():
        torch.cuda.synchronize(); _lib.check(L.fb_sync(s))
        if dist is not None: dist.barrier(); torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None: return x
        tt = torch.tensor([x], dtype=torch.float64, device='cuda'); dist.all_reduce(tt, op=dist.ReduceOp.MAX); return float(tt.item())

    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    nit = 1
    fence()
    for w in range(warmup):
        if w == warmup - 1: e0.record(stream)
        FastS._compute(t, metrics, nit); nit += 1
    e1.record(stream); fence()
    if min_timed_s > 0. and warmup > 0:      # secondary workloads: enough iterations for a timed region of min_timed_s
        per = max_over_ranks(e0.elapsed_time(e1)) * 1e-3
        steps = max(steps, int(min_timed_s / max(per, 1e-6)) + 1)
    _lib.check(L.fb_timing_enable(s, 1)); _lib.check(L.fb_timing_reset(s))
    clk = Clocks(); clk.start()
    fence(); e0.record(stream)
    for k in range(steps):
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
    value = cells_total * steps / (ms * 1e-3) / 1e6
    # roofline of the dominant kernel: algorithmic bytes per launch / mean launch duration (CUDA events on the solver's stream)
    b_sub, b_rhs, b_lu = algorithmic_bytes(pi0)
    peak, pk = peaks()
    cz = cells / len(zones)
    cnt = {}
    for name in ('rhs_tma', 'rhs_l1', 'lu_batch', 'lu_cell'):
        v = ctypes.c_longlong(0); _lib.check(L.fb_counter(name.encode(), ctypes.byref(v))); cnt[name] = int(v.value)
    rhs_kern = 'k_rhs3' if cnt['rhs_tma'] >= cnt['rhs_l1'] else 'k_rhs2'
    per_kernel = {rhs_kern: {'avg_launch_ms_per_zone': tm['rhs'] / (steps * nsub * len(zones)), 'bytes_per_cell': b_rhs,
                             'GBs': cells * b_rhs * nsub * steps / (tm['rhs'] * 1e-3) / 1e9 if tm['rhs'] else None}}
    if implicit and tm['lu']:
        lu_name = ('k_lu_cell' if lu_cell0 < cnt['lu_cell'] else 'k_lu_tile') + ('_b' if lu_batch0 < cnt['lu_batch'] else '')
        per_kernel[lu_name] = {
            'avg_sweep_ms_all_zones': tm['lu'] / (2 * steps * nsub), 'bytes_per_cell': b_lu,
            'GBs': cells * b_lu * 2 * nsub * steps / (tm['lu'] * 1e-3) / 1e9}
    for k in per_kernel.values():
        if k['GBs']: k['frac'] = k['GBs'] / peak
    if tm['lu'] > tm['rhs']:
        kern = lu_name
        nl, bpl = 2 * steps * nsub * len(zones), b_lu   # lower + upper launch
        if lu_batch0 < cnt['lu_batch']:      # one lower + one upper launch sweep ALL zones of the rank
            nl, cz = 2 * steps * nsub, cells
        avg_ms = tm['lu'] / nl
    else:
        kern = rhs_kern
        nl, bpl = steps * nsub * len(zones), b_rhs
        avg_ms = tm['rhs'] / nl
    ach = cz * bpl / (avg_ms * 1e-3) / 1e9
    sub_gbs = cells_total * b_sub * nsub * steps / (ms * 1e-3) / 1e9
    roof = {'bound': 'hbm', 'kernel': kern, 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
            'traffic': measured_traffic(workload, kern) if world == 1 else None,
            'peak_source': pk + ' (MEASURED_PEAKS.json hbm_gbs)' if pk == 'measured' else pk, 'bytes_per_cell_per_launch': bpl,
            'avg_launch_ms': avg_ms, 'kernels': per_kernel,
            'whole_subiteration': {'bytes_per_cell_update': b_sub, 'GBs': sub_gbs, 'frac': sub_gbs / (peak * world)},
            'step_share': {k: (v / tm['all'] if tm['all'] else None) for k, v in tm.items()}}
    res = {'workload': desc, 'cells': cells_total, 'sub_iterations_per_step': nsub, 'steps': steps, 'ms_per_step': ms / steps,
           'value': value, 'cell_updates_per_s': cells_total * nsub * steps / (ms * 1e-3), 'roofline': roof, 'gpu_launches': nlaunch,
           'clocks': clk.summary(), 'timed_region_s': ms * 1e-3, 'implicit': implicit}
    if hx is not None:
        res['halo'] = {'transport': hx.describe() + (', after the upper sweep' if implicit else ', overlapped with the interior tiles of the stage'),
                       'bytes_per_exchange_per_rank': hx.bytes_per_exchange,
                       'exchanges_per_step': nsub * (2 if implicit and int(pi0[V.LU_MATCH]) else 1)}

    # end to end through the C-ABI with host buffers (pinned): H2D inputs + D2H result every step
    if do_e2e:
        nz = len(zones)
        hb = [[torch.empty(FastC.block_of(z, suf).size, dtype=torch.float64, pin_memory=True) for z in zones] for suf in ('', '_M1', '_P1')]
        FastS.sync_tree(t)
        for li, suf in enumerate(('', '_M1', '_P1')):
            for zi, z in enumerate(zones): hb[li][zi].numpy()[:] = FastC.block_of(z, suf)
        PtrArr = ctypes.c_void_p * nz
        def ptrs(li): return PtrArr(*[hb[li][zi].data_ptr() for zi in range(nz)])
        fi = ctypes.c_int(1)
        e2e_steps = max(2, min(steps, 5))
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
        # inputs of an iteration: ro, and ro_M1 where the BDF2 time term reads it (not with local time stepping: core3as2.for:37-46)
        bdf2 = implicit and int(pi0[V.DTLOC]) != 1
        h2d = state_bytes * (2 if bdf2 else 1) * world; d2h = state_bytes * world
        res['e2e'] = {'value': cells_total * e2e_steps / (ms2 * 1e-3) / 1e6, 'unit': 'Mcell.it/s', 'h2d_bytes_per_step': h2d,
                      'd2h_bytes_per_step': d2h, 'steps': e2e_steps, 'ms_per_step': ms2 / e2e_steps,
                      'copies': 'pinned host buffers; uploads zone by zone on a copy stream, overlapped with the first kernels of the zones already resident'}
        a = hb[order[0]][0].numpy()
        assert numpy.isfinite(a).all(), "non-finite state after the benchmark"
        del hb
    else:
        res['e2e'] = None
    fence()
    L.fb_solver_destroy(s); FastC.HOOK['solver'] = None
    return res


def run_gpu(args):
    import torch
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    os.environ['FAST_B200_DEVICE'] = str(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    ctx = {'rank': rank, 'world': world, 'local': local, 'dist': dist, 'lu_batch0': 0}
    main = measure(args.workload, args.steps, args.warmup, ctx, do_e2e=not args.no_e2e)
    secondary = []
    if world == 1 and not args.no_secondary:
        for w in WORKLOADS:
            if w == args.workload: continue
            r = measure(w, min(args.steps, 10), max(3, min(args.warmup, 5)), ctx, do_e2e=not args.no_e2e, min_timed_s=1.0)
            if not args.no_cpu:
                cv, nth, sample, _, _, _ = cpu_oracle_rate(w, n=CPU_SAMPLE[w], budget_s=4.)
                r['cpu_baseline'] = {'value': cv, 'unit': 'Mcell.it/s', 'cores': nth, 'kind': 'port', 'sample': sample}
            secondary.append({k: r[k] for k in ('workload', 'cells', 'sub_iterations_per_step', 'steps', 'ms_per_step', 'value', 'roofline',
                                                 'e2e', 'gpu_launches', 'timed_region_s') + (('cpu_baseline',) if 'cpu_baseline' in r else ())})
    cpu = None
    if not args.no_cpu and world == 1:
        cv, nth, sample, _, _, _ = cpu_oracle_rate(args.workload, n=CPU_SAMPLE[args.workload], budget_s=10.)
        cpu = {'value': cv, 'unit': 'Mcell.it/s', 'cores': nth, 'kind': 'port', 'sample': sample}
    if rank == 0:
        cfg = {'workload': main['workload'], 'cells': main['cells'], 'sub_iterations_per_step': main['sub_iterations_per_step'],
               'l2': 'working set >> 126 MB L2 (no flush needed)', 'cell_updates_per_s': main['cell_updates_per_s'],
               'timed_region_s': main['timed_region_s']}
        if 'halo' in main: cfg['halo'] = main['halo']
        if secondary: cfg['secondary'] = secondary
        out = {'metric': 'FP64 cell-updates/sec', 'value': main['value'], 'unit': 'Mcell.it/s', 'n_gpus': world, 'steps': args.steps,
               'warmup': args.warmup, 'ms_per_step': main['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
               'dtype': 'f64', 'data': 'synthetic', 'config': cfg,
               'roofline': main['roofline'], 'cpu_baseline': cpu, 'e2e': main['e2e'], 'gpu_launches': main['gpu_launches'],
               'clocks': main['clocks']}
        print(json.dumps(out))
    if dist is not None: dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1); ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3); ap.add_argument('--impl', default='fast_b200')
    ap.add_argument('--workload', default='c5', choices=WORKLOADS); ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-e2e', action='store_true'); ap.add_argument('--no-secondary', action='store_true')
    ap.add_argument('--ref-budget', type=float, default=100., help='seconds of timed CPU iterations in the reference arm')
    ap.add_argument('--n', default=None, help='testing only: override the mesh size of the reference arm, e.g. 16,16,8 (or 16 for tgv256 / multiblock)')
    args = ap.parse_args()
    if args.n is not None:
        v = tuple(int(x) for x in args.n.split(','))
        args.n = v[0] if len(v) == 1 else v
    if args.impl == 'reference': run_reference(args)
    else: run_gpu(args)


if __name__ == '__main__':
    main()
