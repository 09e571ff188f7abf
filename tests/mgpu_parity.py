"""Multi-rank parity driver (run under torchrun, one rank per GPU, or with --cpu under gloo for the host logic).

GPU mode: every rank solves its share of a multi-block case with cross-rank NCCL halo exchange, then solves the WHOLE
case alone on its own GPU; the states of its zones must agree (only the transport differs, so the match is exact up to
the 1e-12 guard).  CPU mode (--cpu): no solver; the per-rank join tables are exercised with a numpy pack / exchange /
unpack over gloo and compared with the single-process table applied to the same fields."""
import os, sys
import numpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def np_apply(table, fields, send=None, recv=None, neq=5):
    """numpy restatement of csrc/fb_joins.cu k_join for the three kinds (fields: list of (neq, nd) arrays per local zone)."""
    nj = int(table[0]); idx = table[2 + 7 * nj:]
    for q in range(nj):
        kind, dz, rz, npts, od, orr, boff = [int(x) for x in table[2 + 7 * q: 9 + 7 * q]]
        if kind == 0: fields[rz][:, idx[orr:orr + npts]] = fields[dz][:, idx[od:od + npts]]
        elif kind == 1 and send is not None: send[boff:boff + neq * npts] = fields[dz][:, idx[od:od + npts]].ravel()
        elif kind == 2 and recv is not None: fields[rz][:, idx[orr:orr + npts]] = recv[boff:boff + neq * npts].reshape(neq, npts)


def main_cpu():
    import torch, torch.distributed as dist
    from fast_b200 import synth, dist as fdist
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    n, nblk = 6, (2, 2, 1) if world == 2 else fdist.grid_for(world)
    nz = nblk[0] * nblk[1] * nblk[2]
    layout = {}
    for bk in range(nblk[2]):
        for bj in range(nblk[1]):
            for bi in range(nblk[0]): layout[(bi, bj, bk)] = len(layout)
    periodic = (True, False, True)
    nd = (n + 4) ** 3
    rng = numpy.random.default_rng(5)
    full = [rng.standard_normal((5, nd)) for _ in range(nz)]
    ref = [f.copy() for f in full]
    np_apply(synth.build_joins(layout, nblk, (n, n, n), periodic=periodic), ref)
    rank_of = fdist.partition(nz, world)
    mine = sorted(z for z, r in rank_of.items() if r == rank)
    table, peers = synth.build_joins(layout, nblk, (n, n, n), periodic=periodic, rank_of=rank_of, rank=rank, neq=5)
    loc = [full[z].copy() for z in mine]
    ns = sum(v[1] for v in peers.values()); nr = sum(v[3] for v in peers.values())
    send = numpy.zeros(ns); recv = numpy.zeros(nr)
    np_apply(table, loc, send=send)                       # local joins + pack
    ts, tr = torch.from_numpy(send), torch.from_numpy(recv)
    ops = []
    for p, (so, sn, ro, rn) in sorted(peers.items()):
        if sn: ops.append(dist.P2POp(dist.isend, ts[so:so + sn], p))
        if rn: ops.append(dist.P2POp(dist.irecv, tr[ro:ro + rn], p))
    for r in dist.batch_isend_irecv(ops): r.wait()
    np_apply(table, loc, recv=recv)                       # local joins again (idempotent) + unpack
    for a, z in zip(loc, mine):
        assert numpy.array_equal(a, ref[z]), "rank %d zone %d: sharded joins differ from the single-process table" % (rank, z)
    dist.barrier()
    if rank == 0: print("MGPU_CPU_OK world=%d zones=%d" % (world, nz))
    dist.destroy_process_group()


def main_gpu(workload):
    import torch, torch.distributed as dist
    from fast_b200 import synth, dist as fdist
    from fast_b200.FastS import PyTree as FastS
    from fast_b200.FastC import PyTree as FastC
    from fast_b200.FastC.PyTree import Internal
    rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    n = 16 if workload != 'c5' else (16, 24, 8)
    # last tile of a direction one cell thick (94 = 3 x 31 + 1, 17 = 2 x 8 + 1, 33 = 32 + 1) -- the second donor layer
    # of the cut face then lives in the penultimate tile, which must belong to part 1 of the overlapped stage
    if workload.startswith('tgv_cut_'): n = (94, 17, 33)
    t, table, peers, desc, cells = fdist.local_case(workload, rank, world, n=n)
    t, metrics, hx = fdist.warmup_rank(t, table, peers, local)
    for it in range(1, 4): FastS._compute(t, metrics, it)
    FastS.sync_tree(t)
    mine = [(z[0], FastC.block_of(z, '').copy()) for z in Internal.getZones(t)]
    calls = hx.calls
    # the whole case on this GPU alone
    if workload == 'c5':
        g = fdist.grid_for(world); nb = (2 * g[0], 2 * g[1], 2 * g[2])
        nglob = tuple(n[d] * nb[d] for d in range(3))
        t2, layout = synth.case_flatplate(n=nglob, nblocks=nb, modulo_verif=1000, only=set(range(nb[0] * nb[1] * nb[2])), return_layout=True)
        j2 = synth.build_joins(layout, nb, n, periodic=(False, False, True))
    elif workload == 'multiblock':
        t2, layout = synth.case_multiblock(nblk=(2, 2, 2), n=n, modulo_verif=1000, only=set(range(8)), return_layout=True)
        j2 = synth.build_joins(layout, (2, 2, 2), (n, n, n))
    else:
        nblk = fdist.grid_for(world)
        if workload.startswith('tgv_cut_'): nblk = tuple(world if d == 'ijk'.index(workload[-1]) else 1 for d in range(3))
        t2, layout = synth.case_tgv_blocks(n=n, nblk=nblk)
        j2 = synth.build_joins(layout, nblk, (n, n, n) if isinstance(n, int) else n, periodic=tuple(b > 1 for b in nblk))
    (t2, tc2, m2) = FastS.warmup(t2, j2)
    for it in range(1, 4): FastS._compute(t2, m2, it)
    FastS.sync_tree(t2)
    ref = {z[0]: FastC.block_of(z, '') for z in Internal.getZones(t2)}
    worst = 0.
    for name, a in mine:
        b = ref[name]
        assert numpy.isfinite(a).all()
        worst = max(worst, float(numpy.abs(a - b).max() / numpy.abs(b).max()))
    w = torch.tensor([worst], device='cuda', dtype=torch.float64); dist.all_reduce(w, op=dist.ReduceOp.MAX)
    if rank == 0: print("MGPU_GPU %s world=%d transport=%s exchanges(host callbacks)=%d max_rel_diff=%.3e %s" % (workload, world, "native NCCL" if hx.native else "torch callback", calls, w.item(), "OK" if w.item() <= 1e-12 else "FAIL"))
    dist.destroy_process_group()
    assert w.item() <= 1e-12


if __name__ == '__main__':
    if '--cpu' in sys.argv: main_cpu()
    else: main_gpu(sys.argv[1] if len(sys.argv) > 1 else 'tgv')
