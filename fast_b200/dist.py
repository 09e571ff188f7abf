"""One process per GPU: block partition, cross-rank match-join exchange, multi-GPU benchmark driver.

Reference counterpart: FastS.Mpi (S/Mpi.py:33, same surface with MPI graphs) + the point-to-point exchange of
K_FASTC::setInterpTransfersInter / getTransfersInter (C/Com/setInterpTransfersFast.cpp:962-1790): per sub-iteration one
message per neighbour rank holding the donor values of all joins towards it.  Here the path shards by block (the LU
sweep never crosses a zone, S/Compute/LU/invlu_l.for:79-96), there is no collective on the data path: the library
packs the donor cells of the cross-rank joins on its stream (k_join kind 1), calls back into `HaloExchange`, which
moves the send buffer to the peers' receive buffers with one batched NCCL isend/irecv per peer over NVLink, and
unpacks (kind 2).  Everything stays ordered on the solver's CUDA stream; the host never waits inside an iteration.
"""
import ctypes
import os
import numpy

from . import _lib


def partition(nzones, world):
    """Whole blocks -> ranks, contiguous in zone order (zones are generated k-major, so a rank gets a slab / brick),
    balanced by count (all synthetic cases have equal blocks).  Returns {zone id: rank}."""
    if nzones < world: raise ValueError("fast_b200.dist: %d blocks cannot feed %d ranks" % (nzones, world))
    base, rem = divmod(nzones, world)
    out, z = {}, 0
    for r in range(world):
        for _ in range(base + (1 if r < rem else 0)):
            out[z] = r; z += 1
    return out


def grid_for(world):
    """Block grid of the weak-scaling family: one block per rank, as cubic as possible."""
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(world) or (world, 1, 1)


class _DevArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {'shape': (int(n),), 'typestr': '<f8', 'data': (int(ptr), False), 'version': 2}


class HaloExchange:
    """Cross-rank exchange of the packed donor values.  Default: the library's own NCCL transport (fb_nccl_init + fb_set_peers:
    ncclSend / ncclRecv on the exchange stream, nothing of the host on the data path) -- this class only carries the 128-byte
    NCCL id from rank 0 to the other ranks over the existing torch.distributed group.  FAST_B200_EXCHANGE=torch keeps the
    host-callback transport of round 1 (fb_set_exchange_cb -> torch.distributed batch_isend_irecv), which is also the hook for
    a foreign transport such as the reference's MPI layer."""

    def __init__(self, solver, peers, device, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        L = _lib.lib()
        ns, nr = ctypes.c_longlong(0), ctypes.c_longlong(0)
        _lib.check(L.fb_halo_sizes(solver, ctypes.byref(ns), ctypes.byref(nr)))
        self.peers = dict(peers)
        self.bytes_per_exchange = 8 * (ns.value + nr.value)
        self.calls = 0
        self.native = os.environ.get('FAST_B200_EXCHANGE', 'nccl') != 'torch'
        dev = torch.device('cuda', device)
        if self.native:
            rank, world = dist.get_rank(group), dist.get_world_size(group)
            idbuf = (ctypes.c_char * 128)()
            if rank == 0: _lib.check(L.fb_nccl_unique_id(idbuf))
            t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).clone().to(dev)
            dist.broadcast(t, src=0, group=group)
            raw = bytes(t.cpu().numpy().tobytes())
            _lib.check(L.fb_nccl_init(solver, rank, world, ctypes.create_string_buffer(raw, 128)))
            order = sorted(self.peers.items())
            n = len(order)
            IA = ctypes.c_int * max(n, 1); LA = ctypes.c_longlong * max(n, 1)
            _lib.check(L.fb_set_peers(solver, n, IA(*[p for p, _ in order]), LA(*[v[0] for _, v in order]), LA(*[v[1] for _, v in order]),
                                      LA(*[v[2] for _, v in order]), LA(*[v[3] for _, v in order])))
            return
        ps, pr = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
        _lib.check(L.fb_halo_buffers(solver, ctypes.byref(ps), ctypes.byref(pr)))
        self.send = torch.as_tensor(_DevArray(ps.value, ns.value), device=dev) if ns.value else None
        self.recv = torch.as_tensor(_DevArray(pr.value, nr.value), device=dev) if nr.value else None
        self.stream = torch.cuda.ExternalStream(L.fb_exchange_stream(solver), device=dev)   # the library's exchange stream
        self._cb = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int)(self._exchange)
        L.fb_set_exchange_cb.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib.check(L.fb_set_exchange_cb(solver, ctypes.cast(self._cb, ctypes.c_void_p), None))

    def describe(self):
        if self.native: return 'ncclSend / ncclRecv per neighbour inside the library (fb_nccl_init), on its exchange stream'
        return 'NCCL isend/irecv per neighbour (torch.distributed, host callback) on the exchange stream'

    def _exchange(self, user, which):
        torch, dist = self.torch, self.dist
        self.calls += 1
        with torch.cuda.stream(self.stream):
            ops = []
            for p, (so, sn, ro, rn) in sorted(self.peers.items()):
                if sn: ops.append(dist.P2POp(dist.isend, self.send[so:so + sn], p, self.group))
                if rn: ops.append(dist.P2POp(dist.irecv, self.recv[ro:ro + rn], p, self.group))
            if ops:
                for req in dist.batch_isend_irecv(ops): req.wait()     # stream-ordered for NCCL: the host does not block


def local_case(workload, rank, world, n=None):
    """Per-rank share of a multi-block synthetic case.  Returns (tree with this rank's zones, join table of this rank,
    peers, description, cells of the whole job)."""
    from . import synth
    if workload in ('tgv256', 'tgv', 'tgv_cut_i', 'tgv_cut_j', 'tgv_cut_k'):
        n = n or 256
        nblk = grid_for(world)
        if workload.startswith('tgv_cut_'):        # all blocks in a row along one direction (tests: thin last tiles on the cut faces)
            nblk = [1, 1, 1]; nblk['ijk'.index(workload[-1])] = world; nblk = tuple(nblk)
        n3 = (n, n, n) if isinstance(n, int) else tuple(n)
        nz = nblk[0] * nblk[1] * nblk[2]
        rank_of = partition(nz, world)
        mine = {z for z, r in rank_of.items() if r == rank}
        t, layout = synth.case_tgv_blocks(n=n, nblk=nblk, only=mine)
        periodic = tuple(b > 1 for b in nblk)
        table, peers = synth.build_joins(layout, nblk, n3, periodic=periodic, rank_of=rank_of, rank=rank, neq=5)
        desc = 'laminar TGV %dx%dx%d per GPU (%dx%dx%d blocks, periodic match joins), SENSOR o3, explicit RK3 ' \
               '(BASELINE configs[1], one block per GPU)' % (n3 + nblk)
        return t, table, peers, desc, nz * n3[0] * n3[1] * n3[2]
    if workload == 'multiblock':
        n = n or 128
        nblk = (2, 2, 2)
        rank_of = partition(8, world)
        mine = {z for z, r in rank_of.items() if r == rank}
        t, layout = synth.case_multiblock(nblk=nblk, n=n, modulo_verif=1000, only=mine, return_layout=True)
        table, peers = synth.build_joins(layout, nblk, (n, n, n), rank_of=rank_of, rank=rank, neq=5)
        desc = '8 x %d^3 blocks over %d GPUs, match joins, Roe o3, implicit LU-SGS 3 sub-iterations, lu_match ' \
               '(BASELINE configs[3])' % (n, world)
        return t, table, peers, desc, 8 * n ** 3
    if workload == 'c5':
        # BASELINE configs[4]: RANS-SA plate in blocks of 256x256x64, 8 blocks (a 2x2x2 brick) per GPU; at 8 GPUs this is
        # the 1024x1024x256 mesh (268 M cells) in 64 blocks.  Weak scaling: the plate grows with the number of GPUs.
        nloc = n or (256, 256, 64)
        g = grid_for(world)
        nb = (2 * g[0], 2 * g[1], 2 * g[2])
        rank_of, zid = {}, 0
        for bk in range(nb[2]):
            for bj in range(nb[1]):
                for bi in range(nb[0]):
                    rank_of[zid] = (bi // 2) + g[0] * ((bj // 2) + g[1] * (bk // 2)); zid += 1
        mine = {z for z, r in rank_of.items() if r == rank}
        nglob = tuple(nloc[d] * nb[d] for d in range(3))
        t, layout = synth.case_flatplate(n=nglob, nblocks=nb, modulo_verif=1000, only=mine, return_layout=True)
        table, peers = synth.build_joins(layout, nb, tuple(nloc), periodic=(False, False, True), rank_of=rank_of, rank=rank, neq=6)
        desc = 'RANS-SA flat plate %dx%dx%d in %d blocks of %dx%dx%d (3dhomo), 8 blocks per GPU, match joins, AUSM o3, implicit ' \
               'LU-SGS 5 sub-iterations, local dt (BASELINE configs[4]%s)' % (nglob + (len(rank_of),) + tuple(nloc) + ('' if world == 8 else ', weak-scaling member',))
        return t, table, peers, desc, nglob[0] * nglob[1] * nglob[2]
    raise ValueError("fast_b200.dist: workload %r has no multi-GPU decomposition" % workload)


def warmup_rank(t, table, peers, device):
    """FastS.warmup on this rank's zones + the cross-rank exchange hook."""
    from .FastS import PyTree as FastS
    from .FastC import PyTree as FastC
    os.environ['FAST_B200_DEVICE'] = str(device)
    (t, tc, metrics) = FastS.warmup(t, table, exchange=lambda solver: HaloExchange(solver, peers, device))
    return t, metrics, FastC.HOOK.get('halo_exchange')
