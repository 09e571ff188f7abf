"""N>1 path: join-table sharding + exchange pattern under gloo (CPU, world_size 2 and 4), and the full sharded solve
under NCCL when the box has at least two GPUs."""
import os, subprocess, sys
import numpy
import pytest
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, 'tests', 'mgpu_parity.py')


def _torchrun(nproc, args, port):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(nproc), '--master-addr', '127.0.0.1',
           '--master-port', str(port), DRIVER] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)


@pytest.mark.parametrize('world', [2, 4])
def test_sharded_joins_gloo(world):
    r = _torchrun(world, ['--cpu'], 29610 + world)
    assert r.returncode == 0 and 'MGPU_CPU_OK' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_partition_and_tables():
    sys.path.insert(0, ROOT)
    from fast_b200 import synth, dist as fdist
    assert fdist.partition(8, 2) == {0: 0, 1: 0, 2: 0, 3: 0, 4: 1, 5: 1, 6: 1, 7: 1}
    with pytest.raises(ValueError): fdist.partition(1, 2)
    layout = {(0, 0, 0): 0, (1, 0, 0): 1}
    t0, p0 = synth.build_joins(layout, (2, 1, 1), (4, 4, 4), periodic=(True, False, False), rank_of={0: 0, 1: 1}, rank=0)
    t1, p1 = synth.build_joins(layout, (2, 1, 1), (4, 4, 4), periodic=(True, False, False), rank_of={0: 0, 1: 1}, rank=1)
    assert p0[1][1] == p1[0][3] and p0[1][3] == p1[0][1]          # what 0 sends is what 1 receives and vice versa
    assert int(t0[0]) == 4 and sorted(t0[2:2 + 28:7]) == [1, 1, 2, 2]


@pytest.mark.parametrize('world', [1, 2, 8])
def test_c5_bricks(world):
    """BASELINE configs[4] family: every rank owns a 2x2x2 brick of blocks; what a rank sends to a peer is what the peer expects."""
    sys.path.insert(0, ROOT)
    from fast_b200 import dist as fdist
    from fast_b200.FastC.PyTree import Internal
    n = (4, 4, 4)
    per = {}
    for rank in range(world):
        t, table, peers, desc, cells = fdist.local_case('c5', rank, world, n=n)
        assert len(Internal.getZones(t)) == 8 and cells == 8 * world * 64
        per[rank] = peers
    for r, peers in per.items():
        for p, (so, sn, ro, rn) in peers.items():
            assert per[p][r][3] == sn and per[p][r][1] == rn
    if world == 8: assert '1024' not in desc and 'configs[4]' in desc


@pytest.mark.gpu
@pytest.mark.parametrize('workload', ['tgv', 'multiblock', 'c5', 'tgv_cut_i', 'tgv_cut_j', 'tgv_cut_k'])
def test_sharded_solve_nccl(workload):
    import torch
    if torch.cuda.device_count() < 2: pytest.skip("needs 2 GPUs (run: gpurun --gpus 2 -- python -m pytest tests -m gpu -k nccl)")
    r = _torchrun(2, [workload], 29631)
    assert r.returncode == 0 and 'OK' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
