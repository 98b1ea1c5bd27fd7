"""Host-side logic of the multi-rank path on CPU: slab ownership, ghosted local arrays and the
rank wiring, exercised with world_size-2 and -3 `gloo` process groups (no GPU, no NCCL)."""
import os
import socket

import numpy as np
import pytest

import cases
from taxila_lbm_b200 import geometry as geo
from taxila_lbm_b200 import slab


def test_slab_range_matches_dmda_ownership():
    for NZ in (8, 37, 512, 513):
        for n in (1, 2, 3, 4, 8):
            ranges = [slab.slab_range(NZ, n, r) for r in range(n)]
            assert ranges[0][0] == 0 and sum(zl for _, zl in ranges) == NZ
            for (zs, zl), (zs2, _) in zip(ranges, ranges[1:]):
                assert zs + zl == zs2
            base, extra = divmod(NZ, n)
            assert [zl for _, zl in ranges] == [base + 1] * extra + [base] * (n - extra)
    with pytest.raises(ValueError):
        slab.slab_range(4, 8, 0)


def test_neighbours_ring():
    assert slab.neighbours(4, 0, True) == (3, 1)
    assert slab.neighbours(4, 3, True) == (2, 0)
    assert slab.neighbours(4, 0, False) == (-1, 1)
    assert slab.neighbours(4, 3, False) == (2, -1)
    assert slab.neighbours(1, 0, True) == (0, 0)


def test_local_arrays_reassemble():
    cfg, walls, rho = cases.porous_3d(16, NZ=21, rmin=3.0, rmax=5.0, order=8)
    R = cfg.stencil_size_rho
    parts_w, parts_r = [], []
    for r in range(4):
        c, w_rg, r_rg = slab.local_arrays(cfg, walls, rho, 4, r)
        assert w_rg.shape == (c.zl + 2 * R, cfg.NY + 2 * R, cfg.NX + 2 * R)
        parts_w.append(geo.owned(w_rg, R, 3))
        parts_r.append(geo.owned(r_rg, R, 3))
    assert np.array_equal(slab.assemble(parts_w), walls)
    assert np.array_equal(slab.assemble(parts_r), rho)


def test_local_bc_values_follow_the_slabs():
    """Face arrays per rank as BCSetUp sizes them (lbm_bc.F90:127-213)."""
    from taxila_lbm_b200 import config as tc

    cfg, walls, rho, bcs = cases.drainage_3d(N=8, NZ=11, x_bc=tc.BC_NEUMANN)
    rng = np.random.default_rng(0)
    bcs = {b: rng.uniform(size=v.shape) for b, v in bcs.items()}
    parts = {b: [] for b in bcs}
    for r in range(3):
        zs, zl = slab.slab_range(cfg.NZ, 3, r)
        loc = slab.local_bc_values(cfg, bcs, 3, r)
        assert (tc.BOUNDARY_ZM in loc) == (r == 0) and (tc.BOUNDARY_ZP in loc) == (r == 2)
        for b in (tc.BOUNDARY_XM, tc.BOUNDARY_XP):
            assert loc[b].shape == (zl, cfg.NY, 3, 2)
            parts[b].append(loc[b])
    for b in (tc.BOUNDARY_XM, tc.BOUNDARY_XP):
        assert np.array_equal(np.concatenate(parts[b], axis=0), bcs[b])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, periodic_z, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg, walls, rho = cases.porous_3d(16, NZ=19, rmin=3.0, rmax=5.0, order=8, periodic=(1, 1, periodic_z))
        R = cfg.stencil_size_rho
        c, w_rg, r_rg = slab.local_arrays(cfg, walls, rho, world, rank)
        # rank 0 makes the communicator id and broadcasts it, like bench.py does with the NCCL id
        ids = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        assert ids[0] == bytes(range(128))
        down, up = slab.neighbours(world, rank, bool(periodic_z))
        # what the halo exchange moves: my top R owned planes -> up's bottom ghost planes, and
        # my bottom R owned planes -> down's top ghost planes; check the host
        # arrays already satisfy that relation, for walls and rho alike
        for arr in (w_rg, r_rg):
            own = geo.owned(arr, R, 3)
            top = torch.from_numpy(np.ascontiguousarray(arr[c.zl:c.zl + R, R:-R, R:-R]))  # top R owned planes
            bot = torch.from_numpy(np.ascontiguousarray(arr[R:2 * R, R:-R, R:-R]))  # bottom R owned planes
            assert np.array_equal(own[-R:], top.numpy()) and np.array_equal(own[:R], bot.numpy())
            reqs, from_down, from_up = [], torch.empty_like(top), torch.empty_like(bot)
            # tag 0: planes travelling up, tag 1: planes travelling down (with two ranks on a
            # periodic ring both neighbours are the same peer)
            if up >= 0:
                reqs.append(dist.isend(top, up, tag=0))
                reqs.append(dist.irecv(from_up, up, tag=1))
            if down >= 0:
                reqs.append(dist.isend(bot, down, tag=1))
                reqs.append(dist.irecv(from_down, down, tag=0))
            for rq in reqs:
                rq.wait()
            if down >= 0:
                assert np.array_equal(arr[:R, R:-R, R:-R], from_down.numpy())
            if up >= 0:
                assert np.array_equal(arr[c.zl + R:, R:-R, R:-R], from_up.numpy())
        if not periodic_z:
            if rank == 0:
                assert np.all(w_rg[:R] == 999.0)
            if rank == world - 1:
                assert np.all(w_rg[-R:] == 999.0)
        dist.barrier()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, "FAIL %r" % (e,)))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,periodic_z", [(2, 1), (2, 0), (3, 1)])
def test_halo_relation_gloo(world, periodic_z):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, periodic_z, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert res == [(r, "ok") for r in range(world)], res
    assert all(p.exitcode == 0 for p in procs)
