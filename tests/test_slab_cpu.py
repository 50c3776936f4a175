"""
World-size-2 tests of the slab decomposition plumbing on CPU (gloo): scatter / halo exchange /
gather and the periodicity test used to decide whether U halos must be exchanged.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from odil_b200.engine import wraps_axis0
from odil_b200.slab import HALO, SlabInfo
from oracle import odil_oracle as orc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        slab = SlabInfo.from_environment()
        assert slab is not None and slab.world == world and slab.rank == rank
        g = torch.arange(16 * 3 * 5, dtype=torch.float64).reshape(16, 3, 5)
        gc = torch.arange(8 * 3 * 5, dtype=torch.float64).reshape(8, 3, 5) * 0.5
        ref, refc = slab.scatter(g), slab.scatter(gc)
        z0, n = slab.owned_range(g.shape)
        assert (z0, n) == (rank * 8, 8) and tuple(ref.shape) == (8 + 2 * HALO, 3, 5)
        assert torch.equal(slab.owned(ref), g[z0:z0 + n])
        # wipe the halos, exchange two arrays of different sizes in one batch, compare with scatter (periodic ring)
        a, b = ref.clone(), refc.clone()
        for t in (a, b):
            t[:HALO] = -1
            t[-HALO:] = -1
        slab.exchange([a, b], width=HALO)
        assert torch.equal(a, ref) and torch.equal(b, refc)
        # width-1 exchange only touches the innermost halo plane
        c = ref.clone()
        c[:HALO] = -1
        c[-HALO:] = -1
        slab.exchange([c], width=1)
        assert torch.equal(c[HALO - 1:-(HALO - 1) or None], ref[HALO - 1:-(HALO - 1) or None])
        assert torch.all(c[0] == -1) and torch.all(c[-1] == -1)
        assert torch.equal(slab.gather(ref), g)
        s = torch.tensor([float(rank + 1), 10.0])
        slab.all_reduce_sum(s)
        assert s.tolist() == [3.0, 20.0]
        # accumulate = transpose of exchange: partial sums computed for the planes just outside the slab are added
        # to their owners' first / last owned planes (ring)
        p = torch.zeros(8 + 2 * 3, 2, dtype=torch.float64)           # gradient layout: 3 halo planes per side
        p[3:11] = 1.0                                                 # owned planes
        p[2] = 10.0 * (rank + 1)                                      # partial for the plane below (lower neighbour's)
        p[11] = 100.0 * (rank + 1)                                    # partial for the plane above (upper neighbour's)
        slab.accumulate([(p[2:3], p[11:12], p[3:4], p[10:11])])
        other = 1 - rank                                              # world 2: both neighbours are the other rank
        assert torch.all(p[3] == 1.0 + 100.0 * (other + 1))           # first owned += lower neighbour's upper partial
        assert torch.all(p[10] == 1.0 + 10.0 * (other + 1))           # last owned  += upper neighbour's lower partial
        assert torch.all(p[4:10] == 1.0)
        results[rank] = "ok"
    except Exception as e:  # pragma: no cover
        results[rank] = repr(e)
    finally:
        dist.destroy_process_group()


def test_slab_exchange_gloo_world2():
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        results = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, results)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
        assert dict(results) == {0: "ok", 1: "ok"}


def test_single_rank_exchange_is_periodic_copy():
    slab = SlabInfo(0, 1)
    g = torch.arange(6 * 4, dtype=torch.float32).reshape(6, 4)
    a = slab.scatter(g)
    b = a.clone()
    b[:HALO] = 0
    b[-HALO:] = 0
    slab.exchange([b])
    assert torch.equal(a, b)
    assert SlabInfo.from_environment() is None  # no process group in this process


def test_wraps_axis0():
    offsets, table, rr = orc.poisson_plan(3, [0.1, 0.1, 0.1])
    assert not wraps_axis0((16, 16, 16), offsets, rr, table.reshape(-1, len(offsets)))      # Dirichlet rows
    per = np.ones((1, 3))
    assert wraps_axis0((16,), [(0,), (-1,), (1,)], (0,), per)                                 # periodic Laplacian
    assert not wraps_axis0((16, 8), [(0, 0), (0, 1)], (0, 0), np.ones((1, 2)))                # wraps along axis 1 only
    with pytest.raises(ValueError):
        SlabInfo(0, 3).check((16, 4))
