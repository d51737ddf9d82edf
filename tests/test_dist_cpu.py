"""world_size-2 gloo test of the multi-rank merge (no GPU): nav buffers are assembled by
all-gather (equal shards) or all-reduce over disjoint zero-padded rows (ragged shards), sig
buffers by all-reduce(sum) -- the N>1 path of bench.py / UDFRunner."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from libertem_b200.common import Shape
from libertem_b200.io.memory import MemoryDataSet
from libertem_b200.runner import UDFRunner
from libertem_b200.udf import ApplyMasksUDF, SumUDF, SumSigUDF
from libertem_b200.udf.base import UDFMeta, UDFData


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_parts, nav, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        sig = (4, 4)
        ds_shape = Shape(tuple(nav) + sig, sig_dims=2)
        n = ds_shape.nav.size
        data = torch.zeros(tuple(nav) + sig)
        ds = MemoryDataSet(data=data, num_partitions=n_parts, sig_dims=2)
        parts = list(ds.get_partitions())
        udfs = [ApplyMasksUDF(mask_factories=lambda: np.ones((3, 4, 4), np.float32),
                              mask_count=3, mask_dtype=np.float32),
                SumSigUDF(), SumUDF()]
        cpu = torch.device('cpu')
        for u in udfs:
            u.set_meta(UDFMeta(dataset_shape=ds_shape, dataset_dtype=np.float32,
                               input_dtype=np.float32, device=cpu))
            decl = u.get_result_buffers()
            for b in decl.values():
                b.set_shape_ds(ds_shape, None)
                b.allocate(cpu)
            u.results = UDFData(decl)
        runner = UDFRunner(udfs)
        mine = runner.my_partitions(parts, rank, world)
        # what the local pass would have produced: row i of the nav buffers = f(i)
        for p in mine:
            rows = torch.arange(p.start, p.stop, dtype=torch.float32)
            udfs[0].results.get_buffer('intensity').tensor[p.start:p.stop] = \
                rows[:, None] * torch.tensor([1., 2., 3.])
            udfs[1].results.get_buffer('intensity').tensor[p.start:p.stop] = rows + 0.5
            udfs[2].results.get_buffer('intensity').tensor[:] += float(p.idx + 1)
        damage = np.zeros(n, dtype=bool)
        runner._merge_ranks(dist, udfs, parts, None, damage, cpu)
        full = torch.arange(n, dtype=torch.float32)
        ok = bool(torch.equal(udfs[0].results.get_buffer('intensity').tensor,
                              full[:, None] * torch.tensor([1., 2., 3.])))
        ok &= bool(torch.equal(udfs[1].results.get_buffer('intensity').tensor, full + 0.5))
        want_sig = float(sum(range(1, len(parts) + 1)))
        ok &= bool(torch.all(udfs[2].results.get_buffer('intensity').tensor == want_sig))
        ok &= bool(damage.all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_parts,nav', [(2, (4, 6)), (4, (8, 8)), (3, (5, 7))])
def test_merge_ranks_gloo(n_parts, nav):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_parts, nav, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=10) for _ in range(2))
    assert got == [(0, True), (1, True)]


def _make_custom_udfs():
    from libertem_b200.udf.base import UDF

    class MaxUDF(UDF):
        """non-additive custom merge: running maximum (sig), nav bool flags, a 'single' count"""

        def get_result_buffers(self):
            return {'maxsig': self.buffer(kind='sig', dtype='float32'),
                    'seen': self.buffer(kind='nav', dtype=bool),
                    'count': self.buffer(kind='single', dtype='int64')}

        def process_tile(self, tile):
            pass

        def merge(self, dest, src):
            dest.maxsig[:] = torch.maximum(dest.maxsig, src.maxsig)
            dest.seen[:] = src.seen
            dest.count[:] += src.count

    class FlagUDF(UDF):
        """default merge, bool nav buffer (ragged blocks used to be all-reduced with SUM)"""

        def get_result_buffers(self):
            return {'flag': self.buffer(kind='nav', dtype=bool),
                    'z': self.buffer(kind='nav', extra_shape=(2,), dtype='complex64')}

        def process_tile(self, tile):
            pass

    return MaxUDF(), FlagUDF()


def _worker_custom(rank, world, port, n_parts, nav, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        sig = (4, 4)
        ds_shape = Shape(tuple(nav) + sig, sig_dims=2)
        n = ds_shape.nav.size
        ds = MemoryDataSet(data=torch.zeros(tuple(nav) + sig), num_partitions=n_parts,
                           sig_dims=2)
        parts = list(ds.get_partitions())
        udfs = list(_make_custom_udfs())
        cpu = torch.device('cpu')
        for u in udfs:
            u.set_meta(UDFMeta(dataset_shape=ds_shape, dataset_dtype=np.float32,
                               input_dtype=np.float32, device=cpu))
            decl = u.get_result_buffers()
            for b in decl.values():
                b.set_shape_ds(ds_shape, None)
                b.allocate(cpu)
            u.results = UDFData(decl)
        runner = UDFRunner(udfs)
        mine = runner.my_partitions(parts, rank, world)
        mx, fl = udfs
        for p in mine:
            # what the local merges would have left behind on this rank
            t = mx.results.get_buffer('maxsig').tensor
            t[:] = torch.maximum(t, torch.full_like(t, float(p.idx + 1)))
            mx.results.get_buffer('seen').tensor[p.start:p.stop] = True
            mx.results.get_buffer('count').tensor[:] += p.stop - p.start
            rows = torch.arange(p.start, p.stop)
            fl.results.get_buffer('flag').tensor[p.start:p.stop] = (rows % 3 == 0)
            z = (rows.float()[:, None] * torch.tensor([1., 2.])).to(torch.complex64) * (1 + 2j)
            fl.results.get_buffer('z').tensor[p.start:p.stop] = z
        damage = np.zeros(n, dtype=bool)
        runner._merge_ranks(dist, udfs, parts, None, damage, cpu)
        ok = bool(torch.all(mx.results.get_buffer('maxsig').tensor == float(len(parts))))
        ok &= bool(mx.results.get_buffer('seen').tensor.all())
        ok &= int(mx.results.get_buffer('count').tensor.item()) == n
        full = torch.arange(n)
        ok &= bool(torch.equal(fl.results.get_buffer('flag').tensor, full % 3 == 0))
        zf = (full.float()[:, None] * torch.tensor([1., 2.])).to(torch.complex64) * (1 + 2j)
        ok &= bool(torch.equal(fl.results.get_buffer('z').tensor, zf))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_parts,nav', [(2, (4, 6)), (3, (5, 7))])
def test_merge_ranks_custom_merge_gloo(n_parts, nav):
    """a UDF with a non-additive custom merge (max / bool flags / 'single' count)
    must be merged across ranks by replaying udf.merge, not by all-reduce(SUM); bool and
    complex default-merge nav buffers with ragged blocks are gathered, not summed."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_custom, args=(r, 2, port, n_parts, nav, q))
             for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=10) for _ in range(2))
    assert got == [(0, True), (1, True)]
