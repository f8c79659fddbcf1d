"""World-size-2 gloo test of the point-sharding host logic (2022-entries_b200/sharded.py) on CPU: slicing, the
all-gather of per-rank partials and the combine step, with the CPU oracle standing in for the per-rank CUDA engine."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


class OracleBackend:
    """Checker stand-in for CudaShardBackend: partial = the shard's MSM as a 144-byte Jacobian record (padded to the
    224-byte slot), combine = sum of the gathered records."""
    partial_bytes, out_bytes = 224, 144

    def __init__(self, O, bases):
        self.O, self.bases, self.n = O, bases, len(bases)

    def new_buffer(self, nbytes):
        return torch.zeros(nbytes, dtype=torch.uint8)

    def upload(self, scalars_host):
        return scalars_host

    def full(self, scalars, out):
        out.copy_(torch.from_numpy(self.O.ark_msm(self.bases, scalars.numpy().reshape(-1, 32))))

    def partial(self, scalars, out):
        out.zero_()
        out[:144] = torch.from_numpy(self.O.ark_msm(self.bases, scalars.numpy().reshape(-1, 32)))

    # the peer handshake of ShardedMsm.attach_peers (handles through all_gather_object, barrier, call routing); the checker's
    # "slice of every rank's buckets" is simply its own shard's sum -- the total over ranks is the same
    def peer_export(self, chunks):
        self.exported = chunks
        return torch.full((320,), 7 + chunks, dtype=torch.uint8)

    def peer_attach(self, rank, world, handles):
        assert handles.numel() == 320 * world and int(handles[0]) == 7 + self.exported
        self.attached = (rank, world)

    def peer_detach(self):
        self.attached = None

    def peer_partial(self, scalars, out):
        assert self.attached
        self.peer_calls = getattr(self, "peer_calls", 0) + 1
        self.partial(scalars, out)

    def peer_partial_from_host(self, scalars, out):
        self.peer_partial(scalars, out)

    def combine(self, gathered, count, out):
        recs = gathered.numpy().reshape(count, 224)[:, :144].copy()
        acc = recs[0:1].copy()
        for k in range(1, count):
            acc = self.O.jac_add(acc, recs[k:k + 1])
        out.copy_(torch.from_numpy(acc.reshape(-1)))


def _worker(rank, world, port, n, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib
    import __graft_entry__ as g
    g.load_package()
    SH = importlib.import_module(g.PKG_NAME + ".sharded")
    from oracle import oracle as O
    bases = np.tile(O.gen_points(64, seed=21), (n // 64, 1))
    scalars = O.gen_scalars(n, seed=22)
    lo, hi = SH.shard_bounds(n, world, rank)
    sh = SH.ShardedMsm(OracleBackend(O, bases[lo:hi]))
    got = sh.msm(torch.from_numpy(scalars[lo:hi].reshape(-1).copy()))
    want = O.ark_msm(bases, scalars)
    ok = np.array_equal(O.jac_to_affine(got.numpy()), O.jac_to_affine(want))
    # the same through the bucket reduce-scatter routing
    assert sh.attach_peers(chunks=2) and sh.peers == 2 and sh.b.attached == (rank, world)
    got2 = sh.msm(torch.from_numpy(scalars[lo:hi].reshape(-1).copy()))
    ok = ok and np.array_equal(O.jac_to_affine(got2.numpy()), O.jac_to_affine(want)) and sh.b.peer_calls == 1
    sh.detach_peers()
    ok = ok and sh.peers == 0 and sh.b.attached is None
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_shard_bounds():
    import importlib
    import __graft_entry__ as g
    g.load_package()
    SH = importlib.import_module(g.PKG_NAME + ".sharded")
    assert [SH.shard_bounds(1 << 10, 4, r) for r in range(4)] == [(0, 256), (256, 512), (512, 768), (768, 1024)]
    with pytest.raises(ValueError):
        SH.shard_bounds(10, 4, 0)


def test_two_rank_sharded_msm_matches_single():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, 512, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
