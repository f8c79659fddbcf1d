"""Point sharding across the GPUs of one box (BASELINE.json config 3).

MSM is linear in the point set:  sum_i s_i P_i = sum_g  sum_{i in shard g} s_i P_i.  Rank g keeps the contiguous slice
[g*n/G, (g+1)*n/G) of the bases resident (tables built at init) and receives the same slice of every scalar batch.
One exchange per MSM: each rank's partial sum (one XYZZ record, 224 bytes) is all-gathered as bytes -- NCCL has no
reduction operator for curve addition, so the north-star's "allreduce of the partial sums" is an all-gather followed
by the same G-point sum on every rank -- and every rank ends with the full result, like an all-reduce.

The reference has no multi-GPU path at all (device 0 hard-coded: P1A/yrrid/yrrid-msm/MSM.cu:302,
P1A/matter-labs/src/lib.rs:86).

The per-rank engine is injected so that the host logic (slicing, exchange, combine order) runs under gloo on CPU in the
tests, with a checker standing in for the CUDA engine; the product path constructs it over `Engine` (libb200msm.so).
"""
import os

import torch
import torch.distributed as dist


def shard_bounds(n, world, rank):
    """Contiguous slice of rank `rank`; n must divide evenly (the harness sizes are powers of two)."""
    if n % world:
        raise ValueError(f"{n} points do not shard evenly over {world} ranks")
    per = n // world
    return rank * per, (rank + 1) * per


class CudaShardBackend:
    """One rank's engine: partial sums and the combine step run in libb200msm.so on this rank's GPU."""

    def __init__(self, engine, torch_stream=None):
        self.e = engine
        self.device = torch.device("cuda", engine.device)
        self.stream = torch_stream or torch.cuda.Stream(self.device)
        engine.set_stream(self.stream.cuda_stream)
        from . import PARTIAL_BYTES, PROJECTIVE_BYTES
        self.partial_bytes, self.out_bytes = PARTIAL_BYTES, PROJECTIVE_BYTES
        self.n = engine.n
        self.dev_scalars = None
        self.result_stream = self.stream     # where results of the engine's device-resident calls are ordered

    def set_pipelined(self, on=True):
        """Consecutive MSMs overlap (the tail of one under the planning kernels of the next, Engine.set_pipelined): the partial,
        the all-gather and the combine step then live on the engine's tail stream."""
        self.e.set_pipelined(on)
        self.result_stream = torch.cuda.ExternalStream(self.e.result_stream(), device=self.device) if on else self.stream

    def new_buffer(self, nbytes):
        return torch.zeros(nbytes, dtype=torch.uint8, device=self.device)

    def upload(self, scalars_host):
        """Pinned host tensor (n_local*32 bytes) -> device, on the engine's stream."""
        if self.dev_scalars is None:
            self.dev_scalars = torch.empty(self.n * 32, dtype=torch.uint8, device=self.device)
        with torch.cuda.stream(self.stream):
            self.dev_scalars.copy_(scalars_host, non_blocking=True)
        return self.dev_scalars

    def partial(self, scalars_dev, out):
        self.e.msm_partial_device(scalars_dev.data_ptr(), self.n, out.data_ptr())

    def partial_from_host(self, scalars_host, out):
        """Pinned host scalars -> partial sum; the library chunks the upload against the accumulation."""
        self.e.msm_partial_host(scalars_host.data_ptr(), self.n, out.data_ptr())

    def full(self, scalars_dev, out):
        self.e.msm_device(scalars_dev.data_ptr(), self.n, out.data_ptr())

    def combine(self, gathered, count, out):
        self.e.combine_device(gathered.data_ptr(), count, out.data_ptr())

    # ---- bucket reduce-scatter over NVLink (b200msm_peer_*): optional, set up by ShardedMsm.attach_peers
    def peer_export(self, chunks):
        return torch.from_numpy(self.e.peer_export(chunks))

    def peer_attach(self, rank, world, handles):
        self.e.peer_attach(rank, world, handles.numpy())

    def peer_detach(self):
        self.e.peer_detach()

    def peer_partial(self, scalars_dev, out):
        self.e.msm_peer_partial_device(scalars_dev.data_ptr(), self.n, out.data_ptr())

    def peer_partial_from_host(self, scalars_host, out):
        self.e.msm_peer_partial_host(scalars_host.data_ptr(), self.n, out.data_ptr())


class ShardedMsm:
    """multi_scalar_mult over G ranks.  Every rank calls msm() with ITS slice of the scalars and gets the full result."""

    def __init__(self, backend, group=None):
        self.b = backend
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.partial = backend.new_buffer(backend.partial_bytes)
        self.gathered = backend.new_buffer(backend.partial_bytes * self.world)
        self.out = backend.new_buffer(backend.out_bytes)
        self.peers = 0          # chunks the peer mapping was attached with; 0 = every rank reduces all of its own buckets
        # pageable host scalars are staged by a pool of host threads per process (csrc/copy_pool.h, default: half the cores):
        # G ranks on one host share its cores
        if self.world > 1 and "B200MSM_COPY_THREADS" not in os.environ:
            os.environ["B200MSM_COPY_THREADS"] = str(max(4, (os.cpu_count() or 8) // self.world))

    def attach_peers(self, chunks=1):
        """Collective: switch to the bucket reduce-scatter -- rank g reduces the buckets [g B / G, (g + 1) B / G) of EVERY rank,
        reading the peers' partial bucket sums over NVLink inside its level-0 reduction kernel (csrc/kernels_cold.cuh
        k_reduce_slices).  The handles (CUDA IPC) travel through one all_gather on the group's default backend.
        chunks: how many chunks a host-memory call splits its upload into (1 for device-resident scalars)."""
        b = self.b
        if self.world == 1 or not hasattr(b, "peer_export"):
            return False
        if self.peers:
            self.detach_peers()
        mine = b.peer_export(chunks)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, mine.numpy().tobytes(), group=self.group)
        import numpy as np
        handles = torch.from_numpy(np.frombuffer(b"".join(gathered), np.uint8).copy())
        b.peer_attach(self.rank, self.world, handles)
        dist.barrier(group=self.group)          # nobody runs a peer call before every rank has mapped every other rank
        self.peers = chunks
        return True

    def detach_peers(self):
        if self.peers:
            stream = getattr(self.b, "stream", None)
            if stream is not None:
                stream.synchronize()
            dist.barrier(group=self.group)      # every rank has finished reading this rank's buckets
            self.b.peer_detach()
            self.peers = 0

    def msm_device(self, scalars_dev, host=False):
        """scalars already on this rank's device (host=True: in pinned host memory, uploaded by the backend).
        Asynchronous on the backend's stream; returns the device buffer holding the 144-byte Jacobian result
        (identical on every rank)."""
        b = self.b
        if self.world == 1:
            b.full(b.upload(scalars_dev) if host else scalars_dev, self.out)
            return self.out
        if self.peers and host:
            b.peer_partial_from_host(scalars_dev, self.partial)
        elif self.peers:
            b.peer_partial(scalars_dev, self.partial)
        elif host and hasattr(b, "partial_from_host"):
            b.partial_from_host(scalars_dev, self.partial)
        else:
            b.partial(b.upload(scalars_dev) if host else scalars_dev, self.partial)
        stream = getattr(b, "result_stream", getattr(b, "stream", None))
        if stream is not None:
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(self.gathered, self.partial, group=self.group)
        else:
            dist.all_gather_into_tensor(self.gathered, self.partial, group=self.group)
        b.combine(self.gathered, self.world, self.out)
        return self.out

    def msm(self, scalars_host):
        """End-to-end call: this rank's scalar slice in (pinned) host memory -> 144-byte result on the host."""
        out = self.msm_device(scalars_host, host=True)
        stream = getattr(self.b, "result_stream", getattr(self.b, "stream", None))
        if stream is not None:
            with torch.cuda.stream(stream):
                host = out.to("cpu", non_blocking=False)
            return host
        return out.clone()
