"""Site sharding over the GPUs of one box (one process per GPU, torch.distributed for the plumbing).

Sites are independent given Q, so the only cross-GPU traffic of the sampler is the sum of the sufficient
statistics: one all-reduce of n + n^2 + 1 doubles per iteration for the rate-updating samplers (bf/ks/mt/ksmt),
one all-reduce of the whole N x (n + n^2 + 1) block at the end of a fixed-Q run.  The host-side rate update is
replicated: every rank seeds R's Mersenne-Twister identically and sees the same reduced row, so Q stays identical
everywhere without a broadcast.
"""
import numpy as np


_STREAMS = []


def shard(n_sites, rank, world):
    """Contiguous block [start, start + count) of the global site axis owned by `rank`."""
    base, extra = divmod(int(n_sites), int(world))
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def shard_tree(z, rank, world):
    """The tree restricted to this rank's sites, and the global index of its first site."""
    from .tree import PhyloTree
    z = PhyloTree.from_mapping(z)
    st = z.states if z.states.ndim == 2 else z.states[None, :]
    start, count = shard(st.shape[0], rank, world)
    if count == 0:
        raise ValueError("fewer sites than ranks")
    return PhyloTree(z.edge, z.edge_length, st[start:start + count], z.maps, z.mapnames, z.tip_label), start


class _DevPtr:
    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


def allreduce_callback(stream_ptr=None, group=None):
    """pm_allreduce_fn over torch.distributed (NCCL): sums `count` doubles at a device address, ordered on the
    chain's CUDA stream."""
    import torch
    import torch.distributed as dist

    def cb(ctx, dev_ptr, count):
        try:
            t = torch.as_tensor(_DevPtr(dev_ptr, count), device="cuda")
            if stream_ptr:
                with torch.cuda.stream(torch.cuda.ExternalStream(int(stream_ptr))):
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return 0
        except Exception as e:  # the C side turns a non-zero return into PM_ERR_CUDA
            import sys
            sys.stderr.write("phylomap_b200 allreduce failed: %r\n" % (e,))
            return 1

    return cb


def nccl_clique(rank, world, group=None):
    """(unique id, rank, world) for the `nccl=` option: rank 0 asks the library for an ncclUniqueId and torch.distributed
    (any backend) hands it to the other ranks; from then on the library calls ncclAllReduce itself."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    from . import capi
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        raw = C.create_string_buffer(128)
        err = C.create_string_buffer(512)
        capi.check(capi.lib().pm_nccl_unique_id(raw, err, 512), err)
        buf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
    if dist.get_backend(group) == "nccl":
        dev = buf.cuda()
        dist.broadcast(dev, src=0, group=group)
        buf = dev.cpu()
    else:
        dist.broadcast(buf, src=0, group=group)
    return bytes(buf.numpy().tobytes()), rank, world


def sharded_options(rank, world, start, stream_ptr=None, native=True, group=None):
    """Keyword options for api.* that make a call one shard of a `world`-process run.  native: the library's own NCCL
    all-reduce (default); otherwise the torch.distributed callback, which needs the stream the chain launches on."""
    opts = {"site_offset": start, "device": rank}
    if world > 1:
        if native:
            opts["nccl"] = nccl_clique(rank, world, group)
        else:
            if not stream_ptr:
                import torch
                st = torch.cuda.Stream(device=rank)
                _STREAMS.append(st)  # the chain launches on it: keep it alive
                stream_ptr = st.cuda_stream
            opts["allreduce"] = allreduce_callback(stream_ptr, group)
    if stream_ptr:
        opts["stream"] = stream_ptr
    return opts


def combine_rows_host(rows, group=None):
    """CPU-side (gloo) twin of the device all-reduce, used by the host-logic tests: sums a numpy row block."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(rows))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()
