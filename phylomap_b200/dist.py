"""Site sharding over the GPUs of one box (one process per GPU, torch.distributed for the plumbing).

Sites are independent given Q, so the only cross-GPU traffic of the sampler is the sum of the sufficient
statistics: one all-reduce of n + n^2 + 1 doubles per iteration for the rate-updating samplers (bf/ks/mt/ksmt),
one all-reduce of the whole N x (n + n^2 + 1) block at the end of a fixed-Q run.  The host-side rate update is
replicated: every rank seeds R's Mersenne-Twister identically and sees the same reduced row, so Q stays identical
everywhere without a broadcast.
"""
import numpy as np


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


def sharded_options(rank, world, start, stream_ptr=None):
    """Keyword options for api.* that make a call one shard of a `world`-process run."""
    opts = {"site_offset": start, "device": rank}
    if world > 1:
        opts["allreduce"] = allreduce_callback(stream_ptr)
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
