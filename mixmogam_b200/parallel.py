"""
Multi-GPU plumbing for the EMMAX path: one process per GPU, SNPs sharded along the SNP axis
(SURVEY.md section 8e).  torch.distributed (NCCL over NVLink/NVSwitch) carries the one real exchange
step of each stage:

    kinship : every rank forms the integer Gram of its SNP slice -> all_reduce(SUM, int32) of the n x n
              Gram.  Integer addition is exact and order independent, so K is bit-identical for any
              number of ranks.
    scan    : the quadratic form A = R'R of the int8 scan (2 n^3 / 2 FP64 flops, 28 ms at n = 10k -- as long as an
              8-way shard of the scan itself) is formed cooperatively: every rank multiplies its block of the
              rows of R (mmg_mat_syrk_rows) and the n x n partial sums are all_reduced (SUM, float64).  Then every
              rank scans its own SNP slice; the per-SNP outputs are all_gathered.  Permutations: all_reduce(MAX)
              of the per-permutation ratios.

PyTorch is plumbing here (process group + collectives on device pointers owned by libmixmogam_b200).
"""
import numpy as np


def shard_range(m, rank, world):
    """Contiguous SNP slice [begin, end) of rank `rank`; slices are 128-aligned (the scan's row-block
    size) except the last, differ by at most one block, and cover [0, m) exactly."""
    if world <= 1:
        return 0, m
    blocks = (m + 127) // 128
    base, extra = divmod(blocks, world)
    b0 = rank * base + min(rank, extra)
    b1 = b0 + base + (1 if rank < extra else 0)
    return min(b0 * 128, m), min(b1 * 128, m)


class _DevPtr(object):
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can wrap it zero-copy."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr, 'data': (int(ptr), False),
                                         'version': 2, 'strides': None}


def gram_as_tensor(ctx):
    import torch
    ptr, n, ld = ctx.kinship_gram_ptr()
    return torch.as_tensor(_DevPtr(ptr, (ld, ld), '<i4'), device='cuda:%d' % ctx.device)


def allreduce_gram(ctx, group=None):
    """Sum the per-rank partial Grams in place (int32, exact)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    t = gram_as_tensor(ctx)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    torch.cuda.synchronize(ctx.device)


def world_size(group=None):
    """Number of ranks of the initialised process group (1 without torch.distributed)."""
    try:
        import torch.distributed as dist
    except Exception:
        return 1
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


def nccl_backend(group=None):
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_backend(group) == 'nccl'


def split_rows(rows, rank, world):
    """Contiguous block [begin, end) of `rows` rows for `rank`: sizes differ by at most one, blocks cover [0, rows)."""
    base, extra = divmod(rows, world)
    b = rank * base + min(rank, extra)
    return b, b + base + (1 if rank < extra else 0)


def quad_form_sharded(ctx, R, group=None):
    """A = R'R (DeviceMatrix, row-major lower triangle valid) with the n^3 product split over the ranks:
    rank r forms R[rows_r, :]' R[rows_r, :] and the partial sums are all-reduced over NCCL (SUM, float64)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    b, e = split_rows(R.shape[0], rank, world)
    A = ctx.syrk_rows(R, b, e - b)
    ptr, ld = A.device_ptr()
    n = A.shape[0]
    t = torch.as_tensor(_DevPtr(ptr, (n, ld), '<f8'), device='cuda:%d' % ctx.device)
    ctx.sync()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    torch.cuda.synchronize(ctx.device)
    return A


def allgather_rows(local, m_total, group=None, device=None):
    """Concatenate per-rank 1-D float64 arrays (shard_range order) into the full length-m_total array.
    Works with NCCL (device tensors) and gloo (CPU tensors)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return np.asarray(local)
    world = dist.get_world_size(group)
    sizes = [shard_range(m_total, r, world) for r in range(world)]
    maxlen = max(e - b for b, e in sizes)
    backend = dist.get_backend(group)
    dev = torch.device('cuda:%d' % device) if (backend == 'nccl' and device is not None) else torch.device('cpu')
    buf = torch.zeros(maxlen, dtype=torch.float64, device=dev)
    buf[:len(local)] = torch.as_tensor(np.ascontiguousarray(local), dtype=torch.float64).to(dev)
    outs = [torch.empty(maxlen, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    return np.concatenate([o[:e - b].cpu().numpy() for o, (b, e) in zip(outs, sizes)])


def allreduce_max(values, group=None, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return values
    backend = dist.get_backend(group)
    dev = torch.device('cuda:%d' % device) if (backend == 'nccl' and device is not None) else torch.device('cpu')
    t = torch.as_tensor(np.ascontiguousarray(values), dtype=torch.float64).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t.cpu().numpy()


def calc_ibs_kinship_sharded(local_snps, m_total, snps_data_format='diploid_int', scaled=True, ctx=None, group=None,
                             impl='auto'):
    """kinship.calc_ibs_kinship over SNP shards: partial integer Gram -> int32 all-reduce -> replicated
    FP64 finalisation.  Returns the kinship as a DeviceMatrix (identical on every rank)."""
    from . import _lib, kinship
    ctx = ctx or _lib.get_context()
    kinship.partial_ibs_gram(local_snps, snps_data_format, impl=impl, ctx=ctx)
    allreduce_gram(ctx, group)
    K, _ = ctx.kinship_finalize(kinship._coding(snps_data_format), m_total, scaled)
    return K
