"""
Multi-GPU plumbing for the EMMAX path: one process per GPU, SNPs sharded along the SNP axis
(SURVEY.md section 8e).  torch.distributed (NCCL over NVLink/NVSwitch) carries the one real exchange
step of each stage:

    kinship : every rank forms the integer Gram of its SNP slice -> all_reduce(SUM, int32) of the valid
              256 x 256 blocks of the Gram, packed back to back (half the bytes of the padded square).
              Integer addition is exact and order independent, so K is bit-identical for any number of ranks.
    eigen   : rank 0 solves eigh(K), rank 1 eigh(S(K+I)S), at the same time; both bases are broadcast.
    scan    : the quadratic form A = R'R of the int8 scan (n^3 work, as long as an 8-way shard of the scan
              itself) is formed cooperatively: every rank forms an equal range of the 256 x 256 blocks of its
              lower triangle as exact int8 digit-plane products on the tensor cores (mmg_quad_form_tiles) and
              one in-place all_gather completes the packed matrix everywhere.  Then every rank scans its own
              SNP slice; the per-SNP outputs are all_gathered on the device.  Permutations: all_reduce(MAX)
              of the per-permutation ratios.

Every collective is enqueued on the library's own CUDA stream (torch.cuda.ExternalStream over
mmg_stream_handle): it is ordered between the library's kernels by the stream, the host never waits for it.
PyTorch is plumbing here (process group + collectives on device pointers owned by libmixmogam_b200).
"""
import contextlib

import numpy as np

RESULT_KEYS = ('ps', 'f_stats', 'rss', 'var_perc', 'xx')       # row order of mmg_emmax_scan_quad_dev's output
GATHERED_KEYS = RESULT_KEYS[:4]                                 # what the sharded scan hands back (linear_models.py:1351-1352)


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


def split_rows(rows, rank, world):
    """Contiguous block [begin, end) of `rows` rows for `rank`: sizes differ by at most one, blocks cover [0, rows)."""
    base, extra = divmod(rows, world)
    b = rank * base + min(rank, extra)
    return b, b + base + (1 if rank < extra else 0)


def slot_range(slots, rank, world):
    """Blocks [begin, begin + count) of the packed quadratic form that rank `rank` forms, and the common (padded) count
    `per` every rank contributes to the all-gather: per = ceil(slots / world), the last ranks may own fewer real blocks."""
    per = -(-slots // world)
    begin = rank * per
    return begin, max(0, min(per, slots - begin)), per


class _DevPtr(object):
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can wrap it zero-copy."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr, 'data': (int(ptr), False),
                                         'version': 2, 'strides': None}


def _tensor(ctx, ptr, shape, typestr):
    import torch
    return torch.as_tensor(_DevPtr(ptr, shape, typestr), device='cuda:%d' % ctx.device)


def mat_as_tensor(ctx, M):
    """A DeviceMatrix as a torch tensor over the same memory."""
    ptr, ld = M.device_ptr(sync=False)
    return _tensor(ctx, ptr, (M.shape[0], ld), '<f8')


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


# ----------------------------------------------------------------------------------------------------------
# collectives on the library's stream, timed with events on that stream
# ----------------------------------------------------------------------------------------------------------
def lib_stream(ctx):
    """The library's CUDA stream as a torch stream (cached on the context)."""
    import torch
    s = getattr(ctx, '_torch_stream', None)
    if s is None:
        s = torch.cuda.ExternalStream(ctx.stream_ptr(), device='cuda:%d' % ctx.device)
        ctx._torch_stream = s
    return s


@contextlib.contextmanager
def on_lib_stream(ctx, timer=None):
    """Everything torch enqueues inside runs on the library's stream; `timer` names the stage the device time goes to
    (collective_timers)."""
    import torch
    s = lib_stream(ctx)
    with torch.cuda.stream(s):
        if timer is None:
            yield s
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        yield s
        e1.record(s)
        pend = getattr(ctx, '_coll_pending', None)
        if pend is None:
            pend = ctx._coll_pending = []
        pend.append((timer, e0, e1))


def collective_timers(ctx, reset=False):
    """Seconds of device time per collective stage ('allreduce', 'allgather', 'broadcast') since the last reset.  Waits
    for the recorded events."""
    acc = getattr(ctx, '_coll_seconds', None)
    if acc is None:
        acc = ctx._coll_seconds = {}
    for name, e0, e1 in getattr(ctx, '_coll_pending', []) or []:
        e1.synchronize()
        acc[name] = acc.get(name, 0.0) + 1e-3 * e0.elapsed_time(e1)
    ctx._coll_pending = []
    out = dict(acc)
    if reset:
        ctx._coll_seconds = {}
    return out


# ----------------------------------------------------------------------------------------------------------
# kinship
# ----------------------------------------------------------------------------------------------------------
def allreduce_gram(ctx, group=None):
    """Sum the per-rank partial Grams in place (int32, exact): only the valid blocks travel, packed contiguously."""
    import torch.distributed as dist
    if world_size(group) == 1:
        return
    ptr, count = ctx.kinship_gram_tri(0)
    t = _tensor(ctx, ptr, (count,), '<i4')
    with on_lib_stream(ctx, 'allreduce'):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    ctx.kinship_gram_tri(1)


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


# ----------------------------------------------------------------------------------------------------------
# eigendecompositions: solved once each, on different ranks, broadcast (BASELINE.json north_star)
# ----------------------------------------------------------------------------------------------------------
def shared_eigen(lmm, group=None):
    """(eig_L, eig_R) of `lmm` (same model on every rank): rank 0 runs eigh(K), rank 1 eigh(S(K+I)S) concurrently, the
    eigenvectors (n x n FP64 each) and eigenvalues are broadcast over NCCL.  One rank: both locally."""
    import torch
    import torch.distributed as dist
    from ._lib import DeviceMatrix, LazyHostArray
    from .linear_models import EigenDict
    world = world_size(group)
    if world == 1:
        return lmm._get_eigen_L_(), lmm._get_eigen_R_(X=lmm.X)
    ctx, n = lmm.ctx, lmm.n
    rank = dist.get_rank(group)
    q = lmm.X.shape[1]
    root_L, root_R = 0, 1 % world
    eig_L = lmm._get_eigen_L_() if rank == root_L else None
    eig_R = lmm._get_eigen_R_(X=lmm.X) if rank == root_R else None
    out = []
    for root, eig, rows in ((root_L, eig_L, None), (root_R, eig_R, (q, n))):
        if eig is not None:
            U = eig['vectors'].dev
            w_full = np.zeros(n)
            w_full[n - len(eig['values']):] = eig['values']
        else:
            U = DeviceMatrix(ctx, n, n)
            w_full = np.zeros(n)
        with on_lib_stream(ctx, 'broadcast'):                  # (every torch operation on these tensors stays on the library's stream)
            wt = torch.as_tensor(w_full, device='cuda:%d' % ctx.device)
            src = dist.get_global_rank(group, root) if group is not None else root
            dist.broadcast(mat_as_tensor(ctx, U), src=src, group=group)
            dist.broadcast(wt, src=src, group=group)
            w = wt.cpu().numpy()
        if rows is None:
            out.append(EigenDict(values=w, vectors=LazyHostArray(U)))
        else:
            out.append(EigenDict(values=w[q:], vectors=LazyHostArray(U, rows=rows), _q=q))
    return out[0], out[1]


# ----------------------------------------------------------------------------------------------------------
# scan
# ----------------------------------------------------------------------------------------------------------
def quad_form_sharded(ctx, R, group=None):
    """A = R'R as packed 256 x 256 lower-triangular blocks ([slots x 65536] DeviceMatrix), every rank forming an equal range of
    the blocks on the int8 tensor pipe and one in-place all-gather completing it.  Returns (A, entry-wise error bound)."""
    import torch.distributed as dist
    from ._lib import DeviceMatrix
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    slots = ctx.quad_form_slots(R.shape[1])
    begin, count, per = slot_range(slots, rank, world)
    A = DeviceMatrix(ctx, per * world, 65536, zero=False)
    err = ctx.quad_form_tiles(R, begin, count, A)
    t = mat_as_tensor(ctx, A)
    with on_lib_stream(ctx, 'allgather'):
        dist.all_gather_into_tensor(t, t[rank * per:(rank + 1) * per], group=group)
    return A, err


def check_same_model(ctx, values, group=None):
    """Raises unless `values` (a few scalars that identify the model: delta, h0_rss, ...) agree on every rank -- the sharded
    scan sums pieces of ONE rotation, ranks fitting different models would silently mix them."""
    import torch
    import torch.distributed as dist
    with on_lib_stream(ctx, 'allreduce'):
        v = torch.as_tensor(np.asarray(values, dtype=np.float64), device='cuda:%d' % ctx.device)
        lohi = torch.cat([v, -v])                               # one MIN all-reduce: min(v) and -max(v)
        dist.all_reduce(lohi, op=dist.ReduceOp.MIN, group=group)
        lohi = lohi.cpu().numpy()
    lo, hi = lohi[:len(lohi) // 2], -lohi[len(lohi) // 2:]
    if not np.allclose(lo, hi, rtol=1e-9, atol=0.0):
        raise ValueError('sharded EMMAX scan: the ranks of the process group are not fitting the same model '
                         '(min %r != max %r over ranks); every rank must pass the same phenotype, kinship and cofactors' % (lo, hi))


class GatheredRows(object):
    """Per-SNP result vector of a sharded scan: a row of the gathered, compacted device matrix ([5 x m_total]) that turns
    into a numpy array on first use (np.asarray(x), x[...], len(x))."""

    def __init__(self, gathered, key_index, m_total):
        self._g, self._k, self._m = gathered, key_index, m_total
        self._host = None

    def host(self):
        if self._host is None:
            self._host = self._g.download_rows(self._k, 1, 1)[0]      # one contiguous row: no host-side concatenation
            self._g = None
        return self._host

    @property
    def shape(self):
        return (self._m,)

    def __len__(self):
        return self._m

    def __array__(self, dtype=None, copy=None):
        a = self.host()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, k):
        return self.host()[k]

    def __setitem__(self, k, v):
        self.host()[k] = v


def scan_sharded(ctx, A, a_err, v, h0_rss, n_p, m_total=None, group=None, eager=None):
    """The int8 scan of this rank's resident SNP slice given the packed quadratic form, followed by the all-gather of the
    per-rank outputs on the device.  Returns {'ps', 'f_stats', 'rss', 'var_perc'} covering ALL m_total SNPs (rank
    order = shard_range order): `ps` as a numpy array on every rank; the other four as numpy arrays on rank 0 and as
    GatheredRows (downloaded on first use) elsewhere, unless `eager` says otherwise."""
    import torch
    import torch.distributed as dist
    from ._lib import DeviceMatrix
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    m_loc, n = ctx.snps_shape()
    if m_total is not None:
        sizes = [shard_range(m_total, r, world) for r in range(world)]
        if sizes[rank][1] - sizes[rank][0] != m_loc:
            raise ValueError('rank %d holds %d SNPs, shard_range(%d) says %d' % (rank, m_loc, m_total, sizes[rank][1] - sizes[rank][0]))
    else:
        with on_lib_stream(ctx, 'allreduce'):
            cnt = torch.zeros(world, dtype=torch.int64, device='cuda:%d' % ctx.device)
            cnt[rank] = m_loc
            dist.all_reduce(cnt, group=group)
            cnt = cnt.cpu().numpy()
        ends = np.cumsum(cnt)
        sizes = [(int(e - c), int(e)) for c, e in zip(cnt, ends)]
    m_all = sizes[-1][1]
    maxlen = max(e - b for b, e in sizes)
    out = DeviceMatrix(ctx, len(RESULT_KEYS), maxlen)
    ctx.emmax_scan_quad_dev(A, v, h0_rss, n_p, packed=True, a_err=a_err, out=out)
    nk = len(GATHERED_KEYS)                                     # x~.x~ (the kernel's fifth row) is not part of the result dict: it stays behind
    gathered = DeviceMatrix(ctx, world * nk, maxlen, zero=False)
    final = DeviceMatrix(ctx, nk, m_all, zero=False)
    with on_lib_stream(ctx, 'allgather'):
        tg = mat_as_tensor(ctx, gathered)
        dist.all_gather_into_tensor(tg, mat_as_tensor(ctx, out)[:nk], group=group)
        # compact the padded per-rank blocks into [5 x m_total] on the device: the host then receives finished vectors
        tg = tg.view(world, nk, maxlen)
        torch.cat([tg[r, :, :e - b] for r, (b, e) in enumerate(sizes)], dim=1, out=mat_as_tensor(ctx, final))
    out.free()
    gathered.free()
    if eager is None:
        eager = rank == 0
    if eager:
        host = final.download(pinned=True)                      # one copy; the five vectors are its rows
        final.free()
        return {name: host[k] for k, name in enumerate(GATHERED_KEYS)}
    res = {}
    for k, name in enumerate(GATHERED_KEYS):
        g = GatheredRows(final, k, m_all)
        res[name] = g.host() if name == 'ps' else g
    return res


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
