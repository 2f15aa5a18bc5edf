"""N-GPU diagnostic (torchrun): device time of the collectives the sharded path issues, on library-owned and torch-owned buffers."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda:%d' % local))
    from mixmogam_b200 import _lib, parallel
    from mixmogam_b200._lib import DeviceMatrix
    ctx = _lib.get_context(local)
    slots = 820
    per = -(-slots // world)

    def timed(name, fn, reps=5):
        for _ in range(2):
            fn()
        ctx.sync()
        torch.cuda.synchronize()
        dist.barrier()
        parallel.collective_timers(ctx, reset=True)
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        ctx.sync()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / reps
        tm = parallel.collective_timers(ctx, reset=True)
        if rank == 0:
            print('%-60s wall %.2f ms  device %s' % (name, wall * 1e3, {k: round(1e3 * v / reps, 2) for k, v in tm.items()}), flush=True)

    A = DeviceMatrix(ctx, per * world, 65536, zero=True)
    t = parallel.mat_as_tensor(ctx, A)

    def ag_inplace():
        with parallel.on_lib_stream(ctx, 'allgather'):
            dist.all_gather_into_tensor(t, t[rank * per:(rank + 1) * per])
    timed('all_gather in place, library buffer (%d MB)' % (t.numel() * 8 >> 20), ag_inplace)

    src = DeviceMatrix(ctx, per, 65536, zero=True)
    ts = parallel.mat_as_tensor(ctx, src)

    def ag_sep():
        with parallel.on_lib_stream(ctx, 'allgather'):
            dist.all_gather_into_tensor(t, ts)
    timed('all_gather separate input, library buffers', ag_sep)

    def ag_fresh():
        B = DeviceMatrix(ctx, per * world, 65536, zero=False)
        tb = parallel.mat_as_tensor(ctx, B)
        with parallel.on_lib_stream(ctx, 'allgather'):
            dist.all_gather_into_tensor(tb, tb[rank * per:(rank + 1) * per])
        B.free()
    timed('all_gather in place, fresh library buffer every call', ag_fresh)

    tt = torch.zeros(per * world, 65536, dtype=torch.float64, device='cuda:%d' % local)

    def ag_torch():
        with parallel.on_lib_stream(ctx, 'allgather'):
            dist.all_gather_into_tensor(tt, tt[rank * per:(rank + 1) * per])
    timed('all_gather in place, torch buffer, library stream', ag_torch)

    def ag_torch_default():
        dist.all_gather_into_tensor(tt, tt[rank * per:(rank + 1) * per])
    timed('all_gather in place, torch buffer, torch stream', ag_torch_default)

    ti = torch.zeros(820 * 65536, dtype=torch.int32, device='cuda:%d' % local)

    def ar():
        with parallel.on_lib_stream(ctx, 'allreduce'):
            dist.all_reduce(ti)
    timed('all_reduce int32 (%d MB), torch buffer, library stream' % (ti.numel() * 4 >> 20), ar)

    small = DeviceMatrix(ctx, 5, 62592)
    big = DeviceMatrix(ctx, 5 * world, 62592, zero=False)

    def ag_small():
        with parallel.on_lib_stream(ctx, 'allgather'):
            dist.all_gather_into_tensor(parallel.mat_as_tensor(ctx, big), parallel.mat_as_tensor(ctx, small))
    timed('all_gather of the per-SNP results (5 x 62592 doubles per rank)', ag_small)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
