#!/usr/bin/env python
"""
bench.py -- EMMAX SNP-tests/s on synthetic genotypes (BASELINE.json configs[1]: n=10k individuals x 1M SNPs,
single phenotype, IBS kinship + EMMAX).

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA), one process per GPU under torchrun
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port), host cores

A step is one pass of the hot path over the whole workload: IBS kinship (int8 tensor-core Gram + FP64
finalisation) -> REML delta grid -> SNP scan (rotation + per-SNP OLS + F + p).  The eigendecompositions of
K and S(K+I)S run once per K through cuSOLVER, outside the hot path (north_star), are timed separately and
reported as `eigh_seconds`; each timed step re-uses them through the reference API's own eig_L / eig_R
arguments (linear_models.py:1233).

  value : SNP-tests/s, whole job, genotypes already resident in HBM, max-over-ranks time.
  e2e   : the same metric through the public Python API with HOST buffers: the pinned-host -> device copy of
          the genotypes and of K, and the device -> host copies of K and of the per-SNP results are inside
          the timed region.
With N > 1 the 1M SNPs are sharded across ranks (strong scaling): per-rank partial int32 Gram ->
NCCL all-reduce -> replicated REML -> per-rank scan -> all-gather of the p-values.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 20240601 + 1          # seed 20240601 + config index
GEN_CHUNK = 32768            # SNP rows generated per torch call


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--indivs', dest='n', type=int, default=int(os.environ.get('MMG_BENCH_N', 10000)))
    ap.add_argument('--snps', dest='m', type=int, default=int(os.environ.get('MMG_BENCH_M', 1000000)))
    ap.add_argument('--scan-impl', default=os.environ.get('MMG_BENCH_SCAN_IMPL', 'tcgen05'), choices=['tcgen05', 'dmma'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--profile-host', default=None, help='write a cProfile of one resident step and one e2e step to this file')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
# synthetic genotypes: x ~ Binomial(2, f_s), f_s ~ U(0.11, 0.5) (MAF > 0.1 by construction), SNP-major int8
# ------------------------------------------------------------------------------------------------------
def gen_chunk_torch(chunk_id, rows, n, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(SEED * 1000003 + chunk_id)
    f = 0.11 + 0.39 * torch.rand((rows, 1), generator=g, device=device)
    a = (torch.rand((rows, n), generator=g, device=device) < f).to(torch.int8)
    b = (torch.rand((rows, n), generator=g, device=device) < f).to(torch.int8)
    return a + b


def gen_genotypes_pinned(begin, end, n, device):
    """Rows [begin, end) of the global synthetic genotype matrix, as a numpy array over pinned host memory."""
    import torch
    from mixmogam_b200 import _lib
    host = _lib.pinned_empty((end - begin, n), np.int8)
    th = torch.from_numpy(host)
    c0, c1 = begin // GEN_CHUNK, (end - 1) // GEN_CHUNK
    for c in range(c0, c1 + 1):
        x = gen_chunk_torch(c, GEN_CHUNK, n, device)
        lo, hi = max(begin, c * GEN_CHUNK), min(end, (c + 1) * GEN_CHUNK)
        th[lo - begin:hi - begin].copy_(x[lo - c * GEN_CHUNK:hi - c * GEN_CHUNK])
        del x
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return host


def gen_phenotype(n, device):
    """y = X'beta + e from 10 causal SNPs of chunk 0 (identical on every rank), standardised."""
    import torch
    x = gen_chunk_torch(0, GEN_CHUNK, n, device)[:10].double().cpu().numpy()
    rng = np.random.Generator(np.random.PCG64(SEED))
    y = rng.normal(0, 0.5, size=10) @ x + rng.normal(size=n)
    return (y - y.mean()) / y.std()


def gen_genotypes_numpy(m, n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    f = rng.uniform(0.11, 0.5, size=(m, 1)).astype(np.float32)
    return ((rng.random((m, n), dtype=np.float32) < f).astype(np.int8) + (rng.random((m, n), dtype=np.float32) < f).astype(np.int8))


# ------------------------------------------------------------------------------------------------------
# clocks (nvidia-smi sampled during the timed region)
# ------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop = False
        self.t = None

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = [float(s[0]) for s in self.samples if s[0].replace('.', '').isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({nm for s in self.samples for nm, v in zip(names, s[3:7]) if v.lower().startswith('active')})
        pw = [float(s[2]) for s in self.samples if s[2].replace('.', '').isdigit()]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'power_w_max': max(pw) if pw else None, 'samples': len(self.samples)}


# ------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle (line-faithful float32 port of the reference) on a bounded sample
# ------------------------------------------------------------------------------------------------------
def cpu_baseline(n, m, budget_s=20.0):
    import warnings
    from scipy import linalg, stats
    from oracle import reference_py3 as o
    warnings.simplefilter('ignore')
    cores = os.cpu_count() or 1
    # --- kinship sample: 'diploid_int' through the vectorised integer identity (the literal loop of
    #     kinship.py:33-41 is O(n^2) Python calls per chunk: days at n=10k) ---
    mk = max(256, min(m, 4096))
    snps = gen_genotypes_numpy(mk, n, SEED)
    t0 = time.perf_counter()
    o.ibs_counts_diploid_vectorised(snps)
    t_kin = (time.perf_counter() - t0) / mk
    # --- scan sample: the chunk loop of _emmax_f_test_ (linear_models.py:1316-1349): float32 sgemm of the
    #     chunk with M, one scipy lstsq per SNP, f.sf.  M is a random float32 stand-in (timing only: the
    #     eigendecomposition that would produce it is excluded on both arms). ---
    rng = np.random.default_rng(0)
    M = (rng.standard_normal((n, n), dtype=np.float32) / np.float32(np.sqrt(n)))
    Y = rng.standard_normal((n, 1), dtype=np.float32)
    h0_rss = float(Y.T @ Y)

    def scan_chunk(cnt):
        t0 = time.perf_counter()
        Xs = snps[:cnt].astype(np.float32) @ M                                  # :1317-1318
        rss_list = np.repeat(np.float32(h0_rss), cnt)
        for j in range(cnt):
            (betas, rss, rk, sigma) = linalg.lstsq(Xs[j:j + 1].T, Y)            # :1328
            if rss.size and rss[0] != 0:
                rss_list[j] = rss[0]
        rss_ratio = h0_rss / rss_list
        f_stats = (rss_ratio - 1) * (n - 2)
        stats.f.sf(f_stats, 1, n - 2)                                           # :1349
        return time.perf_counter() - t0

    scan_chunk(32)
    t_probe = scan_chunk(128) / 128
    cnt = int(max(256, min(mk, (budget_s * 0.6) / max(t_probe, 1e-6))))
    t_scan = scan_chunk(cnt) / cnt
    per_snp = t_kin + t_scan
    return {'value': 1.0 / per_snp, 'unit': 'SNP-tests/s', 'cores': cores, 'kind': 'port',
            'sample': 'kinship: %d SNPs (vectorised integer identity for diploid_int, f32 sgemm), %.1f us/SNP; '
                      'scan: chunk loop (f32 sgemm + per-SNP scipy lstsq + f.sf, linear_models.py:1316-1349) on %d SNPs '
                      'with a random stand-in for M, %.1f us/SNP; n=%d; extrapolated linearly in m; eigh excluded on both arms'
                      % (mk, t_kin * 1e6, cnt, t_scan * 1e6, n),
            'kinship_us_per_snp': t_kin * 1e6, 'scan_us_per_snp': t_scan * 1e6}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    base = {'metric': 'EMMAX SNP-tests/sec (kinship + REML + scan; eigendecomposition excluded)', 'unit': 'SNP-tests/s',
            'impl': 'reference', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'configs[1]: synthetic n=%d individuals x %d SNPs, single phenotype, IBS kinship (diploid_int) + EMMAX'
                                   % (args.n, args.m), 'n': args.n, 'm': args.m}}
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        last = cpu_baseline(args.n, args.m, budget_s=10.0 if i < args.warmup else 20.0)
        if i >= args.warmup:
            vals.append(last['value'])
    v = float(np.mean(vals))
    last['value'] = v
    base.update({'value': v, 'ms_per_step': 1e3 * args.m / v, 'cpu_baseline': last,
                 'e2e': {'value': v, 'unit': 'SNP-tests/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                 'gpu_launches': 0})
    print(json.dumps(base))


# ------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    # rank 0 prints exactly ONE line on stdout (the JSON).  NCCL writes its version banner straight to file descriptor 1
    # from C, so the descriptor itself is pointed at stderr for the whole run and the JSON goes out through a saved copy.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=torch.device('cuda:%d' % local))
    device = torch.device('cuda:%d' % local)

    from mixmogam_b200 import _lib, kinship, linear_models as lm, parallel
    ctx = _lib.get_context(local)
    n, m = args.n, args.m
    b, e = parallel.shard_range(m, rank, world)
    m_loc = e - b

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ctx.sync()

    # ---- inputs ----
    t0 = time.perf_counter()
    snps = gen_genotypes_pinned(b, e, n, device)
    y = gen_phenotype(n, device)
    gen_s = time.perf_counter() - t0

    # ---- set-up outside the hot path: K once, its two eigendecompositions (timed separately) ----
    fp64_peak = ctx.microbench('dmma') if rank == 0 else 0.0      # FP64 tensor (DMMA) issue rate, GPU still cool
    imma_peak = ctx.microbench('imma_tcgen05') if rank == 0 else 0.0   # tcgen05 int8 issue rate, smem-resident operands
    ctx.ensure_snps(snps)
    Kd = parallel.calc_ibs_kinship_sharded(snps, m, 'diploid_int', ctx=ctx)
    lmm = lm.LinearMixedModel(y, ctx=ctx, scan_impl=args.scan_impl)
    lmm.add_random_effect(Kd)
    ctx.timer_reset()
    t0 = time.perf_counter()
    eig_L = lmm._get_eigen_L_()
    eig_R = lmm._get_eigen_R_(X=lmm.X)
    eigh_wall = time.perf_counter() - t0
    eigh_dev = ctx.timer('syevd')[0]
    Kd.free()

    # ---- one step with genotypes resident (value) ----
    def step_resident():
        K = parallel.calc_ibs_kinship_sharded(snps, m, 'diploid_int', ctx=ctx)      # pack + Gram (+ all-reduce) + finalize
        mdl = lm.LinearMixedModel(y, ctx=ctx, scan_impl=args.scan_impl)
        mdl.add_random_effect(K)
        K.free()
        r = mdl.emmax_f_test(snps, eig_L=eig_L, eig_R=eig_R, emma_num=0)            # REML + scan
        ps = parallel.allgather_rows(r['ps'], m, device=local)
        return r, ps

    # ---- one step through the public API with host buffers (e2e) ----
    def step_e2e():
        ctx.invalidate_snps()                                                        # force the H2D copy every step
        if world == 1:
            K = kinship.calc_ibs_kinship(snps, 'diploid_int')                        # H2D snps, Gram, D2H K
            r = lm.LinearMixedModel(y, ctx=ctx, scan_impl=args.scan_impl)
            r.add_random_effect(K)                                                   # H2D K
            res = r.emmax_f_test(snps, eig_L=eig_L, eig_R=eig_R, emma_num=0)         # D2H ps, f, rss, var_perc, xx
            return res, res['ps']
        return step_resident()

    for _ in range(args.warmup):
        res, ps = step_resident()            # results held across steps exactly as in the timed loop (same buffer-pool pattern)

    if args.profile_host and rank == 0:
        import cProfile
        import io
        import pstats
        buf = io.StringIO()
        for name, fn in (('resident', step_resident), ('e2e', step_e2e)):
            fn()
            ctx.timer_reset()
            pr = cProfile.Profile()
            t0 = time.perf_counter()
            pr.enable()
            fn()
            barrier()
            pr.disable()
            buf.write('==== %s step: %.1f ms wall; stage timers (ms): %s\n' % (
                name, 1e3 * (time.perf_counter() - t0), {k: round(1e3 * v, 2) for k, v in ctx.timers().items()}))
            pstats.Stats(pr, stream=buf).sort_stats('cumulative').print_stats(45)
        open(args.profile_host, 'w').write(buf.getvalue())

    gram_ms, scan_ms = [], []
    ctx.timer_reset()
    l0 = ctx.launch_count()
    barrier()
    profiling = bool(os.environ.get('MMG_PROFILE_RANGE'))
    if profiling:
        torch.cuda.profiler.start()          # ncu --profile-from-start off: capture the timed steps only
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res, ps = step_resident()
            gram_ms.append(ctx.last_kernel_ms('gram'))
            scan_ms.append(ctx.last_kernel_ms('scan'))
        barrier()
        t_res = time.perf_counter() - t0
    if profiling:
        torch.cuda.profiler.stop()
    launches = ctx.launch_count() - l0
    timers = ctx.timers()

    e2e = None
    e2e_timers = {}
    if not args.no_e2e:
        for _ in range(2):
            res_e, ps_e = step_e2e()
        barrier()
        ctx.timer_reset()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res_e, ps_e = step_e2e()
        barrier()
        t_e2e = time.perf_counter() - t0
        e2e_timers = ctx.timers()
    if world > 1:
        tt = torch.tensor([t_res, t_e2e if not args.no_e2e else 0.0], dtype=torch.float64, device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_res, t_e2e_max = float(tt[0]), float(tt[1])
    else:
        t_e2e_max = t_e2e if not args.no_e2e else 0.0
    if not args.no_e2e:
        # K is downloaded to the caller (read-only, pinned) but its device copy is kept and found again by
        # add_random_effect: no K upload.  Uploads: genotypes + X|Y columns for the two null fits.
        h2d = m_loc * n + n * 8 * 2 * 2
        d2h = (n * n * 8 if world == 1 else 0) + 5 * m_loc * 8
        e2e = {'value': m * args.steps / t_e2e_max, 'unit': 'SNP-tests/s', 'h2d_bytes_per_step': int(h2d),
               'd2h_bytes_per_step': int(d2h), 'ms_per_step': 1e3 * t_e2e_max / args.steps,
               'stage_seconds_per_step': {k: v / args.steps for k, v in e2e_timers.items()},
               'h2d_lanes': dict(zip(('chunks_packed_2bit', 'chunks_raw', 'host_pack_gbs'), ctx.last_h2d_info()))}
        # h2d_bytes_per_step is what the API is handed (the int8 genotype buffer); the packed lane puts a quarter of its
        # chunks' bytes on the PCIe link, so the bytes that actually cross it are fewer
        pk, rw, _ = ctx.last_h2d_info()
        if pk + rw > 0:
            e2e['h2d_pcie_bytes_per_step'] = int(m_loc * n * (rw + 0.25 * pk) / (pk + rw) + n * 8 * 2 * 2)

    if rank == 0:
        value = m * args.steps / t_res
        # ---- roofline of the dominant kernel (the scan), measured live with CUDA events on its stream ----
        scan_s = float(np.mean(scan_ms)) * 1e-3
        alg_flops = 2.0 * n * n * m_loc                     # 2 n^2 FP64 flops per SNP (rotation)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        if args.scan_impl == 'tcgen05':
            S, rho = ctx.last_scan_info()                               # planes chosen by the certified-bound rule, its bound
            npad = (n + 255) // 256 * 256
            nt = npad // 256
            kblocks = sum(min((n + 127) // 128, 2 * (jb + 1)) for jb in range(nt))
            exec_ops = 2.0 * m_loc * 256 * 128 * kblocks * S           # int8 MAC*2 actually issued (lower-triangular K ranges)
            # int8 tcgen05 rate = 2 x bf16 (same pipe, K = 32 vs 16 per instruction).  The BURST figure: twice the sustained one
            # (2792) is below what this kernel executes (2.9-3.1 POP/s), so it is not a ceiling for the int8 pipe; the stricter
            # denominator, the int8 issue rate measured in this very run, is reported beside it (frac_of_issue_rate)
            bf16 = peaks.get('bf16_tflops') or peaks.get('bf16_tflops_sustained') or 1590.0
            int8_peak = 2.0 * bf16
            # dram__bytes_read + dram__bytes_write of this kernel from the committed ncu capture of this very configuration
            # (profiles/r01_ncu_full_scan_quad_1m.txt: CTA-pair schedule, 4 planes); null for any other shape
            traffic = 66381656000 + 170236160 if (n, m, world, S) == (10000, 1000000, 1, 4) and not os.environ.get('MMG_SCAN_SCHED') else None
            roof = {'bound': 'tensor', 'kernel': 'tc_gemm_i8_kernel<QuadEpi>' if os.environ.get('MMG_SCAN_SCHED') == 'table' else 'scan_quad_kernel', 'achieved': exec_ops / scan_s / 1e12,
                    'peak': int8_peak, 'unit': 'TFLOP/s', 'frac': exec_ops / scan_s / 1e12 / int8_peak, 'traffic': traffic,
                    'algorithmic_bytes': float(m_loc) * n + 8.0 * m_loc, 'int8_issue_rate_measured': imma_peak,
                    'frac_of_issue_rate': exec_ops / scan_s / 1e12 / imma_peak if imma_peak else None,
                    'pipe': 'int8 tcgen05 (TOP/s); peak = 2 x bf16 burst of %s' % ('MEASURED_PEAKS.json' if peaks else 'the fallback'),
                    'algorithmic_fp64_tflops': alg_flops / scan_s / 1e12, 'fp64_tensor_peak_measured': fp64_peak,
                    'slices': S, 'certified_rel_bound_xx': rho, 'launch_ms': scan_s * 1e3,
                    'achieved_is': 'int8 ops the kernel executes (S planes x lower-triangular K ranges); the SURVEY 8d figure, '
                                   '2 n^2 FP64 flops per SNP of the rotation it replaces, is algorithmic_fp64_tflops'}
        else:
            roof = {'bound': 'tensor', 'kernel': 'scan_dmma_kernel', 'achieved': alg_flops / scan_s / 1e12, 'peak': fp64_peak,
                    'unit': 'TFLOP/s', 'frac': alg_flops / scan_s / 1e12 / fp64_peak, 'traffic': None,
                    'pipe': 'FP64 DMMA; peak = DMMA issue-rate microbenchmark measured in this run', 'launch_ms': scan_s * 1e3}
        gram_s = float(np.mean(gram_ms)) * 1e-3
        out = {'metric': 'EMMAX SNP-tests/sec (kinship + REML + scan; eigendecomposition excluded)', 'value': value,
               'unit': 'SNP-tests/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
               'ms_per_step': 1e3 * t_res / args.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
               'dtype': 'f64 (scan: exact int8 slices of the FP64 rotation; kinship: int8/int32)' if args.scan_impl == 'tcgen05' else 'f64',
               'data': 'synthetic',
               'config': {'workload': 'configs[1]: synthetic n=%d individuals x %d SNPs, single phenotype, IBS kinship (diploid_int) + EMMAX'
                                      % (n, m), 'n': n, 'm': m, 'scan_impl': args.scan_impl,
                          'l2': 'inputs (%.1f GB genotypes per rank) exceed the 126 MB L2; no flush needed' % (m_loc * n / 1e9),
                          'eigh': 'outside the timed step (north_star), see eigh_seconds'},
               'roofline': roof, 'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clk.summary(),
               'eigh_seconds': {'wall': eigh_wall, 'syevd_device': eigh_dev, 'count': 2},
               'stage_seconds_per_step': {k: v / args.steps for k, v in timers.items()},
               'kinship': {'gram_ms': gram_s * 1e3, 'int8_tops_algorithmic': 2.0 * n * n * 2 * m_loc / gram_s / 1e12,
                           'note': 'thermometer coding c=2; symmetric kernel executes ~half of the algorithmic ops'},
               'setup_seconds': {'generate': gen_s}}
        if not args.no_cpu_baseline and world == 1:
            out['cpu_baseline'] = cpu_baseline(n, m)
        else:
            out['cpu_baseline'] = None
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + '\n').encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
