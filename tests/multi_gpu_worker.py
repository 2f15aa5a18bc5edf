"""Worker of tests/test_gpu_multi.py: one process per GPU under torchrun.  Every rank builds the same small problem, keeps
its SNP slice, runs the SNP-sharded kinship + EMMAX (mixmogam_b200.parallel), and rank 0 compares the gathered results with
the single-GPU path on the whole data and with the oracle."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda:%d' % local))
    from mixmogam_b200 import _lib, kinship, linear_models as lm, parallel
    from oracle import reference_py3 as o
    ctx = _lib.get_context(local)
    n, m = int(os.environ.get('MMG_MULTI_N', 700)), int(os.environ.get('MMG_MULTI_M', 40000))
    snps = o.synth_genotypes(m, n, 'diploid_int', seed=9)
    b, e = parallel.shard_range(m, rank, world)
    mine = np.ascontiguousarray(snps[b:e])
    out = {}
    # kinship: partial Grams summed over ranks == the Gram of everything, bit for bit
    Kd = parallel.calc_ibs_kinship_sharded(mine, m, 'diploid_int', scaled=False, ctx=ctx)
    K_unscaled = Kd.download()
    Kd.free()
    Kd = parallel.calc_ibs_kinship_sharded(mine, m, 'diploid_int', ctx=ctx)
    K = Kd.download()
    y = o.synth_phenotype(snps, K, seed=4)
    mdl = lm.LinearMixedModel(y, ctx=ctx, scan_impl='tcgen05', shard={'group': 'world', 'm_total': m})
    mdl.add_random_effect(Kd)
    eig_L, eig_R = parallel.shared_eigen(mdl)
    r = mdl.emmax_f_test(mine, eig_L=eig_L, eig_R=eig_R, emma_num=0)
    ps = np.asarray(r['ps'])
    assert ps.shape == (m,)
    # without m_total the slice sizes are exchanged instead
    mdl2 = lm.LinearMixedModel(y, ctx=ctx, scan_impl='tcgen05', shard={'group': 'world'})
    mdl2.add_random_effect(Kd)
    r2 = mdl2.emmax_f_test(mine, eig_L=eig_L, eig_R=eig_R, emma_num=0)
    assert np.array_equal(np.asarray(r2['ps']), ps)
    # every rank holds the same p-values
    t = torch.as_tensor(ps, device='cuda:%d' % local)
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi)
    # ranks fitting different models are refused
    bad = lm.LinearMixedModel(y * (1.0 + 0.01 * rank), ctx=ctx, scan_impl='tcgen05', shard={'group': 'world', 'm_total': m})
    bad.add_random_effect(Kd)
    try:
        bad.emmax_f_test(mine, eig_L=eig_L, eig_R=eig_R, emma_num=0)
        refused = False
    except ValueError:
        refused = True
    assert refused
    if rank == 0:
        ctx.invalidate_snps()
        assert np.array_equal(K_unscaled, o.calc_ibs_kinship_diploid_fast(snps, scaled=False))
        single = lm.emmax(snps, y, K, scan_impl='tcgen05')
        ref = o.emmax(list(snps), y, K, dtype='double')
        a, b0 = -np.log10(ps), -np.log10(ref['ps'])
        out['err_vs_oracle'] = float(np.max(np.abs(a - b0) / np.maximum(b0, 1e-3)))
        out['err_vs_single'] = float(np.max(np.abs(np.asarray(r['f_stats']) / single['f_stats'] - 1.0)))
        out['keys'] = sorted(r.keys())
        for k in ('rss', 'var_perc'):
            np.testing.assert_allclose(np.asarray(r[k]), single[k], rtol=1e-7)
        out['her'] = [r['pseudo_heritability'], ref['pseudo_heritability']]
        print('MULTI_RESULT ' + json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
