"""GPU diagnostic: tcgen05 Gram vs numpy on small shapes, with a description of where mismatches fall."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixmogam_b200 import get_context  # noqa: E402

ctx = get_context(0)
print(ctx.device_info())
rng = np.random.default_rng(0)
for (m, n) in [(128, 128), (128, 256), (512, 128), (512, 300), (4096, 520)]:
    x = (rng.random((m, n)) < 0.4).astype(np.int8)
    s = 2 * x.astype(np.int64) - 1
    ref = s.T @ s
    ctx.invalidate_snps()
    ctx.ensure_snps(x)
    for impl in ('simt', 'tcgen05'):
        ctx.kinship_gram(0, impl=impl)
        G = ctx.kinship_gram_download().astype(np.int64)
        bad = G != ref
        print('m=%d n=%d %-8s mismatches=%d/%d' % (m, n, impl, bad.sum(), bad.size), 'gram_ms=%.3f' % ctx.last_kernel_ms('gram'))
        if bad.any():
            iu = np.triu(bad)
            rows, cols = np.nonzero(iu)
            print('   upper-tri mismatch rows: min %d max %d nuniq %d ; cols: min %d max %d nuniq %d' %
                  (rows.min(), rows.max(), len(set(rows)), cols.min(), cols.max(), len(set(cols))))
            print('   row%8 hist', np.bincount(rows % 8, minlength=8), 'col%32 hist', np.bincount(cols % 32, minlength=32))
            i, j = rows[0], cols[0]
            print('   first (%d,%d): got %d want %d ; G[0,:8]=%s ref[0,:8]=%s' % (i, j, G[i, j], ref[i, j], G[0, :8], ref[0, :8]))
            # is the result a permutation / partial sum?  compare against Gram of each 32-SNP group
            if m >= 128:
                for k in (32, 64, 96, 128):
                    part = s[:k].T @ s[:k]
                    print('   equals Gram of first %d SNPs at (i,j)? %s' % (k, G[i, j] == part[i, j]))
print('microbench dmma TF/s', ctx.microbench('dmma'), 'dfma TF/s', ctx.microbench('dfma'), 'copy GB/s', ctx.microbench('copy'))
