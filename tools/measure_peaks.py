#!/usr/bin/env python
"""Measures the pipe rates the rooflines of bench.py divide by and writes them, with the clocks they were measured at,
to profiles/PEAKS_int8_fp64.json (tracked): FP64 tensor (DMMA) and FP64 FMA issue rates, the tcgen05 int8 issue rate
(single-CTA and CTA-pair MMA; one launch = burst, held 2 s = power-capped steady state), device copy bandwidth."""
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def smi():
    q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_thermal_slowdown'
    out = subprocess.run(['nvidia-smi', '-i', '0', '--query-gpu=' + q, '--format=csv,noheader,nounits'], capture_output=True, text=True).stdout
    return [x.strip() for x in out.strip().split(',')]


def main():
    from mixmogam_b200 import get_context
    ctx = get_context(0)
    res = {'gpu': ctx.device_info()['name'], 'when': time.strftime('%Y-%m-%dT%H:%M:%SZ', time.gmtime()), 'idle': smi()}
    for key, which in (('dmma_tflops', 'dmma'), ('dfma_tflops', 'dfma'), ('copy_gbs', 'copy'), ('int8_single_cta_burst_tops', 'imma_tcgen05'),
                       ('int8_burst_tops', 'imma_pair'), ('mxf4_burst_tops', 'mxf4_tcgen05')):
        res[key] = ctx.microbench(which)
    samples = []
    stop = []

    def sampler():
        while not stop:
            samples.append(smi())
            time.sleep(0.1)
    th = threading.Thread(target=sampler, daemon=True)
    th.start()
    res['int8_sustained_tops'] = ctx.microbench('imma_pair_sustained2000')
    res['int8_single_cta_sustained_tops'] = ctx.microbench('imma_tcgen05_sustained2000')
    res['mxf4_sustained_tops'] = ctx.microbench('mxf4_tcgen05_sustained2000')
    res['int8_digits_sustained_tops'] = ctx.microbench('imma_pair_digits_sustained2000')     # B operand: full-range bytes (digit planes)
    stop.append(1)
    th.join()
    sm = sorted(float(s[0]) for s in samples if s and s[0].replace('.', '').isdigit())
    res['under_load'] = {'sm_mhz_median': sm[len(sm) // 2] if sm else None, 'sm_mhz_min': sm[0] if sm else None, 'samples': len(sm),
                         'power_w_max': max([float(s[2]) for s in samples if len(s) > 2 and s[2].replace('.', '').isdigit()] or [0]),
                         'sw_power_cap_seen': any(len(s) > 3 and s[3].lower().startswith('active') for s in samples)}
    res['how'] = ('libmixmogam_b200_bench.so (mixmogam_b200/csrc/microbench.cu): tcgen05.mma kind::i8 128(256)x256x32 issued back to back '
                  'from shared-memory-resident operands, 148 CTAs (mxf4: kind::mxf4.block_scale 128x256x64 on e2m1 operands, unit scales); burst = second of two 40k-K-block launches, sustained = last '
                  'quarter of ~2 s of back-to-back launches; dmma = mma.sync m8n8k4 f64 chains, 4 blocks/SM')
    json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'PEAKS_int8_fp64.json'), 'w'), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
