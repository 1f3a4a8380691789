"""Summarise the per-step timeline dumps written with LIC360_WF_TRACE=1 LIC360_WF_TRACE_FILE=<csv> (codec.cu): window averages of the
device stamps (us after the step's first stamp) and of the host's per-step times."""
import csv, sys
import numpy as np
for fn in sys.argv[1:]:
    rows = list(csv.DictReader(open(fn)))
    A = {k: np.array([float(r[k]) for r in rows]) for k in rows[0]}
    print(fn, 'sum period %.2f ms, host wait %.2f, host decode %.2f, levels wait %.2f' % (
        A['period'].sum() / 1e3, A['host_wait_us'].sum() / 1e3, A['host_decode_us'].sum() / 1e3, A['levels_wait_us'].sum() / 1e3))
    print(' steps     len period   wait    dec levels | chain0  chain  rows0  rows1  tail0  tail1 old_end | rows0+dec')
    n = len(rows)
    for a in range(0, n, 20):
        s = slice(a, min(a + 20, n))
        f = lambda k: A[k][s].mean()
        print(' %3d-%3d %5.0f %6.1f %6.1f %6.1f %6.1f | %6.1f %6.1f %6.1f %6.1f %6.1f %6.1f %7.1f | %6.1f' % (
            a, min(a + 20, n), f('len'), f('period'), f('host_wait_us'), f('host_decode_us'), f('levels_wait_us'), f('chain_start'),
            f('chain_end') - f('chain_start'), f('rows_start'), f('rows_last'), f('tail_start'), f('tail_end'), f('old_end'),
            f('rows_start') + f('host_decode_us')))
