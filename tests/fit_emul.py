"""TEST INFRASTRUCTURE — a host stand-in for one `pbn_fit_scan_host` call, so that the fitter's host orchestration
(tie settling, last-slot rule, assembly) can be exercised without a GPU.  It enumerates the candidates in NumPy and
scores them with the library's own per-candidate solver run on the host (`pbn_fit_eval_host`, the same
__host__ __device__ function the kernel calls).  The -m gpu tests run the real kernel against the oracle."""
import itertools

import numpy as np

from gym_PBN.b200 import abi
from gym_PBN.envs.bittner.gen import predictor_sets as ps


def scan(table, rank, top_l, key_gt=None, arr_lt=None, tie_le=None, tie_cap=1 << 16):
    lib = abi.lib()
    G, S = len(table.genes), table.n_samples
    keys_out, ties_out = [], []
    for g in range(G):
        rem = [i for i in range(G) if i != g]
        arrs, masks, yrows = [], [], []
        for a, b, c in itertools.combinations(range(len(rem)), 3):
            ra, rb, rc = (list(table.rows_of(rem[i])) for i in (a, b, c))
            for y in table.rows_of(g):
                for sc, (ia, ib, ic) in enumerate(itertools.product(ra, rb, rc)):
                    arrs.append((a << 40) | (b << 28) | (c << 16) | ((int(y) - int(table.row_off[g])) << 12) | sc)
                    masks.append((table.masks[ia], table.masks[ib], table.masks[ic], table.masks[y]))
                    yrows.append(y)
        if not arrs:
            keys_out.append([])
            continue
        arrs = np.array(arrs, np.uint64)
        m = np.ascontiguousarray(np.array(masks, np.uint32))
        yrows = np.array(yrows)
        lo, hi = np.zeros(len(arrs), np.int32), np.zeros(len(arrs), np.int32)
        abi.check(lib.pbn_fit_eval_host(m.ctypes.data, len(arrs), S, lo.ctypes.data, hi.ctypes.data))
        r_lo = rank[yrows, np.minimum(lo, S)].astype(np.uint64)
        r_hi = rank[yrows, np.minimum(hi, S)].astype(np.uint64)
        ok = np.ones(len(arrs), bool) if arr_lt is None else arrs < arr_lt[g]
        tie = r_lo != r_hi
        key = (r_lo << np.uint64(ps.ARR_BITS)) | arrs
        sel = ok & ~tie
        if key_gt is not None:
            sel &= key > key_gt[g]
        keys_out.append([int(k) for k in np.sort(key[sel])[:top_l]])
        if tie_le is not None:
            for k in key[ok & tie & (r_lo <= tie_le[g])]:
                ties_out.append((int(k), g))
    return keys_out, ties_out, 0.0
