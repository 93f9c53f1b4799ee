#!/usr/bin/env python
"""Host-side model of the fused kernel's producer / MMA / epilogue barrier protocol
(b2cnn_tc_fused.cuh).  Each role advances only when the barrier it would wait on is satisfied;
a run that stops before every step is processed is a deadlock.  Cheap insurance before GPU time.
    python scripts/sim_fused_protocol.py [slack]   # slack = constant in the W-chunk prefetch rule
"""
import sys


def simulate(ntiles, LAG=4, slack=None):
    slack = 15 - LAG if slack is None else slack
    J = 7 * ntiles; nch = (J + 7) // 8
    tiles_issued = w_issued = mma_j = proj_done = epi_j = pfull_sent = 0
    prod = ("tile", 0)
    changed = True
    while changed:
        changed = False
        kind, i = prod
        if kind == "tile" and i < ntiles:
            if i < 2 or epi_j >= 7 * (i - 1):                 # empty[s]: tile i-2 fully consumed
                tiles_issued = i + 1; prod = ("w", i); changed = True
        elif kind == "w":
            if w_issued < nch and 8 * w_issued <= 7 * i + slack:
                if w_issued < 2 or proj_done >= w_issued - 1:  # wempty: chunk w_issued-2 projected
                    w_issued += 1; changed = True
            else:
                prod = ("tile", i + 1); changed = True
        if prod == ("tile", ntiles) and w_issued < nch:
            if w_issued < 2 or proj_done >= w_issued - 1:
                w_issued += 1; changed = True
        if mma_j < J:
            j = mma_j
            if j >= 8 + LAG and (j - LAG) % 8 == 0 and proj_done == (j - LAG) // 8 - 1:
                if w_issued > proj_done and pfull_sent > proj_done:
                    proj_done += 1; changed = True
            elif tiles_issued > j // 7 and (j < 4 or epi_j >= j - 4):   # tempty: block j-4 loaded by the epilogue
                mma_j += 1; changed = True
        elif proj_done < nch and w_issued > proj_done and pfull_sent > proj_done:
            proj_done += 1; changed = True
        if epi_j < J:
            j = epi_j; m = j // 8
            ok = mma_j > j and not (j % 8 == 0 and m >= 2 and proj_done < m - 1)   # tfull; pempty of chunk m-2
            if ok:
                epi_j += 1; changed = True
                if epi_j % 8 == 0 or epi_j == J:
                    pfull_sent = (epi_j + 7) // 8
    return epi_j == J and proj_done == nch


if __name__ == "__main__":
    slack = int(sys.argv[1]) if len(sys.argv) > 1 else None
    bad = [n for n in range(1, 80) if not simulate(n, slack=slack)]
    print("deadlocks at tiles-per-CTA:", bad)
    sys.exit(1 if bad else 0)
