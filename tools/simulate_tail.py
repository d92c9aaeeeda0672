"""List-scheduling simulation of the strong-scaling tail (CPU only; uses the oracle for the iteration counts -- a measurement aid, like tests/).

One 8192-QP batch (config 3, S1) split over N GPUs with 296 resident CTAs each: a CTA slot draws the next QP of its GPU's slice when it finishes
one. Cost of a QP = its ADMM iterations (+ a setup equivalent of ~40 iterations). Reports the efficiency of the makespan against total/slots for
(a) the production order (contiguous slices, queue in batch order), (b) longest-first inside each slice (an oracle-given upper bound on what any
iteration-count predictor could achieve), (c) time slices of q iterations (a QP re-enters the queue after q iterations).
"""
import heapq
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def makespan(costs, slots):
    h = [0.0] * slots
    heapq.heapify(h)
    for c in costs:
        t = heapq.heappop(h)
        heapq.heappush(h, t + c)
    return max(h)


def main():
    from oracle import qp_oracle as O
    from sqp_solver_b200.synth import make_batch

    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    d = make_batch(B, 64, 128, seed0=0)
    r = O.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], O.default_settings())
    it = np.minimum(r["iter"], 1000).astype(float) + 40.0
    print("QPs %d, iterations mean %.0f min %.0f max %.0f" % (B, it.mean() - 40, it.min() - 40, it.max() - 40))
    for N in (1, 2, 4, 8):
        slots = 296
        per = B // N
        ideal = it.sum() / (N * slots)
        res = {}
        res["batch order"] = max(makespan(it[g * per:(g + 1) * per], slots) for g in range(N))
        res["longest first"] = max(makespan(np.sort(it[g * per:(g + 1) * per])[::-1], slots) for g in range(N))
        for q in (100, 250):
            sl = []
            for g in range(N):
                c = it[g * per:(g + 1) * per]
                pieces = []
                rem = c.copy()
                while (rem > 0).any():  # round-robin re-queueing: every QP contributes one slice per round (+6 iterations' worth of reload)
                    pieces.extend(np.minimum(rem[rem > 0], q) + 6.0)
                    rem = rem - q
                sl.append(makespan(pieces, slots))
            res["time slices of %d" % q] = max(sl)
        print("N=%d ideal %.0f: " % (N, ideal) + "; ".join("%s %.3f" % (k, ideal / v) for k, v in res.items()))


if __name__ == "__main__":
    main()
