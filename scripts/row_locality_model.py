"""CPU model behind DESIGN.md 4 ("index-ordered rows"): how many distinct 128 B lines and distinct 32 B records one
warp-wide gather instruction of the pair kernels touches, for four row layouts, on the C2 cell (bcc Fe 20^3 - 1 + H).

A warp of the G = 4 pair kernels holds 8 consecutive atoms x 4 lanes; in a round of 16 row entries lane t of an atom works,
in its m-th pair, on the entry in storage slot 4 t + m. "contiguous" stores rank k in slot k (round 1's layout), "transposed"
stores rank k in slot row_slot(k) (device_common.cuh), so that the four lanes' m-th pairs are four CONSECUTIVE ranks.
"cell walk" orders a row as the build finds it (3 x 3 rows of x-contiguous cells, index order inside a cell), "by index"
as k_nb_order_rows leaves it. Records are 32 B (pos4), so four consecutive indices share a line.

Usage: python scripts/row_locality_model.py [atom groups sampled = 200]      (numpy + scipy, no GPU)"""
import os
import sys

import numpy as np
from scipy.spatial import cKDTree

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openfly_b200 import inputs  # noqa: E402


def main():
    groups = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    pos, ex, _ = inputs.iron_hydrogen_cell(20, 2.8553)
    L = ex[0]
    pos = pos - L * np.floor(pos / L)
    r_list = (5.3 + 2 * 0.001) * 1.1 ** (1 / 3)  # the dimer's list radius (dimer.cpp:47)
    nb = cKDTree(pos, boxsize=L).query_ball_point(pos, r_list)
    cell = L / int(L // r_list)

    def cell_walk(i, js):
        d = pos[js] - pos[i]
        d -= L * np.round(d / L)
        c = np.floor((pos[i] + d) / cell).astype(int)
        return js[np.lexsort((js, c[:, 0], c[:, 1], c[:, 2]))]

    def by_index(i, js):
        return np.sort(js)

    def count(order, transposed):
        lines = recs = atoms = 0
        for i0 in range(4000, 4000 + 8 * groups, 8):
            rows = [order(i, np.array([j for j in nb[i] if j != i])) for i in range(i0, i0 + 8)]
            for rd in range((max(len(r) for r in rows) + 15) // 16):
                for m in range(4):
                    js = [r[rd * 16 + (4 * m + t if transposed else 4 * t + m)] for r in rows for t in range(4)
                          if rd * 16 + (4 * m + t if transposed else 4 * t + m) < len(r)]
                    lines += len({j // 4 for j in js})
                    recs += len(set(js))
            atoms += 8
        return lines / atoms, recs / atoms

    print(f"C2 cell: {len(pos)} atoms, list radius {r_list:.4f}, mean neighbours {np.mean([len(x) - 1 for x in nb]):.2f}")
    print("per atom and pass:            distinct 128 B lines   distinct 32 B records (of ~58 requested)")
    for name, order, tr in (("cell walk, contiguous", cell_walk, False), ("cell walk, transposed", cell_walk, True),
                            ("by index,  contiguous", by_index, False), ("by index,  transposed", by_index, True)):
        ln, rc = count(order, tr)
        print(f"  {name:26s}  {ln:8.2f}               {rc:8.2f}")


if __name__ == "__main__":
    main()
