import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as entry
pkg = entry.load_package(); synth = pkg.synth
np.set_printoptions(precision=6, suppress=True, linewidth=200)
sw0 = synth.make_sweep(synth.vlp16(), 0); sw1 = synth.make_sweep(synth.vlp16(), 2)
for iters in (1, 2, 3, 10):
    g = pkg.context(0, n_scans=16, gn_max_iters=iters)
    o = entry.load_oracle().context(0, n_scans=16, gn_max_iters=iters, oracle_threads=8)
    _, e0, s0 = o.extract_features(sw0); _, e1, s1 = o.extract_features(sw1)
    for c in (g, o):
        c.map_set(0, e0); c.map_set(1, s0)
    pg, sg = g.register(e1, s1, solver=0); po, so = o.register(e1, s1, solver=0)
    print(iters, "gpu", pg, sg["n_edge_matched"], sg["n_surf_matched"], sg["outer_iters"], sg["degenerate"], sg["final_cost"])
    print(iters, "orc", po, so["n_edge_matched"], so["n_surf_matched"], so["outer_iters"], so["degenerate"], so["final_cost"])
