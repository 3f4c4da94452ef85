"""Run each level operation of the path once or twice on the finest level of a workload, for ncu.

    ncu --set full --clock-control none --import-source on -k regex:k_ -o gpurun_out/prof python tools/prof_ops.py S1
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from afivo_streamer_b200 import mg as M  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "S1"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
tree, bc, ids, rhs, desc = bench.build_workload(name, want_rhs=False)  # data values do not change the traffic
mg = M.mg_t(sides_bc=bc)
M.mg_init(tree, mg)
mg.set_profiling(True)  # no graphs: plain launches
L = tree.highest_lvl
for _ in range(reps):
    mg.gsrb_halfsweep(L, 1)
    mg.gsrb_halfsweep(L, 2)
    mg.update_coarse(L, True)
    mg.correct_children_gc(L - 1)
    mg.residual_lvl(L)
print("done", M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
M.mg_destroy(mg)
