"""One field_from_potential on a workload, for ncu:
    ncu --set full --clock-control none --import-source on -k regex:k_grad3 -c 1 -o gpurun_out/grad python tools/prof_field.py S3
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from afivo_streamer_b200 import mg as M  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "S1"
tree, bc, ids, rhs, desc = bench.build_workload(name, want_rhs=False)
mg = M.mg_t(sides_bc=bc)
M.mg_init(tree, mg)
M.field_from_potential(tree, mg, -1.0)
M.field_from_potential(tree, mg, -1.0)
M.mg_destroy(mg)
