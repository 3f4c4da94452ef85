"""Per-launch device time (CUDA events, library profiling mode) of every level operation on the
finest level of a workload, plus whole-cycle times through CUDA graphs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from afivo_streamer_b200 import mg as M  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "S1"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
tree, bc, ids, rhs, desc = bench.build_workload(name, want_rhs=(name != "S3"))
mg = M.mg_t(sides_bc=bc)
M.mg_init(tree, mg)
if ids is not None:
    mg.set_cc(M.I_RHS, ids, rhs)
M.mg_fas_fmg(tree, mg, True, False)
mg.fas_vcycle_async(True, 0, 3)
mg.sync()
mg.fas_vcycle_async(True, 0, 20)
mg.sync()
print(f"{name}: V-cycle {mg.last_cycle_ms() / 20:.4f} ms  ({mg.cell_updates() * 20 / mg.last_cycle_ms() / 1e6:.1f} G cell-updates/s)")
mg.fas_fmg_async(False, True, 5)
mg.sync()
print(f"{name}: FMG {mg.last_cycle_ms() / 5:.4f} ms")
L = tree.highest_lvl
prev = 0.0
for lv in range(1, L + 1):
    mg.fas_vcycle_async(False, lv, 3)
    mg.sync()
    mg.fas_vcycle_async(False, lv, 20)
    mg.sync()
    cur = mg.last_cycle_ms() / 20
    print(f"  V-cycle(highest_lvl={lv}) {1e3 * cur:9.1f} us   (+{1e3 * (cur - prev):.1f} us for this level)")
    prev = cur
if tree.ndim == 2:
    M.mg_destroy(mg)
    sys.exit(0)
mg.set_profiling(True)
for _ in range(reps):
    for lv in range(L, 1, -1):
        mg.gsrb_halfsweep(lv, 1)
        mg.gsrb_halfsweep(lv, 2)
    mg.update_coarse(L, True)
    mg.correct_children_gc(L - 1)
    mg.gc_lvl(L, M.I_PHI, True)
    mg.residual_lvl(L)
    mg.solve_coarse_grid()
for k, (ms, calls) in sorted(mg.profile().items()):
    print(f"  {k:20s} {1e3 * ms / calls:9.2f} us/launch  x{calls}")
mg.set_profiling(False)
M.field_from_potential(tree, mg, -1.0)
mg.set_profiling(True)
for _ in range(reps):
    M.field_from_potential(tree, mg, -1.0)
for k, (ms, calls) in sorted(mg.profile().items()):
    if k.startswith("field"):
        print(f"  {k:20s} {1e3 * ms / calls:9.2f} us/launch  x{calls}")
M.mg_destroy(mg)
