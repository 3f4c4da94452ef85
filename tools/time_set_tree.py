"""What a refinement costs on the library side: host time of afmg_set_tree (topology up, slot maps, slab), afmg_set_bc
and the first cycle after it (coarse-solver set-up + graph capture), on a bench.py workload:

    python tools/time_set_tree.py [--workload S3]

The second afmg_set_tree on the same handle is the case of a time loop (af_adjust_refinement changed the tree)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="S3")
    args = ap.parse_args()
    import bench as B
    from afivo_streamer_b200 import _lib
    from afivo_streamer_b200 import mg as M

    tree, bc, _, _, desc = B.build_workload(args.workload, want_rhs=False)
    mg = M.mg_t(sides_bc=bc, device=0, lsf_boundary_value=1.0)
    t0 = time.perf_counter()
    M.mg_init(tree, mg)
    t_init = time.perf_counter() - t0
    out = {"workload": args.workload, "n_boxes": int(tree.n_boxes), "mg_init_first_s": t_init}
    for rep in ("second", "third"):
        t0 = time.perf_counter()
        td, keep = M._tree_desc(tree)
        t1 = time.perf_counter()
        mg._check(_lib.lib().afmg_set_tree(mg._h, C.byref(td)))
        t2 = time.perf_counter()
        mg.set_bc(bc)
        t3 = time.perf_counter()
        M.mg_fas_fmg(tree, mg, True, False)
        t4 = time.perf_counter()
        M.mg_fas_fmg(tree, mg, True, True)
        t5 = time.perf_counter()
        out[rep] = {"python_tree_desc_s": t1 - t0, "afmg_set_tree_s": t2 - t1, "afmg_set_bc_s": t3 - t2,
                    "first_fmg_s (coarse set-up + graph capture + run)": t4 - t3, "next_fmg_s": t5 - t4}
    print(json.dumps(out))
    M.mg_destroy(mg)


if __name__ == "__main__":
    main()
