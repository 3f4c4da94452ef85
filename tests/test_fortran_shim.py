"""CPU: the Fortran shim (fortran/m_af_multigrid_gpu.f90) cannot be compiled in this image (no Fortran compiler), so
its ISO_C_BINDING half is checked textually against the C header and the library: every bind(c) procedure names an
exported symbol of include/afmg.h with the same number of arguments, every bind(c) derived type lists the members of
the C struct in the same order with matching kinds, and the module exports one entry point per reference
entry point (afivo/src/m_af_multigrid.f90:43, :111, :137, :185, :1188) with the reference's argument order."""
import ctypes as C
import os
import re

from afivo_streamer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fortran_source():
    raw = open(os.path.join(ROOT, "fortran", "m_af_multigrid_gpu.f90")).read()
    text = "\n".join(ln.split("!")[0].rstrip() for ln in raw.splitlines())  # no string in the file contains '!'
    return re.sub(r"&\s*\n\s*&?", " ", text)  # join continuation lines


def c_prototypes():
    hdr = open(os.path.join(ROOT, "include", "afmg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(afmg_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).strip()
        depth, n = 0, 0 if args in ("", "void") else 1
        for ch in args:
            depth += ch == "("
            depth -= ch == ")"
            n += ch == "," and depth == 0
        protos[m.group(1)] = n
    return protos


def test_every_binding_matches_a_header_prototype():
    src = fortran_source()
    protos = c_prototypes()
    L = _lib.lib()
    found = re.findall(r"(?:function|subroutine)\s+(\w+)\s*\(([^)]*)\)\s*bind\(c,\s*name=\"(\w+)\"\)", src, flags=re.I)
    assert len(found) >= 20, len(found)
    for fname, args, cname in found:
        assert cname in protos, f"{cname} is not declared in include/afmg.h"
        assert hasattr(L, cname), f"{cname} is not exported by libafmg.so"
        n = len([a for a in args.split(",") if a.strip()])
        assert n == protos[cname], f"{cname}: {n} arguments in the shim, {protos[cname]} in the header"
        assert fname.lower() == cname.lower()


KIND = {"integer(c_int32_t)": C.c_int32, "integer(c_int64_t)": C.c_int64, "real(c_double)": C.c_double,
        "type(c_ptr)": C.c_void_p, "integer(c_int)": C.c_int}


def fortran_members(src, tname):
    body = re.search(rf"type,\s*bind\(c\)\s*::\s*{tname}\b(.*?)end type", src, flags=re.S | re.I).group(1)
    out = []
    for ln in body.splitlines():
        ln = ln.strip()
        if not ln:
            continue
        kind, names = [p.strip() for p in ln.split("::")]
        for nm in re.findall(r"(\w+)(?:\((\d+)\))?", names):
            out.append((nm[0], KIND[kind.replace(" ", "")], int(nm[1] or 1)))
    return out


def ctypes_members(struct):
    out = []
    for name, ty in struct._fields_:
        if issubclass(ty, C.Array):
            out.append((name, ty._type_, ty._length_))
        elif issubclass(ty, C._Pointer):
            out.append((name, C.c_void_p, 1))
        else:
            out.append((name, ty, 1))
    return out


def test_derived_types_mirror_the_c_structs():
    src = fortran_source()
    for tname, struct in (("afmg_opts", _lib.Opts), ("afmg_stencil_desc", _lib.StencilDesc), ("afmg_tree", _lib.TreeDesc)):
        f, c = fortran_members(src, tname), ctypes_members(struct)
        assert [m[0] for m in f] == [m[0] for m in c], (tname, [m[0] for m in f], [m[0] for m in c])
        for (fn, fk, fl), (cn, ck, cl) in zip(f, c):
            assert C.sizeof(fk) == C.sizeof(ck) and fl == cl, (tname, fn)
    # ... and the ctypes structs are the header's: same size as the C compiler's (checked through the C driver's use
    # of afmg_opts) and the same member names in the same order
    hdr = re.sub(r"/\*.*?\*/", " ", open(os.path.join(ROOT, "include", "afmg.h")).read(), flags=re.S)
    for tname, struct in (("afmg_opts", _lib.Opts), ("afmg_stencil_desc", _lib.StencilDesc), ("afmg_tree", _lib.TreeDesc),
                          ("afmg_lsf_opts", _lib.LsfOpts), ("afmg_electrode", _lib.Electrode)):
        body = re.search(rf"typedef struct {tname}\s*\{{(.*?)\}}\s*{tname}\s*;", hdr, flags=re.S).group(1)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            first, *rest = decl.split(",")
            names.append(re.search(r"(\w+)\s*(?:\[\d+\])?$", first.strip()).group(1))
            names += [re.match(r"\s*\*?\s*(\w+)", r).group(1) for r in rest]
        assert names == [f[0] for f in struct._fields_], (tname, names)


def test_shim_entry_points_extend_the_reference_signatures():
    """mg_gpu_X(tree, mg, <the reference's arguments>, slot, <the reference's optionals>): the reference's argument
    order first (afivo/src/m_af_multigrid.f90:43, :111, :137, :185, :1188), plus the handle slot that stands for the
    mg_t instance (one per solver: field + 3 Helmholtz modes)."""
    src = fortran_source().lower()
    for sig in ("subroutine mg_gpu_init(tree, mg, slot)", "subroutine mg_gpu_destroy(mg, slot)",
                "subroutine mg_gpu_fas_fmg(tree, mg, set_residual, have_guess, slot)",
                "subroutine mg_gpu_fas_vcycle(tree, mg, set_residual, slot, highest_lvl, standalone)",
                "subroutine mg_gpu_update_operator_stencil(tree, mg, slot)",
                "subroutine mg_gpu_compute_phi_gradient(tree, mg, i_fc, fac, slot, i_norm)"):
        assert sig in src, sig
    public = " ".join(re.findall(r"public\s*::([^\n]*)", src))
    for name in ("mg_gpu_init", "mg_gpu_destroy", "mg_gpu_fas_fmg", "mg_gpu_fas_vcycle", "mg_gpu_update_operator_stencil",
                 "mg_gpu_compute_phi_gradient", "mg_gpu_field_solve", "photoi_gpu_helmh_compute"):
        assert name in public, name


def test_dropin_module_has_the_references_names_and_argument_lists():
    """fortran/m_af_multigrid_dropin.f90: the entry points under the reference's own names and argument lists
    (afivo/src/m_af_multigrid.f90:43, :111, :137, :185, :1188, :1857, :1997), each forwarding to the shim routine of
    the same purpose with the slot resolved from the mg_t."""
    raw = open(os.path.join(ROOT, "fortran", "m_af_multigrid_dropin.f90")).read()
    src = "\n".join(ln.split("!")[0].rstrip() for ln in raw.splitlines()).lower()
    src = re.sub(r"&\s*\n\s*&?", " ", src)
    want = {
        "mg_init": ("tree, mg", "mg_gpu_init"),
        "mg_destroy": ("mg", "mg_gpu_destroy"),
        "mg_fas_fmg": ("tree, mg, set_residual, have_guess", "mg_gpu_fas_fmg"),
        "mg_fas_vcycle": ("tree, mg, set_residual, highest_lvl, standalone", "mg_gpu_fas_vcycle"),
        "mg_update_operator_stencil": ("tree, mg, new_lsf, new_eps", "mg_gpu_update_operator_stencil"),
        "mg_compute_phi_gradient": ("tree, mg, i_fc, fac, i_norm", "mg_gpu_compute_phi_gradient"),
        "mg_compute_field_norm": ("tree, i_fc, i_norm", "mg_gpu_compute_field_norm"),
    }
    shim = fortran_source().lower()
    public = " ".join(re.findall(r"public\s*::([^\n]*)", src))
    for name, (args, target) in want.items():
        m = re.search(rf"subroutine {name}\(([^)]*)\)(.*?)end subroutine {name}", src, flags=re.S)
        assert m, name
        assert re.sub(r"\s+", " ", m.group(1).strip()) == args, (name, m.group(1))
        call = re.search(rf"call {target}\(([^\n]*)\)", m.group(2))
        assert call, (name, target)
        # the forwarded call has as many arguments as the shim routine declares
        decl = re.search(rf"subroutine {target}\(([^)]*)\)", shim).group(1)
        depth, n = 0, 1
        for ch in call.group(1):
            depth += ch == "("
            depth -= ch == ")"
            n += ch == "," and depth == 0
        assert n == len(decl.split(",")), (name, call.group(1), decl)
        assert name in public
    # optional arguments stay optional, intents are the reference's
    assert re.search(r"integer, intent\(in\), optional :: highest_lvl", src)
    assert re.search(r"logical, intent\(in\), optional :: standalone", src)
    assert re.search(r"subroutine mg_fas_vcycle.*?type\(mg_t\), intent\(in\)\s+:: mg", src, flags=re.S)


def test_every_solve_refreshes_what_changes_in_time():
    """Round-1 review findings, kept fixed: the shim includes cpp_macros.h (DTIMES / NDIM), runs the reference's mg_use
    before every solve (tags + stencils of boxes created by af_adjust_refinement), re-evaluates mg%sides_bc and
    mg%lsf_boundary_value before EVERY solve (both carry the time-dependent voltage, src/m_field.f90:481-487,
    590-610), and detects tree changes with a hash of the complete topology instead of three counters."""
    raw = open(os.path.join(ROOT, "fortran", "m_af_multigrid_gpu.f90")).read()
    assert raw.lstrip().startswith('#include "cpp_macros.h"')
    src = fortran_source().lower()
    assert re.search(r"use m_af_stencil, only:[^\n]*af_stencil_index", src)
    assert "use m_af_multigrid, only: mg_use" in src
    prep = re.search(r"subroutine prepare_solve\(tree, mg, slot\)(.*?)end subroutine prepare_solve", src, flags=re.S).group(1)
    order = [prep.index(k) for k in ("call mg_use(tree, mg)", "call sync_tree(tree, mg, slot)", "call sync_bc(tree, mg, slot)",
                                     "afmg_set_lsf_boundary_value(", "call sync_lsf_boundary_values(")]
    assert order == sorted(order)
    for entry in ("mg_gpu_fas_fmg", "mg_gpu_fas_vcycle", "mg_gpu_field_solve", "photoi_gpu_helmh_compute"):
        body = re.search(rf"subroutine {entry}\(.*?end subroutine {entry}", src, flags=re.S).group(0)
        assert "call prepare_solve(" in body and "call finish_solve(tree)" in body, entry
    assert "call mg_use(tree, mg)" in re.search(r"subroutine mg_gpu_compute_phi_gradient\(.*?end subroutine", src, flags=re.S).group(0)
    # sync_tree no longer evaluates boundary conditions and is keyed on the topology hash
    st = re.search(r"subroutine sync_tree\(tree, mg, slot\)(.*?)end subroutine sync_tree", src, flags=re.S).group(1)
    assert "sides_bc" not in st and "topology_hash(tree)" in st and "tree_signature" not in src
    th = re.search(r"function topology_hash\(tree\)(.*?)end function topology_hash", src, flags=re.S).group(1)
    for member in ("%parent", "%children(1)", "%neighbors(n)", "%ix(n)", "lvls(lvl)%ids"):
        assert member in th, member
    # phi goes up only when the device copy is stale; packed buffers are page-locked
    assert "have_guess .and. .not. solvers(slot)%phi_current" in src
    assert "afmg_host_alloc(" in src and "c_f_pointer(buf_ptr, buf" in src
