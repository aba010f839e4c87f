"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/delfem_b200.h declares,
fails loudly without a GPU, and the product never touches the oracle."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(dfb):
    L = dfb.lib()
    names = dfb.header_symbols()
    assert len(names) >= 70
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    # and nothing with the dfb_ prefix is exported that the header does not declare
    out = subprocess.run(["nm", "-D", "--defined-only", dfb.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\b(dfb_[a-z0-9_]+)\b", out)))
    extra = [n for n in exported if n not in names]
    assert not extra, extra


def test_header_is_plain_c_and_cites_the_reference():
    hdr = os.path.join(ROOT, "include", "delfem_b200.h")
    src = open(hdr).read()
    assert "torch" not in src.lower()
    for cite in ["sparse_square.rs:218-261", "sparse_square.rs:264-347", "sparse_square.rs:180-214", "merge.rs:3", "merge.rs:51",
                 "sparse_ilu.rs:87-219", "sparse_square.rs:3-55", "linearsystem.rs:67", "laplace_tri2.rs:3", "solid_linear_tri2.rs:23"]:
        assert cite in src, cite
    # compiles as C
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", hdr], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_status_codes_match_header(dfb):
    src = open(os.path.join(ROOT, "include", "delfem_b200.h")).read()
    for name, val in dfb.STATUS.items():
        key = "DFB_OK" if name == "OK" else f"DFB_ERR_{name}"
        m = re.search(rf"{key}\s*=\s*(\d+)", src)
        assert m and int(m.group(1)) == val, key


def test_no_gpu_means_loud_failure_not_fallback(dfb):
    """on a box without a CUDA device the product must raise, never compute on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(dfb.DfbError) as ei:
        dfb.Ctx(0)
    assert ei.value.status == dfb.STATUS["CUDA"]
    assert "no CPU fallback" in str(ei.value)
    assert dfb.lib().dfb_version().decode().startswith("delfem_b200")


def test_product_never_references_the_oracle():
    pkg = os.path.join(ROOT, "del-fem_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath or dirpath.endswith("lib"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, os.path.join(dirpath, f)
    r = subprocess.run(["ldd", os.path.join(pkg, "lib", "libdelfem_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in r and "torch" not in r


def test_sass_contains_tma_bulk_copies(dfb):
    """the SpMV streams the matrix with TMA bulk copies: SASS shows UBLKCP (B200_PROFILING.md), and the target is sm_100a"""
    out = subprocess.run(["cuobjdump", "-sass", dfb.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    funcs = out.split("Function : ")
    # the shipped default k_spmv_packed<3, false> (variant 20): exactly one bulk-copy site per tile buffer stage + the prologue,
    # no halo logic (the separate <3, true> instantiation carries the flag loads), and the ring kernel of variant 4
    for prefix in ("_Z13k_spmv_packedILi3ELb0ELi8EE", "_Z13k_spmv_packedILi3ELb1ELi8EE", "_Z11k_spmv_ringILi3EE"):
        k = [f for f in funcs if f.startswith(prefix)]
        assert k, f"{prefix} not found in the library"
        assert "UBLKCP" in k[0], prefix
        assert "SYNCS" in k[0], prefix  # mbarrier
    plain = [f for f in funcs if f.startswith("_Z13k_spmv_packedILi3ELb0ELi8EE")][0]
    halo = [f for f in funcs if f.startswith("_Z13k_spmv_packedILi3ELb1ELi8EE")][0]
    assert "LDG.E.64.STRONG.SYS" not in plain and "LDG.E.64.STRONG.SYS" in halo  # only the halo kernel polls peer flags
    # the halo kernel carries the row code twice: interior tiles run the single-domain instruction stream, boundary tiles a copy
    # whose ghost gathers are L2-coherent loads (the neighbours write the landing zone while the kernel runs)
    assert 2 * plain.count("DFMA") == halo.count("DFMA")
    assert "LDG.E.64.STRONG.GPU" in halo


def test_python_mirror_signatures_cover_header(dfb):
    from del_fem_b200 import _lib
    L = dfb.lib()
    for n in dfb.header_symbols():
        fn = getattr(L, n)
        if n in ("dfb_last_error_message", "dfb_version"):
            continue
        assert fn.argtypes is not None, f"{n}: ctypes argtypes not declared"


def test_cpp_mirror_compiles():
    """the header-only C++ mirror and its test program compile and link against the C-ABI library (no GPU needed to build)"""
    import tempfile
    lib_dir = os.path.join(ROOT, "del-fem_b200", "lib")
    with tempfile.TemporaryDirectory() as d:
        r = subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "test_solid_linear_tri2.cpp"), "-o",
                            os.path.join(d, "t"), "-L" + lib_dir, "-ldelfem_b200", "-Wl,-rpath," + lib_dir], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_rust_sys_crate_is_in_sync_with_the_header():
    """rust/delfem-b200-sys/src/lib.rs (the binding a del-fem maintainer adds) is generated from the header:
    every exported function and constant must be declared, and the committed file must be what the generator produces now"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_rust_sys.py"), "--check"])
    assert r.returncode == 0, "run `python tools/gen_rust_sys.py` and commit rust/delfem-b200-sys/src/lib.rs"
    import re
    hdr = open(os.path.join(root, "include", "delfem_b200.h")).read()
    rs = open(os.path.join(root, "rust", "delfem-b200-sys", "src", "lib.rs")).read()
    names = set(re.findall(r"\b(dfb_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)))
    assert len(names) > 80
    for n in names:
        assert re.search(r"pub fn %s\(" % n, rs), n


def test_cpp_mirror_compiles_against_the_header():
    """del-fem_b200/cpp/delfem_b200.hpp is header-only over the C ABI: every template it offers must at least compile and
    instantiate (g++ -fsyntax-only, no CUDA, no GPU); the end-to-end run is a GPU test (tests/cpp/test_solid_linear_tri2.cpp)"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = r'''
#include "delfem_b200.hpp"
using namespace delfem_b200;
template class sparse_square::Matrix<1>;
template class sparse_square::Matrix<3>;
template class sparse_ilu::Preconditioner<2>;
template class sparse_ilu::Preconditioner<3>;
void use(const Ctx &ctx, sparse_square::Matrix<3> &m, sparse_square::Matrix<3> &m2, DevVec<3> &v) {
  sparse_ilu::Preconditioner<3> pc(DFB_PRECOND_JACOBI);
  pc.initialize(ctx, m);
  pc.initialize_ilu0(ctx, m);
  pc.initialize_iluk(ctx, m, 2);
  m.add_scaled(0.5, m2);
  m.add_to_diagonal(2.0, v);
  Mesh mesh(ctx, std::vector<std::size_t>{0, 1, 2, 3}, 4, std::vector<double>(12, 0.0), 3);
  m.assemble(ctx, mesh, DFB_ELEM_SOLID_LINEAR_TET, 1.0, 1.0);
  sparse_ilu::copy_value(pc, m);
  sparse_ilu::decompose(pc);
  sparse_ilu::solve_preconditioning_vec(v, pc);
  auto h = sparse_square::preconditioned_conjugate_gradient(v, v, v, v, 1e-5, 10, m, pc);
  auto g = sparse_square::conjugate_gradient(v, v, v, v, 1e-5, 10, m);
  (void)h; (void)g;
}
'''
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(root, "include"), "-I",
                        os.path.join(root, "del-fem_b200", "cpp"), "-x", "c++", "-"], input=src.encode(), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-3000:]
