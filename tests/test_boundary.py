"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/admm_b200.h
declares, its POD structs have the layout the ctypes / MEX bindings assume, it refuses to run without a
CUDA device (no CPU fallback), and the product never touches oracle/."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "admm_b200.h")


@pytest.fixture(scope="module")
def lib(pkg):
    import __graft_entry__ as graft
    graft._load_build_module().build()
    return pkg.load()


def declared_functions():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(admmb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib, pkg):
    names = declared_functions()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/admm_b200.h but not exported"
    assert set(pkg._lib.EXPORTS) == set(names)


def test_version_matches_header(lib):
    m = re.search(r"#define ADMMB_VERSION (\d+)", open(HDR).read())
    assert lib.admmb_version() == int(m.group(1))


def test_struct_layouts_match_the_c_compiler(pkg, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "admm_b200.h"\nint main(void){'
                   'printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(admmb_problem), sizeof(admmb_opts), sizeof(admmb_result),'
                   'offsetof(admmb_problem, block_par), offsetof(admmb_opts, xupdate), offsetof(admmb_result, stats));'
                   'printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(admmb_generator), sizeof(admmb_scp), sizeof(admmb_scp_result),'
                   'offsetof(admmb_generator, e), offsetof(admmb_scp, control), offsetof(admmb_scp_result, stats));return 0;}')
    exe = tmp_path / "sz"
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    L = pkg._lib
    want = [C.sizeof(L.Problem), C.sizeof(L.Opts), C.sizeof(L.Result), L.Problem.block_par.offset,
            L.Opts.xupdate.offset, L.Result.stats.offset,
            C.sizeof(L.Generator), C.sizeof(L.Scp), C.sizeof(L.ScpResult), L.Generator.e.offset, L.Scp.control.offset,
            L.ScpResult.stats.offset]
    assert got == want


def test_no_cpu_fallback_without_a_device(lib, pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = lib.admmb_create(C.byref(h), None, 1)
    assert rc == pkg._lib.E_NODEVICE
    assert b"no CPU fallback" in lib.admmb_last_error(None)
    with pytest.raises(pkg.AdmmError):
        pkg.Solver()


def test_product_never_imports_or_links_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "admm-library_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".m", ".sh")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle|oracle/_build|libadmm_ocp_cpu|#include\s+\".*oracle", txt, re.M):
                    bad.append(f)
    assert not bad, f"product files reference the oracle: {bad}"


def test_layout_conversion_and_factor_packing(pkg, P):
    prob, _ = P.lqr_tracking(batch=3, N=4, seed=1, per_problem=True)
    m = pkg.solver.to_c_layout(prob)
    # MATLAB column-major [6x6xNxBd]: element (i,j) of stage k of problem b at i + 6j + 36k + 36N b
    A = prob["A"]
    flat = m["A"].reshape(-1)
    assert flat[2 + 6 * 4 + 36 * 1 + 36 * 4 * 2] == A[2, 1, 2, 4]
    assert m["dyn_batched"] == 1 and m["q_batched"] == 1 and m["par_batched"] == 0
    rng = np.random.default_rng(0)
    parts = {k: rng.standard_normal((4, r, c)) for k, (_, r, c, _) in pkg.solver.FAC_LAYOUT.items()}
    back = pkg.solver.unpack_factor(pkg.solver.pack_factor(parts, 4))
    for k in parts:
        assert np.array_equal(parts[k], back[k])


def test_mex_gateway_syntax_against_stub_header():
    """MATLAB is absent: the gateway is only syntax-checked against a stub mex.h."""
    subprocess.run(["bash", os.path.join(ROOT, "admm-library_b200", "mex", "check_syntax.sh")], check=True)


def test_matlab_sources_present_and_consistent():
    m = open(os.path.join(ROOT, "admm-library_b200", "matlab", "admm_solve.m")).read()
    o = open(os.path.join(ROOT, "oracle", "admm_ocp.m")).read()
    assert "function [x, z, u, hist] = admm_solve(prob, opts)" in m
    assert "function [x, z, u, hist] = admm_ocp(prob, opts)" in o
    assert "UNEXECUTED" in o and "unpinned" in o
