"""CPU tests of the drop-in boundary: the C ABI library loads, exports every
symbol include/gpuexec.h declares, and refuses to run without a GPU (no fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import opentenbase_b200 as g
from conftest import HAS_GPU

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = g.lib()
    decl = g.declared_symbols()
    assert len(decl) >= 40
    missing = [s for s in decl if not hasattr(L, s)]
    assert not missing, missing
    assert L.gx_abi_version() == 1


def test_header_is_plain_c():
    """The boundary must be consumable from the provider's C code: compile the header as C11."""
    src = '#include "gpuexec.h"\nint main(void){ gx_agg_plan p; gx_heap_desc d; (void)p; (void)d; return GX_OK; }\n'
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-fsyntax-only", "-I", g.INCLUDE_DIR, "-x", "c", "-"],
                   input=src.encode(), check=True)


def test_struct_layout_matches_header():
    """ctypes mirrors vs. sizeof() as gcc sees the header."""
    prog = ('#include <stdio.h>\n#include "gpuexec.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu\\n",'
            'sizeof(gx_pred),sizeof(gx_expr),sizeof(gx_agg),sizeof(gx_agg_plan),sizeof(gx_heap_desc),sizeof(gx_host_table));return 0;}')
    exe = os.path.join(ROOT, "tests", ".sizeof_probe")
    subprocess.run(["gcc", "-I", g.INCLUDE_DIR, "-x", "c", "-", "-o", exe], input=prog.encode(), check=True)
    out = subprocess.run([exe], capture_output=True, check=True).stdout.split()
    os.unlink(exe)
    want = [C.sizeof(x) for x in (g.GxPred, g.GxExpr, g.GxAgg, g.GxAggPlan, g.GxHeapDesc, g.GxHostTable)]
    assert [int(x) for x in out] == want
    import oracle as O
    assert C.sizeof(O.GxAggPlan) == C.sizeof(g.GxAggPlan)


@pytest.mark.skipif(HAS_GPU, reason="this box has a GPU")
def test_no_gpu_means_loud_failure_not_fallback():
    with pytest.raises(g.GxError) as e:
        g.Context(0)
    assert e.value.status == g.GX_ERR_NODEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under opentenbase_b200/ may import, link or dlopen it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "opentenbase_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"liboracle|import oracle|from oracle|otb_oracle\.h|orc_[a-z_]+\(", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
    deps = subprocess.run(["ldd", g.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in deps
