"""CPU test of the drop-in boundary: the C-ABI library loads without a GPU and exports every function that
include/ksw2.h and include/ksw2_b200.h declare (no compute call is made here); the result struct keeps the reference ABI."""
import ctypes as C
import os
import re

import ksw2_b200 as K

INC = os.path.join(K.ROOT, "include")


def declared_functions(path):
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", " ", txt)
    txt = re.sub(r"#[^\n]*", " ", txt)
    names = []
    for stmt in txt.split(";"):
        stmt = " ".join(stmt.split())
        if "(" not in stmt or stmt.startswith("typedef") or "{" in stmt.split("(")[0]:
            continue
        m = re.match(r"^(?:extern \"C\" \{ )?(?:const )?[A-Za-z_][A-Za-z0-9_ ]*?[ \*]+([A-Za-z_][A-Za-z0-9_]*) ?\(", stmt)
        if m:
            names.append(m.group(1))
    return names


def test_library_exports_every_declared_symbol():
    K.build()
    L = C.CDLL(K.LIB_PATH)
    names = declared_functions(os.path.join(INC, "ksw2.h")) + declared_functions(os.path.join(INC, "ksw2_b200.h"))
    assert len(names) >= 25, names
    for must in ("ksw_extz2_sse", "ksw_extd2_sse", "ksw_exts2_sse", "ksw_extz", "ksw_extd", "ksw2b_align", "ksw2b_plan_run"):
        assert must in names, (must, names)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_result_struct_keeps_reference_abi():
    # reference ksw2.h:33-42 on x86-64: sizeof 56, cigar pointer at offset 48
    assert C.sizeof(K.ExtzT) == 56 and K.ExtzT.cigar.offset == 48 and K.ExtzT.m_cigar.offset == 32
    assert K.RESULT_DTYPE.itemsize == 64


def test_no_cpu_fallback_without_a_device():
    """without a CUDA device the library must refuse to create a context (never compute on the CPU)"""
    import torch
    if torch.cuda.is_available():
        return
    L = K.lib()
    assert not L.ksw2b_create(0)
    assert b"no CPU path" in L.ksw2b_last_error() or b"CUDA" in L.ksw2b_last_error()


def test_headers_are_plain_c_and_the_example_links(tmp_path):
    """include/*.h compile as C99 (the callers are C programs) and examples/dropin.c links against the library without unresolved symbols"""
    import subprocess
    K.build()
    exe = os.path.join(str(tmp_path), "dropin")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + INC, os.path.join(K.ROOT, "examples", "dropin.c"),
                           "-L" + K.PKG_DIR, "-lksw2_b200", "-Wl,-rpath," + K.PKG_DIR, "-Wl,--no-undefined", "-o", exe])
    assert os.path.exists(exe)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + INC, os.path.join(K.ROOT, "examples", "multi_gpu.c"),
                           "-L" + K.PKG_DIR, "-lksw2_b200", "-Wl,-rpath," + K.PKG_DIR, "-Wl,--no-undefined", "-o", exe + "_multi"])
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + INC, os.path.join(K.ROOT, "examples", "batch_adapter.c"),
                           "-L" + K.PKG_DIR, "-lksw2_b200", "-Wl,-rpath," + K.PKG_DIR, "-Wl,--no-undefined", "-o", exe + "_adapter"])
    # every function the two headers declare can be named from C (prototype check: take the addresses)
    names = declared_functions(os.path.join(INC, "ksw2.h")) + declared_functions(os.path.join(INC, "ksw2_b200.h"))
    src = os.path.join(str(tmp_path), "all.c")
    with open(src, "w") as f:
        f.write('#include "ksw2.h"\n#include "ksw2_b200.h"\nvoid *tab[] = {' + ", ".join(f"(void*)(size_t)&{n}" for n in names) + "};\nint main(void) { return tab[0] == 0; }\n")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + INC, src, "-L" + K.PKG_DIR, "-lksw2_b200", "-Wl,-rpath," + K.PKG_DIR, "-o", exe + "2"])
