"""CPU tests of the boundary: libsdpb200.so loads, exports every symbol include/sdpb200.h declares,
agrees on struct layout, and refuses to work without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sdpb200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdpb_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(S):
    lib = S.abi.load()
    declared = _declared_functions()
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in sdpb200.h but not exported"
    assert sorted(S.abi.EXPORTS) == declared
    assert lib.sdpb_abi_version() == 4


def test_struct_layout_matches_c_compiler(S, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sdpb200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(sdpb_model), sizeof(sdpb_options),'
                   'offsetof(sdpb_model,gamma), offsetof(sdpb_model,q_mul), offsetof(sdpb_model,price_t),'
                   'offsetof(sdpb_model,reserve2));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    M = S.abi.SdpbModel
    assert [int(x) for x in out] == [C.sizeof(M), C.sizeof(S.abi.SdpbOptions), M.gamma.offset,
                                     M.q_mul.offset, M.price_t.offset, M.reserve2.offset]
    lib = S.abi.load()
    assert lib.sdpb_sizeof_model() == C.sizeof(M)


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "c.c"
    src.write_text('#include "sdpb200.h"\nint main(void){return SDPB_ABI_VERSION-1;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    "-c", str(src), "-o", str(tmp_path / "c.o")], check=True)


def test_argument_errors_need_no_gpu(S):
    lib = S.abi.load()
    h = C.c_void_p()
    assert lib.sdpb_create(None, None, C.byref(h)) == S.abi.SDPB_ERR_ARG
    spec = S.inventory_model(S.poisson_pmf([3, 3]), max_order=5, inv_min=-5, inv_max=5)
    m = spec.to_struct()
    m.struct_size = 12
    assert lib.sdpb_create(C.byref(m), None, C.byref(h)) == S.abi.SDPB_ERR_ARG
    assert b"struct_size" in lib.sdpb_last_error(None)
    # a grid that is not exactly representable is refused before any device work
    bad = S.inventory_model(S.poisson_pmf([3, 3]), max_order=5, inv_min=-5, inv_max=5, step=0.3)
    assert lib.sdpb_create(C.byref(bad.to_struct()), None, C.byref(h)) == S.abi.SDPB_ERR_OFFGRID
    bad2 = S.inventory_model([np.array([[0.5, 1.0]])], max_order=5, inv_min=-5, inv_max=5)
    assert lib.sdpb_create(C.byref(bad2.to_struct()), None, C.byref(h)) == S.abi.SDPB_ERR_OFFGRID


def test_no_cpu_fallback(S):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    spec = S.inventory_model(S.poisson_pmf([3, 3]), max_order=5, inv_min=-5, inv_max=5)
    with pytest.raises(S.SdpbError) as e:
        S.Solver(spec)
    assert e.value.code == S.abi.SDPB_ERR_NO_DEVICE


def test_product_never_touches_the_oracle():
    """The shipped package and library must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "stochastic-inventory_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"
    so = os.path.join(pkg, "libsdpb200.so")
    deps = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    assert "oracle" not in deps


def test_getpmf_matches_survey_shapes(S):
    rows = S.poisson_pmf([20, 40, 60, 40], 0.9999)
    assert [len(r) for r in rows] == [40, 67, 92, 67]
    assert [len(r) for r in S.poisson_pmf([151.0], 0.9999)] == [200]
    assert [len(r) for r in S.poisson_pmf([10.0], 0.9999)] == [25]
    for r in rows:
        assert abs(r[:, 1].sum() - 1.0) < 1e-12
        assert np.array_equal(r[:, 0], np.arange(len(r)))
    # continuous branch (GetPmf.java:126-129): Normal(30, 7.5), step 1
    r = S.GetPmf([S.NormalDist(30, 7.5)], 0.9999, 1).getpmf()[0]
    assert r[0, 0] == float(int(S.NormalDist(30, 7.5).inverseF(1 - 0.9999)))
    assert abs(r[:, 1].sum() - 1.0) < 1e-12
    # UniformInt branch (GetPmf.java:97-111)
    r = S.GetPmf([S.UniformIntDist(2, 5)] * 2, 0.99, 1).getpmf()
    assert np.array_equal(r[1][:, 0], [2, 3, 4, 5]) and np.all(r[1][:, 1] == 0.25)


# ---- struct layouts: compiler == committed table == ctypes binding == Java StructLayouts -------------------------
def _layout_from_compiler(tmp_path):
    exe = tmp_path / "print_layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "print_layout.c"),
                    "-o", str(exe)], check=True)
    return subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout


def _parse_layout(text):
    out = {}
    for line in text.splitlines():
        if line.startswith("#") or not line.strip():
            continue
        name, off, size = line.split()
        out[name] = (int(off), int(size))
    return out


def test_layout_table_is_current(tmp_path):
    """java/LAYOUT.txt is what the C compiler produces from include/sdpb200.h today, and it lists EVERY field."""
    now = _layout_from_compiler(tmp_path)
    committed = open(os.path.join(ROOT, "java", "LAYOUT.txt")).read()
    assert _parse_layout(now) == _parse_layout(committed), "regenerate java/LAYOUT.txt with tools/print_layout.c"
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for struct in ("sdpb_model", "sdpb_options", "sdpb_grid", "sdpb_stats", "sdpb_reached_model"):
        body = re.search(r"typedef struct " + struct + r" \{(.*?)\} " + struct + ";", src, flags=re.S).group(1)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                fields.append(re.findall(r"[A-Za-z_][A-Za-z_0-9]*", re.sub(r"\[[^\]]*\]", "", part))[-1])
        listed = [k.split(".")[1] for k in _parse_layout(committed) if k.startswith(struct + ".") and not k.endswith(".sizeof")]
        assert listed == fields, struct


def test_ctypes_binding_matches_layout_table(S):
    lay = _parse_layout(open(os.path.join(ROOT, "java", "LAYOUT.txt")).read())
    for cname, cls in (("sdpb_model", S.abi.SdpbModel), ("sdpb_options", S.abi.SdpbOptions),
                       ("sdpb_grid", S.abi.SdpbGrid), ("sdpb_stats", S.abi.SdpbStats),
                       ("sdpb_reached_model", S.abi.SdpbReachedModel)):
        assert lay[cname + ".sizeof"][0] == C.sizeof(cls)
        for fname, _ in cls._fields_:
            f = getattr(cls, fname)
            assert lay[f"{cname}.{fname}"] == (f.offset, f.size), (cname, fname)


def test_java_struct_layouts_match_layout_table():
    """The Java binding cannot be compiled here (no JDK), but its StructLayouts can be read: field order, names and
    value layouts of SdpB200.MODEL / OPTIONS / GRID must give exactly the offsets of java/LAYOUT.txt under the C
    alignment rules Panama applies (natural alignment, no implicit padding -- so the C struct must need none)."""
    lay = _parse_layout(open(os.path.join(ROOT, "java", "LAYOUT.txt")).read())
    src = open(os.path.join(ROOT, "java", "sdp", "b200", "SdpB200.java")).read()
    sizes = {"JAVA_INT": 4, "JAVA_DOUBLE": 8, "JAVA_LONG": 8, "ADDRESS": 8, "JAVA_DOUBLE2": 16}
    for jname, cname in (("MODEL", "sdpb_model"), ("OPTIONS", "sdpb_options"), ("GRID", "sdpb_grid"),
                         ("REACHED_MODEL", "sdpb_reached_model")):
        body = re.search(r"StructLayout " + jname + r" = MemoryLayout\.structLayout\((.*?)\);\n", src, flags=re.S).group(1)
        body = body.replace("MemoryLayout.sequenceLayout(2, JAVA_DOUBLE)", "JAVA_DOUBLE2")
        fields = re.findall(r"(JAVA_INT|JAVA_DOUBLE2|JAVA_DOUBLE|JAVA_LONG|ADDRESS)\.withName\(\"([a-zA-Z_0-9]+)\"\)", body)
        off = 0
        for typ, name in fields:
            assert off % min(sizes[typ], 8) == 0, f"{jname}.{name}: Panama struct layouts have no implicit padding"
            assert lay[f"{cname}.{name}"] == (off, sizes[typ]), (jname, name)
            off += sizes[typ]
        assert off == lay[cname + ".sizeof"][0], jname
        assert len(fields) == sum(1 for k in lay if k.startswith(cname + ".")) - 1


def test_python_constants_match_header_enums(S):
    """Every enumerator of include/sdpb200.h that _abi.py mirrors has the header's value (kernel names, cost kinds,
    quantisers, flags, status codes): a renumbered enum would otherwise select the wrong kernel silently."""
    import re
    text = open(os.path.join(ROOT, "include", "sdpb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    enums = dict(re.findall(r"\b(SDPB_[A-Z0-9_]+)\s*=\s*(\d+u?\s*<<\s*\d+|-?\d+)", text))
    assert len(enums) >= 50, len(enums)

    def val(expr):
        m = re.match(r"(\d+)u?\s*<<\s*(\d+)", expr)
        return int(m.group(1)) << int(m.group(2)) if m else int(expr)

    A = S.abi
    checked = 0
    for name, expr in enums.items():
        short = name[len("SDPB_"):]
        for cand in (short, name):
            if hasattr(A, cand):
                assert getattr(A, cand) == val(expr), (name, getattr(A, cand), expr)
                checked += 1
                break
    assert checked >= 40, checked
    for k in ("KERNEL_LEAD_Q2M", "KERNEL_COLLAPSED", "KERNEL_CASH_TAIL", "KERNEL_LEAD_Q2"):
        assert getattr(A, k) == val(enums["SDPB_" + k])
    # the (uncompiled) Java binding carries a few of them as literals
    java = open(os.path.join(ROOT, "java", "sdp", "b200", "SdpB200.java")).read()
    lits = re.findall(r"\b(KERNEL_[A-Z0-9_]+)\s*=\s*(\d+)", java)
    assert len(lits) >= 5
    for name, v in lits:
        assert int(v) == val(enums["SDPB_" + name]), name
