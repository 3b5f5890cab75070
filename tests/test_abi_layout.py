"""ABI guard for the reference-side binding (/root/reference/src/dynamic_ob.jl:30,36 are what julia/TBCuda.jl overrides):
the struct layouts a C compiler gives include/tbcuda.h must equal the ctypes mirror (tensorbranching.jl_b200/_lib.py) and
the struct definitions of julia/TBCuda.jl, field by field.  No Julia toolchain exists in this image, so this -- plus
a plain-C client that drives the library exactly as a `ccall` host would -- is the executable evidence for that file."""
import ctypes as C
import json
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import tbcuda
from helpers import golden_branches, load_golden, to_sliced
from tensorbranching_lib import L  # noqa: F401  (see conftest: alias of tensorbranching.jl_b200._lib)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ABI = os.path.join(ROOT, "tests", "abi")


@pytest.fixture(scope="module")
def c_layout(tmp_path_factory):
    exe = tmp_path_factory.mktemp("abi") / "abi_layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ABI, "abi_layout.c"), "-o", str(exe)])
    return json.loads(subprocess.check_output([str(exe)]))


@pytest.mark.parametrize("name", ["tb_options", "tb_network", "tb_plan_stats", "tb_step_info"])
def test_ctypes_mirror_matches_the_header(c_layout, name):
    cls = getattr(L, name)
    want = c_layout[name]
    assert C.sizeof(cls) == want["__sizeof__"][0]
    fields = [f[0] for f in cls._fields_]
    assert fields == [k for k in want if k != "__sizeof__"]  # same fields, same order
    for f in fields:
        d = getattr(cls, f)
        assert [d.offset, d.size] == want[f], f


JL_SIZES = {"Int32": 4, "UInt32": 4, "Int64": 8, "UInt64": 8, "Float64": 8, "Cdouble": 8, "Cint": 4, "UInt8": 1}


def _julia_struct(name):
    """field list [(name, size)] of `struct name ... end` in julia/TBCuda.jl (isbits structs follow the C layout rules)"""
    src = open(os.path.join(ROOT, "julia", "TBCuda.jl")).read()
    m = re.search(r"^struct " + name + r"\n(.*?)^end", src, re.S | re.M)
    assert m, name
    body = re.sub(r"#.*", "", m.group(1))
    out = []
    for decl in re.split(r"[;\n]", body):
        decl = decl.strip()
        if not decl:
            continue
        fname, ftype = [x.strip() for x in decl.split("::")]
        out.append((fname, 8 if ftype.startswith("Ptr{") else JL_SIZES[ftype]))
    return out


@pytest.mark.parametrize("jl,c", [("TbOptions", "tb_options"), ("TbNetwork", "tb_network")])
def test_julia_structs_match_the_header(c_layout, jl, c):
    want = c_layout[c]
    off = 0
    fields = _julia_struct(jl)
    assert [f for f, _ in fields] == [k for k in want if k != "__sizeof__"]
    for f, size in fields:
        off = (off + size - 1) // size * size  # natural alignment, as C and Julia isbits structs both do
        assert [off, size] == want[f], (jl, f)
        off += size
    align = max(s for _, s in fields)
    assert (off + align - 1) // align * align == want["__sizeof__"][0]


def test_julia_ccall_signatures_name_exported_symbols():
    src = open(os.path.join(ROOT, "julia", "TBCuda.jl")).read()
    lib = L.load()
    for sym in set(re.findall(r"ccall\(\(:(\w+), LIB\)", src)):
        assert sym in L.EXPORTS and hasattr(lib, sym), sym


def write_branch_file(path, branches):
    with open(path, "wb") as f:
        f.write(struct.pack("<i", len(branches)))
        for b in branches:
            s = to_sliced(b)
            if s.code is None or b.nv == 0:
                f.write(struct.pack("<iiid", 0, 0, 0, float(b.r)))
                continue
            code = s.code
            f.write(struct.pack("<iiid", b.nv, len(code.ixs), len(code.leaf_labels), float(b.r)))
            for arr in (code.leaf_off, code.leaf_labels, code.node_left, code.node_right):
                f.write(np.ascontiguousarray(arr, dtype="<i4").tobytes())


@pytest.fixture(scope="module")
def c_client(tmp_path_factory):
    exe = tmp_path_factory.mktemp("abi") / "c_client"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ABI, "c_client.c"), "-o", str(exe),
                           "-L", os.path.dirname(L.LIB_PATH), "-ltbcuda", "-Wl,-rpath," + os.path.dirname(L.LIB_PATH)])
    return str(exe)


def test_c_client_builds_against_the_header(c_client):
    assert os.path.exists(c_client)


def _run_client(c_client, path, devices):
    out = subprocess.check_output([c_client, str(path), devices], text=True, timeout=600).strip().splitlines()
    while out and not out[0].startswith("devices "):  # NCCL may print a version banner first
        out.pop(0)
    head = out[0].split()
    vals = np.array([float(ln.split()[0]) for ln in out[1:]])
    stat = np.array([int(ln.split()[1]) for ln in out[1:]])
    return dict(devices=int(head[1]), max=float(head[3])), vals, stat


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["0", "0,0", "0,0,0"])
def test_plain_c_client_contracts_a_branch_list(c_client, tmp_path, devices):
    """one device, and the multi-device sharding path on repeated device 0 (host combine: NCCL cannot span one GPU twice;
    the real NCCL path is tests/test_multi_gpu_device.py on a >= 2 GPU box)"""
    rec = load_golden("rr100_sc10_unit.json")
    brs = golden_branches(rec)
    path = tmp_path / "branches.bin"
    write_branch_file(path, brs)
    head, vals, stat = _run_client(c_client, path, devices)
    assert head["devices"] == len(devices.split(","))
    assert not stat.any() and np.array_equal(vals, np.asarray(rec["values"])) and head["max"] == rec["exact"]


@pytest.mark.gpu
def test_plain_c_client_cfg2(c_client, tmp_path):
    """BASELINE config 2 (2 997 branches) from plain C on every GPU of the box, against the committed golden values"""
    import bench
    import torch
    rec = load_golden("baseline_cfg2.json")
    brs = bench.make_workload("cfg2")
    path = tmp_path / "cfg2.bin"
    write_branch_file(path, brs)
    n = max(1, torch.cuda.device_count())
    head, vals, stat = _run_client(c_client, path, ",".join(str(d) for d in range(n)))
    assert head["devices"] == n and not stat.any()
    assert np.array_equal(vals, np.asarray(rec["values"])) and head["max"] == rec["mis"]


@pytest.mark.gpu
def test_plain_c_client_branching_tables(tmp_path):
    """tb_branching_table / tb_table_configs / tb_branching_tables from plain C on known tables (tests/abi/c_table.c)"""
    exe = tmp_path / "c_table"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ABI, "c_table.c"), "-o", str(exe),
                           "-L", os.path.dirname(L.LIB_PATH), "-ltbcuda", "-Wl,-rpath," + os.path.dirname(L.LIB_PATH)])
    out = subprocess.check_output([str(exe)], text=True, timeout=300)
    assert out.strip().splitlines()[-1] == "ok"
