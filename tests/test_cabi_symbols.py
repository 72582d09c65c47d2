"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/ttb200.h declares (no compute calls without a GPU), and argument errors that are decided on
the host keep the reference's error classes."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    import __graft_entry__ as g
    g.build()


def test_library_exports_every_declared_symbol():
    _build()
    hdr = open(os.path.join(ROOT, "include", "ttb200.h")).read()
    names = sorted(set(re.findall(r"\b(ttb_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 35
    lib = ctypes.CDLL(os.path.join(ROOT, "tensortrains.jl_b200", "ttb200", "libttb200.so"))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.ttb_version() >= 100


def test_library_targets_sm100a_only():
    _build()
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "tensortrains.jl_b200", "ttb200", "libttb200.so")],
                         capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tensortrains.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src, os.path.join(dp, f)


def test_host_decided_argument_errors():
    _build()
    import numpy as np
    import ttb200 as T
    with pytest.raises(ValueError):
        T.TensorTrain([np.zeros((1, 2)), np.zeros((2, 1))])       # N > 2 required (tensor_train.jl:13)
    with pytest.raises(ValueError):
        T.TensorTrain([np.zeros((1, 2, 2)), np.zeros((3, 1, 2))])  # check_bond_dims
    with pytest.raises(ValueError):
        T._range([3, 2])                                           # issorted(indices)
    assert T._range(range(2, 5)) == (2, 4) and T._range([]) == (1, 0)
    assert T._per_bonds((3, 5, 2))[0] == [3, 3, 3, 3]              # (d, L) builds L-1 sites (periodic_tensor_train.jl:44)
    assert T._open_bonds((3, 5, 2))[0] == [1, 3, 3, 3, 3, 1]
