"""CPU: properties of the compiled kernels that the measured performance rests on, read from the built library with
cuobjdump (no GPU needed) — so that a source change which silently loses one of them fails here, not in a profile.

* PISCES tendency kernel: 168 registers (3 blocks of 128 threads per SM), and ALL of its input loads are issued before
  its first data-dependent branch (one DRAM round trip per cell; a guard branch ahead of the arithmetic once split them
  11 + 27).
* NPD tendency kernels: no spills; the parameter-sweep instantiations stay in registers as well.
* scaling + Ω prologue and the PAR scans of the bench workloads: no local memory."""
import os
import re
import shutil
import subprocess

import pytest

from oceanbiome_b200 import _lib

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not available")


def lib_path():
    _lib.load()
    return _lib.LIB_PATH if hasattr(_lib, "LIB_PATH") else os.path.join(os.path.dirname(_lib.__file__), "lib", "libobm_b200.so")


@pytest.fixture(scope="module")
def resources():
    out = subprocess.run([CUOBJDUMP, "-res-usage", lib_path()], capture_output=True, text=True, check=True).stdout
    res = {}
    for name, line in re.findall(r"Function (\S+):\n\s*(REG:.*)", out):
        res[name] = {k: int(v) for k, v in re.findall(r"(\w+)(?:\[0\])?:(\d+)", line)}
    assert res
    return res


def kernels(resources, fragment):
    return {n: r for n, r in resources.items() if fragment in n}


def test_pisces_tendency_kernel_occupancy(resources):
    ks = kernels(resources, "pisces_tendency_kernel")
    assert len(ks) == 4  # <accumulate, all destinations present>
    for name, r in ks.items():
        assert r["REG"] <= 168, (name, r)  # 3 × 128 threads × 168 registers ≤ 65 536
        assert r["STACK"] <= 192, (name, r)  # the spill that L1 absorbs (≈ 116 B of it live)


def test_pisces_input_loads_form_one_batch():
    sym = "_ZN3obm22pisces_tendency_kernelILb1ELb1EEEvNS_10PiscesArgsE"
    sass = subprocess.run([CUOBJDUMP, "-sass", "-fun", sym, lib_path()], capture_output=True, text=True, check=True).stdout
    ops = re.findall(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)
    assert len(ops) > 3000
    loads = [i for i, op in enumerate(ops) if op.startswith("LDG")]
    # the fast path reads 38 values per cell (28 3-D inputs, 2 × 2 faces of w, 4 column fields, z, …); the out-of-line exact
    # path has its own copy further down
    first_batch = [i for i in loads if i < loads[0] + 400]
    assert len(first_batch) >= 37, len(first_batch)
    between = ops[first_batch[0]:first_batch[-1]]
    assert not any(op == "BRA" for op in between), "a branch splits the input loads of the PISCES kernel"
    assert not any(op.startswith(("DSETP", "BSSY")) for op in between), "a data-dependent test sits between the input loads"


def test_npd_kernels_do_not_spill(resources):
    ks = kernels(resources, "npd_tendency_kernel")
    assert len(ks) == 24  # 3 nutrient × 4 detritus choices × (plain, parameter sweep)
    for name, r in ks.items():
        assert r["STACK"] <= 40 and r["REG"] <= 144, (name, r)  # (40 B: the richest parameter-sweep instantiation only)


def test_prologue_and_scans_use_no_local_memory(resources):
    # 48 registers for 5 blocks per SM cost the DIAG scans 16 – 24 B; 64 registers for 8 blocks per SM cost the fused
    # scaling + Ω prologue 40 B with the FP32 pre-solve (timed against 7 blocks / 72 registers: profiles/r04_kernel_variants.txt;
    # ncu counts no executed local loads / stores on the sea-water path: the spilled values live in the cold branches)
    for fragment, stack in (("scale_negative_calcite_kernel", 40), ("scale_negative_kernel", 0), ("par_twoband_kernel", 16),  # 4 blocks / 64 registers: 16 B, timed −3 % against 3 blocks without
                           
                            ("par_multiband_kernel", 40)):  # 4 bands + diagnostics: 40 B; the 3-band PISCES scan: ≤ 24 B
        ks = kernels(resources, fragment)
        assert ks, fragment
        for name, r in ks.items():
            assert r["STACK"] <= stack, (name, r)
