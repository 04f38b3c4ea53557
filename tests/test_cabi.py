"""CPU tests: the C-ABI library loads and exports every symbol include/oceananigans_b200.h declares; the host
layer rejects out-of-scope configurations loudly (no fallbacks)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "oceananigans_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(oc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_bound_in_python():
    from oceananigans_b200 import _lib
    assert set(_declared_symbols()) == set(_lib.SYMBOLS)


def test_cuda_library_loads_and_exports_every_symbol():
    import __graft_entry__ as ge
    from oceananigans_b200 import _lib
    if not os.path.exists(ge.LIB):
        ge.build()
    lib = _lib.Library(ge.LIB)           # resolves every symbol, checks the ABI version
    assert lib.oc_abi_version() == _lib.OC_ABI_VERSION
    import ctypes as C
    for name in _declared_symbols():
        assert hasattr(lib.dll, name), name
    cfg = _lib.oc_config()
    lib.oc_config_init(C.byref(cfg))     # pure host call; no compute without a GPU
    assert cfg.abi_version == _lib.OC_ABI_VERSION == 6 and cfg.amd_has_Cb == 0 and cfg.ab2_chi == 0.1 and abs(cfg.gravity - 9.80665) < 1e-15


def test_missing_library_fails_loudly(tmp_path):
    from oceananigans_b200 import _lib
    with pytest.raises(_lib.OceananigansB200Error, match="no CPU fallback"):
        _lib.Library(str(tmp_path / "nope.so"))


def test_out_of_scope_configurations_are_errors():
    import oceananigans_b200 as ob
    assert ob.WENO(order=7).buffer == 4 and ob.WENO(order=9).buffer == 5          # in scope since round 2
    with pytest.raises(NotImplementedError):
        ob.WENO(order=11)
    with pytest.raises(ValueError):
        ob.WENO(order=6)                                                             # "defined only for odd orders"
    with pytest.raises(NotImplementedError):
        ob.Centered(order=6)
    with pytest.raises(NotImplementedError):
        ob.UpwindBiased(order=7)
    with pytest.raises(NotImplementedError):
        ob.RectilinearGrid(np.float64, size=(4, 4, 4), x=[0, 0.1, 0.3, 0.6, 1.0], y=(0, 1), z=(0, 1))      # stretched x: out of scope
    with pytest.raises(ValueError):
        ob.RectilinearGrid(np.float64, size=(4, 4, 4), x=(0, 1), y=(0, 1), z=[0, 0.1, 0.3, 0.6, 1.0],
                           topology=(ob.Periodic, ob.Periodic, ob.Periodic))                                # stretched z must be Bounded
    gs = ob.RectilinearGrid(np.float64, size=(4, 4, 4), x=(0, 1), y=(0, 1), z=[0, 0.1, 0.3, 0.6, 1.0])
    assert gs.z_faces is not None and gs.Lz == 1.0 and gs.dz is None
    with pytest.raises(ValueError):
        ob.RectilinearGrid(np.float64, size=(4, 4), extent=(1, 1, 1))
    g = ob.RectilinearGrid(np.float32, size=(4, 4, 4), extent=(1, 1, 1))
    assert g.H == (3, 3, 3) and g.FT is np.float32
    assert abs(g.dz - np.float32(0.25)) == 0


def test_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof of every C-ABI struct, printed by a C program compiled against include/oceananigans_b200.h with gcc,
    equal the ctypes (and therefore the Julia `struct`) layouts."""
    import ctypes as C
    import subprocess
    from oceananigans_b200 import _lib
    structs = {"oc_config": _lib.oc_config, "oc_field_info": _lib.oc_field_info, "oc_clock": _lib.oc_clock, "oc_bc": _lib.oc_bc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void) {']
    for name, st in structs.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for fname, _ in st._fields_:
            lines.append(f'  printf("{name}.{fname} %zu\\n", offsetof({name}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for name, st in structs.items():
        assert int(out[name]) == C.sizeof(st), name
        for fname, _ in st._fields_:
            assert int(out[f"{name}.{fname}"]) == getattr(st, fname).offset, (name, fname)
