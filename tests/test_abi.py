"""CPU: the C-ABI library builds, loads, exports every symbol include/cm3env.h declares, and
refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import cm3_b200
from cm3_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "cm3env.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cm3_[a-z_0-9]+)\s*\(", src)))


def test_library_builds_and_loads():
    cm3_b200.build_library()
    lib = L.load_library()
    assert lib.cm3_abi_version() == L.ABI_VERSION == 2


def test_every_declared_symbol_is_exported_and_bound():
    cm3_b200.build_library()
    lib = C.CDLL(cm3_b200.library_path())
    names = header_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), "library does not export %s" % name
        assert name in L.SYMBOLS, "ctypes binding does not cover %s" % name
    assert sorted(L.SYMBOLS) == names


def test_struct_sizes_match_header():
    # natural alignment, no packing: sizes derived by hand from include/cm3env.h
    assert C.sizeof(L.CheckersConfig) == 5 * 4 + 2 * 32 + 4 * 4 + 4 + 8 + 8  # padded to 8
    assert C.sizeof(L.CheckersState) == 32
    assert C.sizeof(L.CheckersOutputs) == 72
    assert C.sizeof(L.ParticleConfig) == 6 * 4 + 8 + 8 * 8 + 4 * 64 + 24
    assert C.sizeof(L.ParticleState) == 48
    assert C.sizeof(L.ParticleOutputs) == 64


def test_header_is_plain_c_and_layouts_match_the_ctypes_binding(tmp_path):
    """include/cm3env.h compiles as C11 (no C++ leaking through the boundary) and every field of
    every struct sits at the offset the ctypes binding assumes."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    structs = {"cm3_checkers_config": L.CheckersConfig, "cm3_checkers_state": L.CheckersState,
               "cm3_checkers_outputs": L.CheckersOutputs, "cm3_particle_config": L.ParticleConfig,
               "cm3_particle_state": L.ParticleState, "cm3_particle_outputs": L.ParticleOutputs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "cm3env.h"', 'int main(void) {']
    for cname, ct in structs.items():
        lines.append('printf("%s sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in ct._fields_:
            lines.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines.append('printf("abi %d max_agents %d max_dst %d\\n", CM3_ABI_VERSION, CM3_MAX_AGENTS, CM3_MAX_DST);')
    lines.append("return 0; }")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call([cc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)], text=True).split("\n")
    seen = 0
    for line in out:
        parts = line.split()
        if len(parts) == 3 and parts[0] in structs:
            ct = structs[parts[0]]
            want = C.sizeof(ct) if parts[1] == "sizeof" else getattr(ct, parts[1]).offset
            assert int(parts[2]) == want, line
            seen += 1
        elif parts and parts[0] == "abi":
            assert (int(parts[1]), int(parts[3]), int(parts[5])) == (L.load_library().cm3_abi_version(), L.MAX_AGENTS, L.MAX_DST)
    assert seen == sum(len(ct._fields_) + 1 for ct in structs.values())


def test_bad_geometry_is_rejected_like_the_reference():
    lib = L.load_library()
    cfg = L.CheckersConfig(n_rows=4, n_columns=8, n_obs=2, n_agents=2, max_steps=33, num_envs=4)
    h = C.c_void_p()
    assert lib.cm3_checkers_create(C.byref(cfg), C.byref(h)) == -2  # checkers.py:16
    cfg = L.CheckersConfig(n_rows=3, n_columns=7, n_obs=2, n_agents=2, max_steps=33, num_envs=4)
    assert lib.cm3_checkers_create(C.byref(cfg), C.byref(h)) == -2  # checkers.py:17
    assert b"odd" in lib.cm3_last_error()
    cfg = L.CheckersConfig(n_rows=7, n_columns=30, n_obs=2, n_agents=2, max_steps=33, num_envs=4)
    assert lib.cm3_checkers_create(C.byref(cfg), C.byref(h)) == -4  # 210 cells: beyond the 64-bit bitboard
    assert b"bitboards" in lib.cm3_last_error()
    cfg = L.CheckersConfig(n_rows=3, n_columns=8, n_obs=2, n_agents=1, max_steps=33, num_envs=4, random_goal=2)
    assert lib.cm3_checkers_create(C.byref(cfg), C.byref(h)) == -1
    cfg = L.CheckersConfig(n_rows=3, n_columns=8, n_obs=2, n_agents=2, max_steps=33, num_envs=4, random_goal=1)
    assert lib.cm3_checkers_create(C.byref(cfg), C.byref(h)) == -1  # the goal redraw is the stage-1 protocol


def test_null_arguments():
    lib = L.load_library()
    assert lib.cm3_checkers_create(None, None) == -1
    assert lib.cm3_checkers_destroy(None) == -1
    assert lib.cm3_particle_destroy(None) == -1
    assert lib.cm3_checkers_step(None, None, None, None, None) == -1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = L.load_library()
    assert L.device_count() == 0
    cfg = L.CheckersConfig(n_rows=3, n_columns=8, n_obs=2, n_agents=2, max_steps=33, num_envs=4)
    cfg.agents_r[0], cfg.agents_r[1], cfg.agents_c[0], cfg.agents_c[1] = 0, 2, 8, 8
    h = C.c_void_p()
    assert lib.cm3_checkers_create(C.byref(cfg), C.byref(h)) == -5
    assert b"no CPU fallback" in lib.cm3_last_error()
    pc = L.ParticleConfig()
    lib.cm3_particle_default_config(C.byref(pc), 4, 33)
    assert (pc.dt, pc.damping, pc.contact_force, pc.contact_margin) == (0.1, 0.25, 100.0, 1e-3)
    assert lib.cm3_particle_create(C.byref(pc), C.byref(h)) == -5
    with pytest.raises(L.Cm3Error):
        cm3_b200.VecCheckers(4, device="cpu")
    with pytest.raises(L.Cm3Error):
        cm3_b200.VecCheckers(4, 3, 8, 2, [0, 2], [8, 8], 2, 33, device="cuda:0")
