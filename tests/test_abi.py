"""CPU: the C-ABI shared library loads and exports every symbol include/ppc_b200.h declares.  No compute calls."""
import ctypes as C
import os
import re
import subprocess

import pytest

from particlephysicsc_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ppc_b200.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(api.LIB_PATH):
        api.build_library()
    return api.load_library()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b((?:particles|ppc_)[A-Za-z_0-9]+)\s*\(", text)
    return sorted(set(names))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(api.REFERENCE_SYMBOLS + api.EXTENSION_SYMBOLS + api.SLAB_SYMBOLS + api.SNAPSHOT_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", api.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    for name in declared_symbols():
        assert name in exported, name


def test_reference_entry_points_keep_their_names():
    # exactly the solver / store symbols the reference's main.c links against
    assert set(api.REFERENCE_SYMBOLS) == {"particlesCreate", "particlesFree", "particlesAddParticle", "particlesRemoveParticle",
                                          "particlesUpdate", "particlesCollisionsCheck_Grid"}


def test_struct_layout_matches_reference():
    # includes/structs.h:9-22 is 72 bytes: seven pointers + two size_t, in this order
    assert C.sizeof(api.Particles_t) == 72
    names = [f[0] for f in api.Particles_t._fields_]
    assert names == ["v_x", "v_y", "p_x", "p_y", "colors", "mass", "radius", "length", "size"]
    assert api.Particles_t.length.offset == 56 and api.Particles_t.size.offset == 64
    assert C.sizeof(api.Vector2) == 8


def test_header_compiles_as_c_and_next_to_the_oracle_stub():
    """The header is plain C, and coexists with a raylib.h / the reference's structs.h included first."""
    src = '#include "ppc_b200.h"\nint main(void){ struct Particles_t p; (void)p; return sizeof(struct Pair) == 8 ? 0 : 1; }\n'
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                   input=src, text=True, check=True)
    src2 = '#include <raylib.h>\n#include "ppc_b200.h"\nint main(void){ Color c = RED; (void)c; return 0; }\n'
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "oracle", "stub"), "-I",
                    os.path.join(ROOT, "include"), "-x", "c", "-"], input=src2, text=True, check=True)


def test_product_package_never_imports_the_oracle():
    """oracle/ is the checker: nothing under particlephysicsc_b200/ may import, link or execute it."""
    pkg = os.path.join(ROOT, "particlephysicsc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for pat in (r'#include\s*[<"][^>"]*oracle', r"^\s*(from|import)\s+oracle", r"libppc_oracle", r"libref", r"dlopen"):
                    assert not re.search(pat, text, flags=re.M), (os.path.join(dirpath, f), pat)
