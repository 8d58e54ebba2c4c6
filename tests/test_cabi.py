"""CPU: the C-ABI library loads and exports every symbol include/softrod.h declares.
No compute calls here (no GPU in this container); argument validation that happens
before any CUDA call is exercised."""
import ctypes as C
import os
import re

import pytest

from gym_softrobot_b200 import _native as nat
from gym_softrobot_b200 import build as b

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "softrod.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sr_[a-z0-9_]+)\s*\(", hdr)))


def test_library_builds_and_exports_header_symbols():
    path = b.build_library()
    assert os.path.exists(path)
    lib = C.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 24
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in softrod.h but not exported"
    assert sorted(nat.EXPORTED_SYMBOLS) == syms, "python binding list out of sync with the header"
    assert lib.sr_abi_version() == 1


def test_config_struct_layout_matches_header(tmp_path):
    """ctypes mirror == the C header, as a C compiler lays it out (sizeof and the offset of the last field)."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "softrod.h"\n'
                   'int main(void){printf("%zu %zu %zu\\n", sizeof(sr_config), offsetof(sr_config, spline_max_rate),'
                   ' sizeof(sr_state_view));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    size, off_last, view = map(int, subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split())
    assert C.sizeof(nat.SrConfig) == size
    assert nat.SrConfig.spline_max_rate.offset == off_last
    assert C.sizeof(nat.SrStateView) == view == 8 + 12 * 4


def _cfg(**kw):
    cfg = nat.SrConfig()
    cfg.struct_size = C.sizeof(nat.SrConfig)
    cfg.model, cfg.n_env, cfg.n_elem, cfg.bc_kind = nat.MODEL_SOFT_PENDULUM, 4, 50, nat.BC_PENDULUM_SLIDER
    cfg.dt, cfg.base_length, cfg.base_radius, cfg.density, cfg.youngs_modulus = 1e-4, 1.0, 0.05, 1000.0, 1e6
    cfg.damping_constant = 2e-3
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


@pytest.mark.parametrize("kw,msg", [
    (dict(struct_size=8), "struct_size"),
    (dict(n_env=0), "n_env"),
    (dict(n_elem=2), "n_elem"),
    (dict(dt=0.0), "dt"),
    (dict(model=77), "model"),
    (dict(bc_kind=9), "bc_kind"),
    (dict(muscle_layers_on=1), "muscle_layers_on needs"),          # no transverse muscle / head / tapered assembly
    (dict(head_fixed=1), "muscle-layer kernel only"),
    (dict(n_fixed_sucker=2), "muscle-layer kernel only"),
])
def test_create_rejects_bad_config(kw, msg):
    lib = nat.load_library()
    h = C.c_void_p()
    rc = lib.sr_create(C.byref(_cfg(**kw)), C.byref(h))
    assert rc == -1 and not h.value
    assert msg in lib.sr_last_error().decode()


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = nat.load_library()
    h = C.c_void_p()
    rc = lib.sr_create(C.byref(_cfg()), C.byref(h))
    assert rc == -3 and b"no CUDA device" in lib.sr_last_error()
    with pytest.raises(nat.SoftRodError):
        nat.Handle(model=nat.MODEL_SOFT_PENDULUM, n_env=1, n_elem=50, dt=1e-4, base_length=1.0,
                   base_radius=0.05, density=1000.0, youngs_modulus=1e6)
    out = C.c_double()
    assert lib.sr_measure_fp64_peak(0, C.byref(out)) == -3


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (no CPU fallback)."""
    pkg = os.path.join(ROOT, "gym_softrobot_b200")
    bad = re.compile(r"^\s*(import|from)\s+(rod_oracle|oracle|ref_loader|elastica)\b|#include\s*[<\"].*oracle", re.M)
    n = 0
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                n += 1
                assert not bad.search(open(os.path.join(dirpath, f)).read()), f
    assert n >= 8


def test_integration_doc_stub_lists_the_header_fields_in_order():
    """INTEGRATION.md shows the ctypes mirror a reference maintainer would write: its field list must be the
    header's `sr_config`, name for name, in order (and so must the package's own mirror)."""
    hdr = open(os.path.join(ROOT, "include", "softrod.h")).read()
    body = re.search(r"typedef struct sr_config \{(.*?)\} sr_config;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1]              # drop the type
        fields += [re.sub(r"\[.*?\]", "", n).strip() for n in names.split(",")]
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    stub = re.search(r"class SrConfig\(C\.Structure\):(.*?)\n\n", doc, re.S).group(1)
    doc_fields = re.findall(r'\("([a-z0-9_]+)",', stub)
    assert doc_fields == fields
    assert [f for f, _ in nat.SrConfig._fields_] == fields
