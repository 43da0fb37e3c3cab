"""The C-ABI shared library loads and exports every symbol include/simgan_b200.h declares, and the ctypes
binding table covers exactly that set.  Host-only entry points (layouts, workspace sizes, argument
validation) are exercised; nothing here launches a kernel."""
import ctypes as C
import os
import re

import pytest

from simgan_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "simgan_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", src))
    names -= {"sg_allreduce_fn"}
    return names


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.lib()


def test_every_declared_symbol_is_exported_and_bound(lib):
    decl = declared_symbols()
    assert len(decl) >= 20
    raw = C.CDLL(build.LIB_PATH)
    for name in decl:
        assert hasattr(raw, name), "header declares %s but the library does not export it" % name
    assert decl == set(_lib.SIGNATURES), (decl ^ set(_lib.SIGNATURES))


def test_no_stray_exports():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", build.LIB_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    ours = {s for s in exported if s.startswith("sg_")}
    assert ours == declared_symbols()


def test_layouts_match_parameter_counts(lib):
    offs, total = _lib.policy_layout(14, 64, 7)
    sizes = [64 * 14, 64, 64 * 64, 64, 64 * 14, 64, 64 * 64, 64, 64, 1, 7 * 64, 7, 7]
    assert sum(sizes) == 10767                                   # policy params cfg 2
    for i in range(12):
        assert offs[i + 1] - offs[i] >= sizes[i] and offs[i] % 4 == 0
    assert total >= offs[12] + 7 and total % 4 == 0
    offs, total = _lib.disc_layout(25, 100)
    assert sum([2500, 100, 10000, 100, 100, 1]) == 12801          # D params cfg 2
    assert offs == [0, 2500, 2600, 12600, 12700, 12800] and total == 12804


def test_argument_validation_reports_messages(lib):
    cfg = _lib.PpoConfig()
    assert lib.sg_ppo_workspace_bytes(C.byref(cfg)) == -1
    assert b"non-positive" in lib.sg_last_error()
    cfg.obs_dim, cfg.hidden, cfg.act_dim, cfg.T, cfg.N = 14, 64, 7, 8, 4
    cfg.ppo_epoch, cfg.num_mini_batch, cfg.mini_batch_size = 1, 2, 16
    cfg.row_begin, cfg.row_end, cfg.first_adam_step = 0, 17, 1
    assert lib.sg_ppo_workspace_bytes(C.byref(cfg)) == -1
    assert b"shard" in lib.sg_last_error()
    cfg.row_end = 16
    assert lib.sg_ppo_workspace_bytes(C.byref(cfg)) > 0
    assert lib.sg_ppo_phase_cycles_offset(C.byref(cfg)) > 0
    cfg.mode = 7
    assert lib.sg_ppo_workspace_bytes(C.byref(cfg)) == -1
    # null pointers are rejected before anything touches the device
    rc = lib.sg_compute_returns(None, None, None, None, None, None, 4, 4, 0.99, 0.95, 1, 1, None)
    assert rc == 1 and b"null pointer" in lib.sg_last_error()
    dcfg = _lib.DiscConfig()
    dcfg.feat_dim, dcfg.hidden, dcfg.batch_size, dcfg.n_steps = 25, 100, 128, 3
    dcfg.row_begin, dcfg.row_end, dcfg.first_adam_step, dcfg.gp_lambda = 0, 128, 1, 10.0
    assert lib.sg_disc_workspace_bytes(C.byref(dcfg)) > 0
    assert lib.sg_relabel_workspace_bytes(0, 4) == -1
    assert lib.sg_launch_count() == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.SgError, match="no CPU fallback"):
        _lib.lib()
