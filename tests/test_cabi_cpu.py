"""CPU-only checks of the boundary: the C-ABI library builds, loads, exports every symbol that
include/timet_b200.h declares, validates arguments, and refuses to compute without a GPU."""
import ctypes as C
import os
import re

import pytest
import torch

from timetuning_b200 import _build, _cabi, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "timet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(timet_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_loads():
    lib = _cabi.lib()
    assert os.path.isfile(_build.LIB)
    assert lib.timet_abi_version() == 2


def test_every_declared_symbol_is_exported_and_bound():
    declared = header_functions()
    assert len(declared) >= 20
    handle = C.CDLL(_build.LIB)
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in timet_b200.h but not exported"
    assert sorted(_cabi.EXPORTS) == declared, "ctypes prototypes out of sync with the header"


def test_ff_params_struct_layout():
    assert C.sizeof(_cabi.FFParams) == 48


def test_workspace_queries_are_host_only():
    lib = _cabi.lib()
    p = _cabi.FFParams(32, 8, 28, 28, 384, 200, 7, 6, 5, 1, 0.1, 0)
    n = lib.timet_ff_workspace_bytes(C.byref(p))
    assert 100e6 < n < 2e9
    assert lib.timet_ff_slots(C.byref(p)) == 16
    assert lib.timet_sinkhorn_workspace_bytes(25088, 200) > 200 * 4 * 300


def _tc_plan(*args, **env):
    """timet_ff_tc_plan for FFParams(*args) under the given TIMET_* switches (host-side query: no GPU needed)."""
    import os
    lib = _cabi.lib()
    old = {k: os.environ.get(k) for k in env}
    try:
        for k, v in env.items():
            os.environ[k] = v
        lib.timet_debug_reload_env()
        plan = (C.c_int32 * 8)()
        assert lib.timet_ff_tc_plan(C.byref(_cabi.FFParams(*args)), plan) == 0
        return dict(zip(("kernel", "colblk", "slots", "stages", "a_bytes", "key_cols", "tile_rows", "smem"), plan))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        lib.timet_debug_reload_env()


def test_tc_plan_of_the_baseline_shapes():
    """The persistent tcgen05 kernel's configuration is decided on the host: BASELINE configs[1] gets
    column-blocked query tiles, an 84 KB query tile, 24-slot candidate lists and a three-stage key ring inside 227 KB."""
    cfg2 = _tc_plan(32, 8, 28, 28, 384, 200, 7, 6, 5, 1, 0.1, 0)
    assert cfg2 == dict(kernel=2, colblk=1, slots=24, stages=3, a_bytes=6 * 112 * 128, key_cols=224, tile_rows=4, smem=cfg2["smem"])
    assert cfg2["smem"] <= 227 * 1024
    # the switch that keeps the padding rows inside every K chunk: 96 KB query tile, 32-slot lists, two stages
    old = _tc_plan(32, 8, 28, 28, 384, 200, 7, 6, 5, 1, 0.1, 0, TIMET_TC_PFLAGS="16384")
    assert (old["colblk"], old["slots"], old["stages"], old["a_bytes"]) == (1, 32, 2, 6 * 16384)
    raster = _tc_plan(32, 8, 28, 28, 384, 200, 7, 6, 5, 1, 0.1, 0, TIMET_TC_PFLAGS="256")
    assert (raster["colblk"], raster["slots"], raster["stages"], raster["a_bytes"]) == (0, 32, 2, 6 * 16384)
    # configs[0] (14 x 14), configs[3] (60 x 60, radius 12, k 7), configs[4] (56 x 56, D = 768: streamed query tile)
    cfg1 = _tc_plan(2, 4, 14, 14, 384, 200, 7, 6, 5, 1, 0.1, 0)
    assert (cfg1["kernel"], cfg1["colblk"], cfg1["tile_rows"]) == (2, 0, 9)
    cfg4 = _tc_plan(1, 80, 60, 60, 384, 11, 7, 12, 7, 1, 0.1, 0)
    assert (cfg4["kernel"], cfg4["colblk"], cfg4["key_cols"], cfg4["tile_rows"]) == (2, 0, 240, 2)
    cfg5 = _tc_plan(8, 16, 56, 56, 768, 300, 7, 6, 5, 1, 0.1, 0)
    assert (cfg5["kernel"], cfg5["colblk"], cfg5["a_bytes"], cfg5["key_cols"]) == (2, 0, 0, 224)
    for c in (cfg1, cfg4, cfg5):
        assert c["smem"] <= 227 * 1024 and c["stages"] >= 2
    # other column-blocked widths: the last column block holds W - 24 columns
    for W, rows in ((26, 104), (30, 120), (32, 128)):
        pl = _tc_plan(2, 4, 12, W, 128, 8, 7, 3, 5, 1, 0.1, 0)
        assert (pl["colblk"], pl["a_bytes"]) == (1, 2 * rows * 128), (W, pl)
    # per-item kernel / exact engine
    assert _tc_plan(32, 8, 28, 28, 384, 200, 7, 6, 5, 1, 0.1, 0, TIMET_TC_PERSIST="0")["kernel"] == 1
    assert _tc_plan(2, 4, 28, 28, 384, 200, 7, 6, 12, 1, 0.1, 0)["kernel"] == 0          # top-k 12: exact engine


@pytest.mark.parametrize("field,value", [("n_clips", 0), ("n_frames", 1), ("topk", 0), ("topk", 17), ("n_last_frames", 0),
                                         ("t_begin", 0), ("t_begin", 8), ("radius", -1), ("temperature", 0.0), ("dim", 0)])
def test_ff_validation(field, value):
    lib = _cabi.lib()
    p = _cabi.FFParams(2, 8, 28, 28, 384, 200, 7, 6, 5, 1, 0.1, 0)
    setattr(p, field, value)
    assert lib.timet_ff_workspace_bytes(C.byref(p)) == 0
    assert field.split("_")[0] in lib.timet_last_error().decode() or "ff:" in lib.timet_last_error().decode()


def test_error_codes_without_pointers():
    lib = _cabi.lib()
    rc = lib.timet_sinkhorn(None, 10, 10, 0, 0.05, 3, 1, None, None, None, 0, None)
    assert rc == -1 and b"NULL" in lib.timet_last_error()
    rc = lib.timet_restrict_neighborhood(0, 4, 1, None, None)
    assert rc == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.sinkhorn(torch.rand(5, 7), 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.restrict_neighborhood(4, 4, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.propagate_labels_batched(torch.rand(1, 3, 16, 8), torch.rand(1, 16, 4))


def test_spatial_resolution_duck_typing():
    class FE:
        spatial_resolution = 28

    class Wrapper:
        feature_extractor = FE()

    assert ops._spatial_resolution(FE()) == (28, 28)
    assert ops._spatial_resolution(Wrapper()) == (28, 28)

    class NonSquare:
        spatial_resolution = (60, 106)       # additive: DAVIS 480x854 at patch 8
    assert ops._spatial_resolution(NonSquare()) == (60, 106)


def test_install_binds_the_real_reference_modules_and_uninstall_restores_them():
    """install() against the reference's own modules (imported in place / from the shipped copy): the names the
    reference resolves at call time (time_tuning.py:49,51,147,165; mask_propagation.py:485,821) are rebound, and
    uninstall() puts the originals back.  Binding only -- no compute without a GPU."""
    import ref_loader
    import timetuning_b200 as tb
    from timetuning_b200 import training
    if not ref_loader.available():
        pytest.skip("reference not available (neither /root/reference nor baseline/_ref)")
    mu, mp, tt, _ = ref_loader.load()
    orig = (tt.sinkhorn, tt.propagate_labels, mp.label_propagation, mp.propagate_labels, mp.restrict_neighborhood,
            mp.norm_mask, mu.sinkhorn, tt.TimeT.get_loss, tt.TimeT.get_scores)
    assert tt.sinkhorn is mu.sinkhorn and tt.propagate_labels is mp.propagate_labels
    inst = tb.install(tt, mp, mu, fast_get_loss=True)
    try:
        assert tt.sinkhorn is ops.sinkhorn and tt.propagate_labels is ops.propagate_labels
        assert mp.label_propagation is ops.label_propagation and mp.norm_mask is ops.norm_mask
        assert mp.propagate_labels is ops.propagate_labels and mp.restrict_neighborhood is ops.restrict_neighborhood
        assert mu.sinkhorn is ops.sinkhorn
        assert tt.TimeT.get_loss is training.fast_get_loss and tt.TimeT.get_scores is training.get_scores
        # the reference's own callers look the names up in their module globals at call time
        assert tt.TimeT.find_optimal_assignment.__globals__["sinkhorn"] is ops.sinkhorn
        assert tt.TimeT.make_seg_maps.__globals__["propagate_labels"] is ops.propagate_labels
        assert mp.propagate_labels is not orig[3]
    finally:
        inst.uninstall()
    now = (tt.sinkhorn, tt.propagate_labels, mp.label_propagation, mp.propagate_labels, mp.restrict_neighborhood,
           mp.norm_mask, mu.sinkhorn, tt.TimeT.get_loss, tt.TimeT.get_scores)
    assert all(a is b for a, b in zip(orig, now))


def test_install_on_partial_modules():
    import types
    import timetuning_b200 as tb
    tt = types.SimpleNamespace(sinkhorn=None, propagate_labels=None)
    with tb.install(time_tuning=tt):
        assert tt.sinkhorn is ops.sinkhorn and tt.propagate_labels is ops.propagate_labels
    assert tt.sinkhorn is None and tt.propagate_labels is None
    with pytest.raises(ValueError):
        tb.install(fast_get_loss=True)
