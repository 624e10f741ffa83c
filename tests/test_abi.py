"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports
every symbol include/apdgicp.h declares; parameter defaults equal the reference
constructors' values. No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(REPO, "include", "apdgicp.h")).read()
    return sorted(set(re.findall(r"APD_API [\w\s\*]*?\b(apd_\w+)\(", text)))


def test_header_declares_the_reference_surface():
    syms = declared_symbols()
    for must in ["apd_create", "apd_set_source", "apd_set_target", "apd_align", "apd_linearize", "apd_compute_error",
                 "apd_update_correspondences", "apd_fitness", "apd_swap_source_and_target", "apd_clear_source",
                 "apd_clear_target", "apd_set_source_covariances", "apd_get_target_covariances", "apd_align_batch",
                 "apd_comm_init"]:
        assert must in syms
    assert len(syms) >= 35


def test_library_exports_every_declared_symbol(gorio):
    if not os.path.exists(gorio.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    lib = gorio.load()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    lib.apd_abi_version.restype = ctypes.c_int
    assert lib.apd_abi_version() == 3  # 2: apd_params.variant (FastGICP behind the same kernels); 3: FastVGICP (voxel_* fields, hooks)


def test_default_params_match_reference_constructors(gorio):
    lib = gorio.load()
    p = gorio.ApdParams()
    assert lib.apd_default_params(ctypes.byref(p)) == 0
    # fast_apdgicp_impl.hpp:14-28, fast_apdgicp.hpp:116-118, lsq_registration_impl.hpp:11-24
    assert p.k_correspondences == 20
    assert p.regularization == gorio.REG_PLANE
    assert p.max_correspondence_distance == pytest.approx(3.4028234663852886e38)
    assert (p.dist_var, p.azimuth_var, p.elevation_var) == (0.86, 0.5, 1.0)
    assert p.max_iterations == 64 and p.optimizer == gorio.OPT_LM
    assert (p.rotation_epsilon, p.transformation_epsilon) == (2e-3, 5e-4)
    assert p.lm_max_iterations == 10 and p.lm_init_lambda_factor == 1e-9
    assert p.variant == 0 and ctypes.sizeof(p) == 112  # APD_VARIANT_APDGICP; the struct the header declares
    assert p.voxel_search == 2 and p.voxel_resolution == 1.0 and p.voxel_mode == 0  # DIRECT1, 1.0, ADDITIVE (fast_vgicp_impl.hpp:22-24)


def test_no_cpu_fallback_without_gpu(gorio):
    """Without an sm_100 device apd_create must fail (the product has no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = gorio.load()
    h = ctypes.c_void_p()
    assert lib.apd_create(0, ctypes.byref(h)) != 0
    assert not h.value


def test_oracle_and_product_share_the_abi(gorio):
    from oracle_binding import oracle_lib
    lib = oracle_lib()
    for s in gorio._binding.CORE_SYMBOLS:
        assert hasattr(lib, "apdo_" + s), s
