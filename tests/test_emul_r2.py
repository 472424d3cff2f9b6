"""Round-2 CPU checks through the test-only emulation build: the PCS seam, the verifier transcript prefix, the multi-device
dispatcher (two emulated devices), max_ring_size below capacity against reference-generated proofs."""

import pytest

from dot_ring_b200 import _native
from dot_ring_b200 import engine as engine_mod
from tests.host.emul import emulation_library


@pytest.fixture(scope="module")
def api():
    lib = emulation_library()
    old = _native._default
    _native.set_default_library(lib)
    eng = engine_mod.Engine(0, window_bits=4, library=lib, srs_points=1537)
    engine_mod.set_default_engine(eng, 0)
    import dot_ring_b200 as pkg

    yield pkg
    engine_mod.set_default_engine(None, 0)
    eng.close()
    _native.set_default_library(old)


def test_pcs_seam_matches_reference(api):
    from tests import pcs_cases

    pcs_cases.check_pcs(api)


def test_verifier_transcript_prefix_matches_reference(api):
    from tests import pcs_cases

    pcs_cases.check_transcript_prefix(api)


def test_small_max_ring_reference_proofs(api):
    from tests import pcs_cases

    pcs_cases.check_small_ring_golden(api, domains=(512,))


def test_engine_pool_shards_like_one_device(api):
    from tests import pool_cases

    lib = emulation_library()
    pool = engine_mod.EnginePool(devices=[0, 0], window_bits=4, library=lib, srs_points=1537)
    try:
        pool_cases.check_pool_matches_single(api, pool, engine_mod.default_engine(), n=5)
        pool_cases.check_range_split_msm(pool, engine_mod.default_engine().ctx)
    finally:
        pool.close()


def test_oversized_suite_fields_raise(api):
    """an oversized DST / suite id must raise before any memmove into the fixed-size struct fields."""
    with pytest.raises(ValueError):
        _native.make_suite(b"x" * 33, b"dst", (0, 1), (0, 1))
    with pytest.raises(ValueError):
        _native.make_suite(b"id", b"d" * 65, (0, 1), (0, 1))
