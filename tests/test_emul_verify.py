"""CPU run of the verification parity cases: kernels through the test-only emulation build (tests/host/emul.py)."""

import pytest

from dot_ring_b200 import _native
from tests import edge_cases
from tests import msm_cases
from tests import verify_cases as cases
from tests.host.emul import emulation_library
from tests.ring_fixtures import native_srs


@pytest.fixture(scope="module")
def ctx():
    c = _native.Context(0, emulation_library())
    yield c
    c.close()


@pytest.fixture(scope="module")
def srs(ctx):
    s = native_srs(ctx, 1537, 4)
    yield s
    s.close()


def test_pairing(ctx):
    cases.pairing_kats(ctx)


def test_pedersen_vectors(ctx):
    cases.pedersen_vectors(ctx)


def test_tiny_vectors(ctx):
    cases.tiny_vectors(ctx)


def test_thin_vectors(ctx):
    cases.thin_vectors(ctx)


def test_vrf_batch_fixtures(ctx):
    cases.vrf_batch_fixtures(ctx)


def test_ring8_verify(srs):
    cases.ring8_vectors(srs, count=2)


def test_w3f_verifier_vectors(ctx):
    cases.w3f_vectors(ctx)


def test_g1_msm_vs_oracle(ctx):
    msm_cases.msm_vs_oracle(ctx, [1, 5, 130, 700])
    msm_cases.synthetic_property(ctx, [300])
    msm_cases.large_ntt(ctx, [8192], {8192})


def test_edge_cases(ctx, srs):
    edge_cases.empty_batches(ctx, srs)
    edge_cases.te_msm_matches_oracle(ctx)
    edge_cases.ring_capacity_and_bad_keys(srs)
    edge_cases.ragged_inputs_match_oracle(srs, ((512, 5),), n_items=2)


def test_vrf_vectors_through_the_large_batch_kernels(ctx):
    """Tiny / Thin / Pedersen batches above the cross-over use the one-thread-per-item kernels: same vectors, same verdicts."""
    lib = ctx.library.lib
    lib.dr_vrf_verify_set_coop_threshold(0)
    try:
        cases.pedersen_vectors(ctx)
        cases.tiny_vectors(ctx)
        cases.thin_vectors(ctx)
        cases.vrf_batch_fixtures(ctx)
    finally:
        lib.dr_vrf_verify_set_coop_threshold(8192)
