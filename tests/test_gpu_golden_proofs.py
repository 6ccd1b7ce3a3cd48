"""GPU tests against the reference's GOLDEN PROOFS: `halo2_proofs/tests/plonk_api_proof.bin` (k = 5, two instances) and the fifteen
stored proofs of halo2_gadgets' test circuits (k = 11) are ACCEPTED through the engine -- Params::new on the device
(hash_to_curve generators, g_lagrange by EC-iFFT), the instance column's commit_lagrange, point decompression, the multiopen
MSMs, compute_s built on the device into the resident g_scalars, and MSM::eval over the resident generator table
(halo2_b200.multiopen / halo2_b200.verifier) -- and a flipped bit, a wrong public input or another key is rejected.
Reference-held known answers: this pins the device's verifier path on vectors the reference's own tests hold."""
import pytest

pytestmark = pytest.mark.gpu

from oracle import pasta  # noqa: E402
from tests import plonk_verifier as PV  # noqa: E402

CASES = PV.load_golden_proofs()
DELTA = PV.scalar_delta(pasta.P_MOD)


@pytest.fixture(scope="module")
def arms():
    import halo2_b200
    from halo2_b200 import lib as L
    L.init()
    out = {}
    for k in (5, 11):
        prm = halo2_b200.Params.new("vesta", k)                    # Params::<EqAffine>::new(k), whole, on the device
        out[k] = PV.EngineArm(halo2_b200, "vesta", k, prm.g, prm.g_lagrange, prm.w, prm.u)
        prm.close()
    yield out
    for a in out.values():
        a.close()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_reference_golden_proof_verifies_on_device(arms, case):
    vk = PV.PinnedKey(case["key_text"])
    arm = arms[vk.k]
    assert PV.verify_proof(arm, vk, case["proof"], case["instances"], DELTA)
    assert PV.verify_proof(arm, vk, case["proof"], case["instances"], DELTA)      # again: pooled buffers, replayed graphs
    proof = case["proof"]
    for off in (5, len(proof) // 2, len(proof) - 40, len(proof) - 3):
        bad = bytearray(proof)
        bad[off] ^= 0x04
        assert not PV.verify_proof(arm, vk, bytes(bad), case["instances"], DELTA), off
    assert not PV.verify_proof(arm, vk, proof[:-32], case["instances"], DELTA)


def test_golden_proof_binds_instances_and_key_on_device(arms):
    case = CASES[0]
    vk = PV.PinnedKey(case["key_text"])
    arm = arms[5]
    assert not PV.verify_proof(arm, vk, case["proof"], [[[2]], [[3]]], DELTA)
    assert not PV.verify_proof(arm, vk, case["proof"], [[[2]]], DELTA)
    other = PV.PinnedKey(case["key_text"].replace("query_index: 2,\n                                    column_index: 2,",
                                                  "query_index: 3,\n                                    column_index: 2,", 1))
    assert not PV.verify_proof(arm, other, case["proof"], case["instances"], DELTA)
