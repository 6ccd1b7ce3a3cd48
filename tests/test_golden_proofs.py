"""CPU checks against the reference's GOLDEN PROOFS (no GPU): `halo2_proofs/tests/plonk_api_proof.bin` (two instances, k = 5) and
the fifteen stored proofs of halo2_gadgets' test circuits (k = 11: ECC chip, Sinsemilla, Merkle, range checks) are accepted by the
restated verifier -- plonk::verify_proof driven by the pinned keys (tests/plonk_verifier.py), multiopen::verify_proof,
commitment::verify_proof, the final MSM over the hash_to_curve generators of Params::new (oracle/pasta.py) -- and by the engine's
host mirror over the test-only ABI stand-in.  These are reference-held known answers: they pin the transcript, the key's
transcript representative (the compact Debug string rebuilt from the pretty literal), hash_to_curve, commit_lagrange, MSM,
compute_s / compute_b, lagrange_interpolate and the whole verifier-side control flow at once; any flipped bit is rejected."""
import numpy as np
import pytest

from oracle import cref, pasta
from tests import fake_engine
from tests import plonk_verifier as PV

CASES = PV.load_golden_proofs()
DELTA = PV.scalar_delta(pasta.P_MOD)


class FastOracleArm(PV.OracleArm):
    """The oracle's verifier with its one big multiexp (msm.rs:175) through the C restatement."""

    def finish(self, guard):
        sc, bs = guard.use_challenges().terms()
        return not cref.best_multiexp(self.curve, cref.ints_to_bytes(sc), cref.affines_to_bytes(bs), 4).any()


@pytest.fixture(scope="module")
def gens():
    """Params::<EqAffine>::new(k) (poly/commitment.rs:38-114) for k = 5 (with g_lagrange: the instance column is committed) and 11."""
    out = {}
    P5 = pasta.Params.new(pasta.VESTA, 5)
    out[5] = (cref.affines_to_bytes(P5.g), cref.affines_to_bytes(P5.g_lagrange), cref.affines_to_bytes([P5.w]), cref.affines_to_bytes([P5.u]))
    g, w, u = pasta.params_generators(pasta.VESTA, 11)
    gb = cref.affines_to_bytes(g)
    out[11] = (gb, gb, cref.affines_to_bytes([w]), cref.affines_to_bytes([u]))     # no instance columns at k = 11: g_lagrange is not used
    return out


def test_fixture_is_what_the_reference_holds():
    assert [c["name"] for c in CASES][0] == "plonk_api" and len(CASES) == 16
    sizes = {c["name"]: len(c["proof"]) for c in CASES}
    # expected_proof_size of the gadget tests (e.g. halo2_gadgets/src/utilities/lookup_range_check.rs, ecc.rs, sinsemilla.rs)
    assert sizes["plonk_api"] == 4160 and sizes["lookup_range_check"] == 1888 and sizes["ecc_chip"] == 3872 and sizes["sinsemilla_chip"] == 4576


def test_compact_debug_form():
    vk = PV.PinnedKey(CASES[0]["key_text"])
    s = vk.compact
    assert "\n" not in s and s.startswith('PinnedVerificationKey { base_modulus: "0x4000') and s.endswith("] } }")
    assert "rotation: Rotation(0) }" in s and "Rotation(-1)" in s and "constants: [], minimum_degree: None }" in s
    assert "Column { index: 1, column_type: Advice }" in s and ", )" not in s and ",)" not in s and "( " not in s and "[ " not in s
    assert PV.pretty_to_compact("A {\n    b: [\n        1,\n        (\n            2,\n            3,\n        ),\n    ],\n    c: D(\n        4,\n    ),\n}") \
        == "A { b: [1, (2, 3)], c: D(4) }"
    assert (vk.k, vk.extended_k, vk.degree(), vk.blinding_factors(), len(vk.gates), len(vk.lookups)) == (5, 7, 4, 5, 2, 1)
    assert vk.scalar_modulus == pasta.P_MOD and vk.base_modulus == pasta.Q_MOD
    assert vk.omega == pasta.omega_for_k("fp", 5)
    big = PV.PinnedKey(next(c for c in CASES if c["name"] == "sinsemilla_chip")["key_text"])
    assert (big.k, big.extended_k, big.degree(), len(big.gates), len(big.lookups)) == (11, 14, 9, 76, 3)
    assert big.omega == pasta.omega_for_k("fp", 11)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_reference_golden_proof_verifies(gens, case):
    vk = PV.PinnedKey(case["key_text"])
    arm = FastOracleArm(case["curve"], vk.k, *gens[vk.k])
    assert PV.verify_proof(arm, vk, case["proof"], case["instances"], DELTA)
    proof = case["proof"]
    for off in (5, len(proof) // 2, len(proof) - 40, len(proof) - 3):          # a commitment, an evaluation, c, f
        bad = bytearray(proof)
        bad[off] ^= 0x04
        assert not PV.verify_proof(arm, vk, bytes(bad), case["instances"], DELTA), off
    assert not PV.verify_proof(arm, vk, proof[:-32], case["instances"], DELTA)


def test_golden_proof_binds_instances_and_key(gens):
    case = CASES[0]
    vk = PV.PinnedKey(case["key_text"])
    arm = FastOracleArm("vesta", 5, *gens[5])
    assert not PV.verify_proof(arm, vk, case["proof"], [[[2]], [[3]]], DELTA)          # another public input in the second instance
    assert not PV.verify_proof(arm, vk, case["proof"], [[[2]]], DELTA)                 # one instance only: the proof carries two
    assert not PV.verify_proof(arm, vk, case["proof"], [[[2], [2]], [[2]]], DELTA)     # Error::InvalidInstances (verifier.rs:77-81)
    # any change to the key changes its transcript representative (src/plonk.rs:75-86), hence every challenge
    other = PV.PinnedKey(case["key_text"].replace("query_index: 2,\n                                    column_index: 2,", "query_index: 3,\n                                    column_index: 2,", 1))
    assert other.compact != vk.compact and other.transcript_repr() != vk.transcript_repr()
    assert not PV.verify_proof(arm, other, case["proof"], case["instances"], DELTA)
    # the plain (Python multiexp) oracle arm agrees with the fast one
    assert PV.verify_proof(PV.OracleArm("vesta", 5, *gens[5]), vk, case["proof"], case["instances"], DELTA)


@pytest.mark.parametrize("name", ["plonk_api", "lookup_range_check", "merkle_chip"])
def test_golden_proof_through_the_host_mirror(gens, name):
    """halo2_b200.multiopen.verify_proof / halo2_b200.verifier (MSM with resident g_scalars, fused compute_s) on the same golden
    proofs, the C ABI replaced by tests/fake_engine.py: the mirror's host logic is pinned on reference-held vectors too."""
    import halo2_b200
    case = next(c for c in CASES if c["name"] == name)
    vk = PV.PinnedKey(case["key_text"])
    with fake_engine.installed() as fake:
        arm = PV.EngineArm(halo2_b200, case["curve"], vk.k, *gens[vk.k])
        assert PV.verify_proof(arm, vk, case["proof"], case["instances"], DELTA)
        assert fake.calls.count("h2_poly_compute_s") == 1 and fake.calls.count("h2_msm_registered_polys") == 1
        bad = bytearray(case["proof"])
        bad[len(bad) - 40] ^= 1
        assert not PV.verify_proof(arm, vk, bytes(bad), case["instances"], DELTA)
        arm.close()


@pytest.mark.parametrize("engine", [False, True], ids=["oracle", "host-mirror"])
def test_golden_proof_under_every_strategy(gens, engine):
    """The three strategies the reference's test runs its proofs through (tests/plonk_api.rs:497-583) on the stored proof:
    SingleVerifier, the AccumulationVerifier (Guard::compute_g + use_g, verifier.rs:45-62) and BatchVerifier (the same proof
    twice, as the reference does at :567-577, every MSM scaled by a random factor and accumulated: plonk/verifier/batch.rs:83-131);
    a batch with one tampered proof fails."""
    import contextlib
    import halo2_b200
    case = CASES[0]
    vk = PV.PinnedKey(case["key_text"])
    with (fake_engine.installed() if engine else contextlib.nullcontext()):
        arm = PV.EngineArm(halo2_b200, "vesta", 5, *gens[5]) if engine else FastOracleArm("vesta", 5, *gens[5])
        if not engine:
            arm.accumulate = lambda guard: PV.OracleArm.accumulate(arm, guard)
        assert PV.verify_proof(arm, vk, case["proof"], case["instances"], DELTA)
        assert PV.verify_proof(arm, vk, case["proof"], case["instances"], DELTA, process=arm.accumulate)
        bad = bytearray(case["proof"])
        bad[-40] ^= 1
        assert not PV.verify_proof(arm, vk, bytes(bad), case["instances"], DELTA, process=arm.accumulate)
        factors = cref.bytes_to_ints(cref.gen_scalars("fp", 12345, 2))
        for proofs, want in (([case["proof"], case["proof"]], True), ([case["proof"], bytes(bad)], False)):
            kept = []
            keep = lambda guard: kept.append(arm.use_challenges(guard)) or True
            assert all(PV.verify_proof(arm, vk, p, case["instances"], DELTA, process=keep) for p in proofs)
            assert arm.batch_eval(kept, factors) == want
        arm.close()
