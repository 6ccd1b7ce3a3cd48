"""GPU test (runs last): a REAL PLONK proof of the reference's own test circuit (halo2_proofs/tests/plonk_api.rs:21-420, k = 5,
two instances) is produced ON THE DEVICE by plonk::create_proof composed from the engine's reference-facing API
(tests/plonk_prover.create_proof_engine: resident polynomials, device transforms, Ast programs in both bases, batch_invert and
the running product, the lookup permutation, fixed-base commits, batched evaluations, the multi-point opening and the opening
argument) under the reference's GOLDEN verifying key -- the same 4 160 bytes as the oracle's prover with the same randomness --
and is accepted by the engine's verifier and by the restated reference verifier that the reference's sixteen golden proofs pin.

This composition was validated without a GPU (tests/test_real_proof.py: the same code over the ABI stand-in, device bodies on the
host emulation, identical bytes); every entry point it calls has its own GPU parity test."""
import pytest

pytestmark = pytest.mark.gpu

from oracle import cref, pasta  # noqa: E402
from tests import multiopen_cases as MC  # noqa: E402
from tests import plonk_api_circuit as circ  # noqa: E402
from tests import plonk_prover as PP  # noqa: E402
from tests import plonk_verifier as PV  # noqa: E402
from tests import prover_replay as R  # noqa: E402
from tests import test_real_proof as TR  # noqa: E402


def test_real_proof_on_the_device():
    import halo2_b200
    from halo2_b200 import lib as L
    L.init()
    c = pasta.VESTA
    vk = PV.PinnedKey(TR.CASE["key_text"])
    prm = halo2_b200.Params.new("vesta", 5)                          # Params::<EqAffine>::new(5) on the device
    try:
        gens = (prm.g, prm.g_lagrange, prm.w, prm.u)
        P = pasta.Params.from_generators(c, 5, [cref.bytes_to_affine(x) for x in prm.g], cref.bytes_to_affine(prm.w[0]), cref.bytes_to_affine(prm.u[0]))
        fixed = circ.fixed_columns(TR.M, TR.ZETA)
        sigma = circ.permutation_columns(TR.M, vk.omega, TR.DELTA)
        inst = [[[2]], [[2]]]
        want = TR.prove((c, P, vk, fixed, sigma, gens), [TR.witness(), TR.witness()], inst, 777)
        T = R.Blake2bTranscript(TR.M)
        PP.create_proof_engine(halo2_b200, prm, vk, fixed, sigma, [TR.witness(), TR.witness()], inst, MC.SeededRng("fp", 777, True), T, TR.ZETA, TR.DELTA)
        got = bytes(T.proof)
        assert len(got) == 4160
        assert got == want                                           # bit-identical to the oracle's proof
        earm = PV.EngineArm(halo2_b200, "vesta", 5, *gens)
        try:
            assert PV.verify_proof(earm, vk, got, inst, TR.DELTA)
            bad = bytearray(got)
            bad[2000] ^= 1
            assert not PV.verify_proof(earm, vk, bytes(bad), inst, TR.DELTA)
        finally:
            earm.close()
        assert PV.verify_proof(PV.OracleArm("vesta", 5, *gens), vk, got, inst, TR.DELTA)
        # a witness that breaks a gate: the device prover still runs, the verifier rejects
        T2 = R.Blake2bTranscript(TR.M)
        PP.create_proof_engine(halo2_b200, prm, vk, fixed, sigma, [TR.witness(break_row=5)], [[[2]]], MC.SeededRng("fp", 901, True), T2, TR.ZETA, TR.DELTA)
        assert not PV.verify_proof(PV.OracleArm("vesta", 5, *gens), vk, bytes(T2.proof), [[[2]]], TR.DELTA)
    finally:
        prm.close()
