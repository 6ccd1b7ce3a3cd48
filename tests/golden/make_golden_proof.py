#!/usr/bin/env python3
"""Extracts the reference's golden PROOFS and the pinned verifying keys they verify under into
tests/golden/golden_proofs.json.gz.  Run in the build container only (reads /root/reference); the archive travels.

Sources (relative to /root/reference):
  halo2_proofs/tests/plonk_api_proof.bin     the proof `plonk_api` checks with verify_proof (halo2_proofs/tests/plonk_api.rs:462-476):
                                             two instances of the test circuit, public input [2] each (:398, :435, :472)
  halo2_proofs/tests/plonk_api.rs:586-985    format!("{:#?}", vk.pinned()) of its key (k = 5)
  halo2_gadgets/src/test_circuits/circuit_data/{vk_*.rdata, proof_*.bin}
                                             fifteen k = 11 circuits (ECC chip, Sinsemilla, Merkle, range checks): the same pinned form
                                             and a stored proof each, verified by test_against_stored_circuit with one empty instance
                                             (halo2_gadgets/src/test_circuits/test_utils.rs:49-58, :68-111)
A pinned key holds everything the verifier reads -- domain, gate polynomials, queries, permutation columns, lookups, fixed and
permutation commitments -- and its compact {:?} form is the string whose BLAKE2b hash enters every transcript
(halo2_proofs/src/plonk.rs:75-86).  Identical key texts are stored once.
"""
import glob
import gzip
import hashlib
import json
import os

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_proofs.json.gz")


def main():
    keys, cases = {}, []

    def key_id(text):
        h = hashlib.sha256(text.encode()).hexdigest()[:16]
        keys[h] = text
        return h

    api = open(os.path.join(REF, "halo2_proofs/tests/plonk_api.rs")).read()
    start = api.index('r#####"PinnedVerificationKey') + len('r#####"')
    cases.append({"name": "plonk_api", "source": "halo2_proofs/tests/plonk_api.rs:586-985 + halo2_proofs/tests/plonk_api_proof.bin",
                  "curve": "vesta", "key": key_id(api[start:api.index('"#####', start)]), "instances": [[["0x2"]], [["0x2"]]],
                  "proof_hex": open(os.path.join(REF, "halo2_proofs/tests/plonk_api_proof.bin"), "rb").read().hex()})
    d = os.path.join(REF, "halo2_gadgets/src/test_circuits/circuit_data")
    for pf in sorted(glob.glob(os.path.join(d, "proof_*.bin"))):
        name = os.path.basename(pf)[len("proof_"):-len(".bin")]
        text = open(os.path.join(d, f"vk_{name}.rdata")).read().replace("\r\n", "\n")
        cases.append({"name": name, "source": f"halo2_gadgets/src/test_circuits/circuit_data/{{vk_{name}.rdata, proof_{name}.bin}}",
                      "curve": "vesta", "key": key_id(text), "instances": [[]], "proof_hex": open(pf, "rb").read().hex()})
    blob = json.dumps({"keys": keys, "cases": cases}, sort_keys=True).encode()
    with open(OUT, "wb") as f:                                   # mtime = 0: the archive is reproducible byte for byte
        with gzip.GzipFile(fileobj=f, mode="wb", mtime=0, compresslevel=9) as z:
            z.write(blob)
    print("wrote", OUT, os.path.getsize(OUT), "bytes:", len(cases), "proofs,", len(keys), "distinct keys,", len(blob), "bytes of JSON")


if __name__ == "__main__":
    main()
