#!/usr/bin/env python3
"""Extracts the known-answer data the reference's own tests hold for the MSM+FFT hot path
into tests/golden/reference_goldens.json.  Run in the build container only (it reads
/root/reference, which does not exist on the GPU box); the JSON is what travels.

Sources (relative to /root/reference):
  halo2_proofs/tests/plonk_api.rs:591-592     field moduli
  halo2_proofs/tests/plonk_api.rs:594-596     k=5 domain omega
  halo2_proofs/tests/plonk_api.rs:958-982     k=5 commit_lagrange outputs (Vesta points)
  halo2_gadgets/src/test_circuits/circuit_data/vk_lookup_range_check.rdata   k=11 omega + commitments
  halo2_poseidon/src/{fp,fq}.rs               Poseidon round constants + MDS (from_raw limbs)
  halo2_poseidon/src/test_vectors.rs          permutation known-answer vectors (32-byte LE reprs)
"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")


def read(p):
    with open(os.path.join(REF, p)) as f:
        return f.read()


def parse_vk(text):
    d = {}
    d["base_modulus"] = re.search(r'base_modulus: "(0x[0-9a-f]+)"', text).group(1)
    d["scalar_modulus"] = re.search(r'scalar_modulus: "(0x[0-9a-f]+)"', text).group(1)
    d["k"] = int(re.search(r"\bk: (\d+),", text).group(1))
    d["extended_k"] = int(re.search(r"extended_k: (\d+),", text).group(1))
    d["omega"] = re.search(r"omega: (0x[0-9a-f]+)", text).group(1)
    fixed = text[text.index("fixed_commitments"):text.index("permutation: VerifyingKey")]
    perm = text[text.index("permutation: VerifyingKey"):]
    pt = re.compile(r"\((0x[0-9a-f]{64}), (0x[0-9a-f]{64})\)")
    d["fixed_commitments"] = [list(m) for m in pt.findall(fixed)]
    d["permutation_commitments"] = [list(m) for m in pt.findall(perm)]
    return d


def parse_from_raw(block):
    """All from_raw([l0,l1,l2,l3]) literals in order -> ints."""
    vals = []
    for m in re.finditer(r"from_raw\(\[\s*(0x[0-9a-f_]+),\s*(0x[0-9a-f_]+),\s*(0x[0-9a-f_]+),\s*(0x[0-9a-f_]+),?\s*\]\)", block):
        limbs = [int(x.replace("_", ""), 16) for x in m.groups()]
        vals.append(hex(limbs[0] | limbs[1] << 64 | limbs[2] << 128 | limbs[3] << 192))
    return vals


def parse_poseidon_consts(path):
    t = read(path)
    i_rc = t.index("const ROUND_CONSTANTS")
    i_mds = t.index("const MDS:")
    i_inv = t.index("const MDS_INV")
    rc = parse_from_raw(t[i_rc:i_mds])
    mds = parse_from_raw(t[i_mds:i_inv])
    assert len(rc) == 192 and len(mds) == 9, (len(rc), len(mds))
    return {"round_constants": rc, "mds": mds}


def parse_permute_vectors(text, mod):
    i0 = text.index("pub mod %s" % mod)
    sec = text[i0:]
    sec = sec[sec.index("pub fn permute()"):sec.index("pub fn hash()")]
    byts = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{2})\b", sec)]
    assert len(byts) % 192 == 0
    out = []
    for i in range(0, len(byts), 192):
        blk = byts[i:i + 192]
        els = [hex(int.from_bytes(bytes(blk[j:j + 32]), "little")) for j in range(0, 192, 32)]
        out.append({"initial_state": els[:3], "final_state": els[3:]})
    return out


def main():
    g = {}
    api = read("halo2_proofs/tests/plonk_api.rs")
    vk5 = api[api.index('r#####"PinnedVerificationKey'):]
    g["vk_plonk_api_k5"] = parse_vk(vk5)
    g["vk_lookup_range_check_k11"] = parse_vk(read("halo2_gadgets/src/test_circuits/circuit_data/vk_lookup_range_check.rdata"))
    tv = read("halo2_poseidon/src/test_vectors.rs")
    g["poseidon"] = {
        "fp": dict(parse_poseidon_consts("halo2_poseidon/src/fp.rs"), permute=parse_permute_vectors(tv, "fp")),
        "fq": dict(parse_poseidon_consts("halo2_poseidon/src/fq.rs"), permute=parse_permute_vectors(tv, "fq")),
        "full_rounds": 8, "partial_rounds": 56,  # halo2_poseidon/src/p128pow5t3.rs:12-24
    }
    with open(OUT, "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;",
          len(g["poseidon"]["fp"]["permute"]), "+", len(g["poseidon"]["fq"]["permute"]), "permute vectors")


if __name__ == "__main__":
    main()
