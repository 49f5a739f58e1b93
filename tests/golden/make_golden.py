#!/usr/bin/env python3
"""Regenerates tests/golden/* from the reference tree (run in the build container; needs /root/reference).

Copies the reference's golden STARK seal and records the known-answer vectors its own unit tests assert:
  proof.bin    <- risc0/zkp/src/verify/proof.bin            (verify_v3_stark_proof, verify/mod.rs:713-726)
  bigint_*.blob <- risc0/bigint2/src/field/*.blob            (bigint2 programs the bigint-ecall test guests execute)
  kats.json    <- literals asserted in:
      risc0/core/src/field/baby_bear.rs:893-894 (5^1000), :815-853 (FpExt linear)
      risc0/zkp/src/core/hash/poseidon2/mod.rs:330-351 (permutation), :354-401 (hash 32 / 17 elems)
      risc0/zkp/src/prove/merkle.rs:161-172 (Fiat-Shamir RNG)
      risc0/zkp/src/hal/cpu.rs:726-733 (SHA hash_rows of a 1x16 zero matrix)
      risc0/zkp/src/merkle.rs:73-102 (MerkleTreeParams)
The literal values are parsed out of those files so a change in the reference shows up as a diff here.
"""
import json, os, re, shutil

REF = os.environ.get("R0_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

def read(p):
    return open(os.path.join(REF, p)).read()

def ints(s):
    return [int(x, 16) if x.lower().startswith("0x") else int(x) for x in re.findall(r"0x[0-9a-fA-F]+|\d+", s)]

shutil.copyfile(os.path.join(REF, "risc0/zkp/src/verify/proof.bin"), os.path.join(HERE, "proof.bin"))
# bigint2 programs (blob = header + bibc 'nondet' program + verify program + constants) that the bigint-ecall guests of
# tests/test_preflight.py / tests/test_gpu_witgen.py run: risc0/bigint2/src/field/<name>.blob -> bigint_<name>.blob
for name in ("modmul_256", "modinv_256", "modsub_256", "modadd_256", "modmul_384", "extfield_deg2_mul_256", "modmul_4096"):
    shutil.copyfile(os.path.join(REF, "risc0/bigint2/src/field/%s.blob" % name), os.path.join(HERE, "bigint_%s.blob" % name))

bb = read("risc0/core/src/field/baby_bear.rs")
m = re.search(r"Elem::new\(5\)\.pow\(1000\),\s*Elem::new\((\d+)\)", bb)
pow_kat = int(m.group(1))
lin = re.search(r"pub fn linear\(\) \{(.*?)\n    \}\n", bb, re.S).group(1)
vals = [int(x) for x in re.findall(r"Elem::new\((\d+)\)", lin)]
assert len(vals) == 20
p2 = read("risc0/zkp/src/core/hash/poseidon2/mod.rs")
perm = re.search(r"fn poseidon2_test_vectors\(\) \{(.*?)\n    \}\n", p2, re.S).group(1)
arrs = re.findall(r"\[\s*((?:0x[0-9a-fA-F]+,?\s*)+)\]", perm)
perm_in, perm_out = ints(arrs[0]), ints(arrs[1])
assert len(perm_in) == 24 and len(perm_out) == 24
def hash_kat(fn):
    body = re.search(r"fn " + fn + r"\(\) \{(.*?)\n    \}\n", p2, re.S).group(1)
    inp = ints(re.search(r"baby_bear_array!\[(.*?)\];", body, re.S).group(1))
    goal = [int(x, 16) for x in re.findall(r"from\(0x([0-9a-fA-F]+)_u32\)", body)]
    assert len(goal) == 8
    return {"input": inp, "digest_normal_form": goal}
mk = read("risc0/zkp/src/prove/merkle.rs")
rng = re.search(r"fn basic_read_iop\(\) \{(.*?)\n    \}\n", mk, re.S).group(1)
rng_vals = [int(x) for x in re.findall(r"as_u32\(\), (\d+)\)", rng)]
assert len(rng_vals) == 2
cpu = read("risc0/zkp/src/hal/cpu.rs")
sha = re.search(r'do_hash_rows\(\s*1,\s*16,\s*&\["([0-9a-f]{64})"\]', cpu).group(1)
kats = {
    "fp_5_pow_1000": pow_kat,
    "fpext_linear": {"x": vals[0:4], "c0": vals[4:8], "c1": vals[8:12], "x_mul_c1": vals[12:16], "c0_plus_x_mul_c1": vals[16:20]},
    "poseidon2_perm": {"input": perm_in, "output": perm_out},
    "poseidon2_hash32": hash_kat("hash_elem_slice_compare_golden"),
    "poseidon2_hash17": hash_kat("hash_elem_slice_compare_golden_unaligned"),
    "poseidon2_rng": {"after_commit_zero": rng_vals[0], "after_commit_x2345678": rng_vals[1]},
    "sha_hash_rows_1x16_zero": sha,
    "merkle_params": [
        {"rows": 1024, "cols": 1234, "queries": 50, "layers": 10, "top_layer": 5, "top_size": 32},
        {"rows": 2048, "cols": 31337, "queries": 128, "layers": 11, "top_layer": 7, "top_size": 128},
    ],
    "golden_seal": {"file": "proof.bin", "po2": 12, "circuit": "HelloCircuit (verify/mod.rs:614-727)", "hash": "poseidon2"},
}
json.dump(kats, open(os.path.join(HERE, "kats.json"), "w"), indent=1)
print("wrote proof.bin, kats.json")
