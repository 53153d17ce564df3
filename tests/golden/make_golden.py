#!/usr/bin/env python3
"""Regenerate the golden fixtures for the unordered c64 known-answer test.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes, next to this script,
  unordered_n2048_dif4_b32_input.f64   4096 little-endian f64 (re, im interleaved)
  unordered_n2048_dif4_b32_target.f64  4096 little-endian f64 (re, im interleaved)

* target = the 2048-entry vector hard-coded in the reference's own test
  `test_equivalency` (src/unordered.rs:1176-9396), extracted textually.
* input  = what that test feeds the plan: rand 0.8 `StdRng::seed_from_u64(0)` followed by
  `gen_range(0.0..1.0)` for re then im of each element (src/unordered.rs:1180-1188).
  rand/rand_chacha are dev-dependencies that are not vendored in the reference tree
  (Cargo.toml:33), so the generator is restated here from their published algorithm:
  PCG32 seed expansion -> ChaCha12 block function -> next_u64 = lo | hi << 32 ->
  f64::from_bits(u >> 12 | 0x3FF << 52) - 1.0.   

Both blobs are checked against the sha256 digests recorded in SURVEY.md section 8c.
"""
import hashlib
import os
import re
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/unordered.rs"
M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF

SHA_INPUT = "efb841a11a3d8325f6fac757487d37282c4ad37cb2801f48afd52a227304ebb1"
SHA_TARGET = "fd92a48d1f5d05edba9530a270931a0dcfed0d41a6b4cd8e77c2cc5d955e1cba"


def rotl32(x, r):
    return ((x << r) | (x >> (32 - r))) & M32


def seed_from_u64(state):
    """rand_core 0.6 SeedableRng::seed_from_u64: PCG32 (XSH-RR) fills the 32-byte seed."""
    MUL, INC = 6364136223846793005, 11634580027462260723
    words = []
    for _ in range(8):
        state = (state * MUL + INC) & M64
        xorshifted = (((state >> 18) ^ state) >> 27) & M32
        rot = state >> 59
        words.append(((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & M32)
    return words


def chacha12_block(key, counter):
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key) + [
        counter & M32, (counter >> 32) & M32, 0, 0]
    x = list(st)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & M32; x[d] = rotl32(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & M32; x[b] = rotl32(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & M32; x[d] = rotl32(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & M32; x[b] = rotl32(x[b] ^ x[c], 7)

    for _ in range(6):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(a + b) & M32 for a, b in zip(x, st)]


def std_rng_unit_f64(seed, count):
    key = seed_from_u64(seed)
    words = []
    blk = 0
    while len(words) < 2 * count:
        words += chacha12_block(key, blk)
        blk += 1
    out = []
    for j in range(count):
        u = words[2 * j] | (words[2 * j + 1] << 32)
        bits = (u >> 12) | (0x3FF << 52)
        out.append(struct.unpack("<d", struct.pack("<Q", bits))[0] - 1.0)
    return out


def extract_target():
    src = open(REF).read()
    a = src.index("let target:")
    b = src.index("assert_eq!", a)
    vals = re.findall(r"(?:re|im):\s*(-?[0-9.eE+-]+),", src[a:b])
    assert len(vals) == 4096, len(vals)
    return [float(v) for v in vals]


def main():
    inp = std_rng_unit_f64(0, 4096)
    blob_in = struct.pack("<4096d", *inp)
    assert hashlib.sha256(blob_in).hexdigest() == SHA_INPUT, "input digest mismatch"
    assert inp[0] == 0.7311134158637045 and inp[1] == 0.773460184353238

    if not os.path.exists(REF):
        sys.exit("reference tree not present; fixtures can only be regenerated in the build container")
    tgt = extract_target()
    blob_t = struct.pack("<4096d", *tgt)
    assert hashlib.sha256(blob_t).hexdigest() == SHA_TARGET, "target digest mismatch"

    with open(os.path.join(HERE, "unordered_n2048_dif4_b32_input.f64"), "wb") as f:
        f.write(blob_in)
    with open(os.path.join(HERE, "unordered_n2048_dif4_b32_target.f64"), "wb") as f:
        f.write(blob_t)
    print("wrote golden fixtures; sum re = %.16g, sum im = %.16g" % (sum(inp[0::2]), sum(inp[1::2])))


if __name__ == "__main__":
    main()
