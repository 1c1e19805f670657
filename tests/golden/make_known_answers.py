#!/usr/bin/env python
"""Writes tests/golden/known_answers.json: the known-answer anchors for the ciphertext-evaluation path
that can be stated WITHOUT running SEAL -- the reference ships no golden vectors and SEAL
cannot be built offline, so these are recomputed here from the published definitions with plain Python
integers (independent of oracle/ and of the product), then checked against the oracle by
tests/test_oracle.py.

  primes     CoeffModulus::Create / get_primes: candidates 2^bits - k*2N + 1, descending; per bit size the
             chain takes the smallest-of-the-top-k first and the largest last (the special prime)
  plain      PlainModulus::Batching(N, 20): largest 20-bit prime = 1 mod 2N   (1032193 for N=8192 is the value
             SEAL's own examples print -- the one externally published anchor)
  psi        minimal primitive 2N-th root of unity per prime (NTTTables)
  naf        util/numth.h naf()
  galois     GaloisTool::get_elt_from_step / get_elts_all
Usage: python tests/golden/make_known_answers.py
"""
import json
import os
import random


def is_prime(n):
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def get_primes(factor, bits, count):
    out, v = [], (1 << bits) - factor + 1
    while len(out) < count and v > (1 << (bits - 1)):
        if is_prime(v):
            out.append(v)
        v -= factor
    assert len(out) == count
    return out


def coeff_modulus_create(N, bit_sizes):
    need = {}
    for b in bit_sizes:
        need[b] = need.get(b, 0) + 1
    pools = {b: get_primes(2 * N, b, c) for b, c in need.items()}
    return [pools[b].pop() for b in bit_sizes]


def minimal_primitive_root(two_n, q):
    # any primitive 2N-th root generates all of them by odd powers; take the minimum
    rng = random.Random(1)
    while True:
        g = pow(rng.randrange(2, q - 1), (q - 1) // two_n, q)
        if pow(g, two_n // 2, q) == q - 1:
            break
    best, g2, cur = g, g * g % q, g
    for _ in range(two_n // 2):
        best = min(best, cur)
        cur = cur * g2 % q
    return best


def naf(value):
    sign, v, out, i = (-1 if value < 0 else 1), abs(value), [], 0
    while v:
        z = 2 - (v & 3) if v & 1 else 0
        v = (v - z) >> 1
        if z:
            out.append(sign * z * (1 << i))
        i += 1
    return out


def elt_from_step(step, N):
    m = 2 * N
    if step == 0:
        return m - 1
    pos = abs(step)
    e = (N // 2 - pos) if step < 0 else pos
    return pow(3, e, m)


def elts_all(N):
    m, out = 2 * N, []
    logn = N.bit_length() - 1
    pos, neg = 3, pow(3, -1, m)
    for _ in range(logn - 1):
        out += [pos, neg]
        pos, neg = pos * pos % m, neg * neg % m
    return out + [m - 1]


CONFIGS = {
    "C1": (8192, [60, 40, 60]), "C2": (8192, [60, 45, 60]), "C3": (16384, [60, 40, 60]),
    "C4": (16384, [60, 45, 45, 45, 45, 45, 60]), "C5": (32768, [60, 45, 45, 45, 45, 45, 60]),
}

if __name__ == "__main__":
    ka = {"chains": {}, "plain_modulus": {}, "psi": {}, "naf": {}, "galois_elt_from_step": {}, "galois_elts_all": {}}
    for name, (N, bits) in CONFIGS.items():
        chain = coeff_modulus_create(N, bits)
        ka["chains"][name] = {"N": N, "bits": bits, "moduli": [hex(q) for q in chain]}
        for q in chain:
            ka["psi"][f"{N}:{hex(q)}"] = minimal_primitive_root(2 * N, q)
    for N in (8192, 16384):
        ka["plain_modulus"][str(N)] = get_primes(2 * N, 20, 1)[0]
    for v in (100, -3, 455, 1, -1, 7, 99, -8191):
        ka["naf"][str(v)] = naf(v)
    for N in (8192, 32768):
        ka["galois_elt_from_step"][str(N)] = {str(s): elt_from_step(s, N) for s in (0, 1, -1, 2, 64, -64, 100, N // 2 - 1)}
        ka["galois_elts_all"][str(N)] = elts_all(N)
    # anchors quoted in SURVEY.md §8(c) (computed there independently): fail loudly if this script disagrees
    assert ka["chains"]["C2"]["moduli"] == ["0xffffffffffe8001", "0x1ffffff8c001", "0xfffffffffffc001"]
    assert ka["chains"]["C1"]["moduli"][1] == "0xfffffdc001"
    assert ka["chains"]["C4"]["moduli"][1:6] == ["0x1fffffde8001", "0x1fffffe28001", "0x1fffffe58001", "0x1fffffee8001", "0x1ffffff18001"]
    assert ka["chains"]["C5"]["moduli"][0] == "0xfffffffff840001" and ka["chains"]["C5"]["moduli"][-1] == "0xffffffffffc0001"
    assert ka["plain_modulus"] == {"8192": 1032193, "16384": 786433}
    assert ka["psi"]["8192:0xffffffffffe8001"] == 100406242475323 and ka["psi"]["8192:0x1ffffff8c001"] == 2229466015
    assert ka["psi"]["16384:0xffffe80001"] == 42618759
    assert ka["naf"]["100"] == [4, -32, 128] and ka["naf"]["-3"] == [1, -4] and ka["naf"]["455"] == [-1, 8, -64, 512]
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "known_answers.json"), "w") as f:
        json.dump(ka, f, indent=1)
    print("wrote known_answers.json")
