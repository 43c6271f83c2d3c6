"""Jump-ahead polynomials for mt19937 (Haramoto, Matsumoto, Nishimura, Panneton, L'Ecuyer 2008) -> csrc/mt_jump_table.h.

The victim-way stream of the reference is ONE sequential mt19937 stream (torch's CPU generator).  To generate it on
many SMs at once the device needs the generator state J words ahead: with phi(x) the characteristic polynomial of the
word-step recurrence (degree 19937) and g(x) = x^J mod phi(x) = sum g_i x^i,
        s[n + J] = XOR over {i : g_i = 1} of s[n + i]          for every n and every bit of the 32-bit words.
This script derives phi by Berlekamp-Massey from the generator's own output, computes g for J = C * 2^m
(C = 624 * BLOCKS_PER_CHUNK words, m = 0 .. LEVELS-1), VERIFIES every g against sequential generation and writes the
bit maps as a C header.  Pure Python integers are the GF(2)[x] polynomials (bit i = coefficient of x^i).

    python tools/gen_mt_jump.py [--verify-levels N]      (regenerates cdlrm_b200/csrc/mt_jump_table.h)
"""
import argparse
import os
import sys

import numpy as np

N, M = 624, 397
DEG = 19937
BLOCKS_PER_CHUNK = 4096
LEVELS = 12


def mt_blocks(state, nblocks):
    """state: uint32[624] (a full block); returns the next nblocks blocks, untempered, uint32[nblocks, 624]."""
    out = np.empty((nblocks, N), dtype=np.uint32)
    x = state.astype(np.uint32).copy()
    U, L, A = np.uint32(0x80000000), np.uint32(0x7fffffff), np.uint32(0x9908b0df)

    def tw(a, b, c):
        y = (a & U) | (b & L)
        return c ^ (y >> np.uint32(1)) ^ np.where(y & np.uint32(1), A, np.uint32(0))

    for k in range(nblocks):
        x[0:227] = tw(x[0:227], x[1:228], x[397:624])
        x[227:454] = tw(x[227:454], x[228:455], x[0:227])
        x[454:623] = tw(x[454:623], x[455:624], x[227:396])
        x[623] = tw(x[623], x[0], x[396])
        out[k] = x
    return out


def seed_state(seed):
    h = np.empty(N, dtype=np.uint64)
    h[0] = seed & 0xffffffff
    for i in range(1, N):
        h[i] = (1812433253 * (int(h[i - 1]) ^ (int(h[i - 1]) >> 30)) + i) & 0xffffffff
    return h.astype(np.uint32)


def berlekamp_massey(bits):
    """Berlekamp-Massey over GF(2) with a running reversed window (bit j = s_{i-j}).  Returns (L, C), C the
    connection polynomial as an int (bit i = c_i, c_0 = 1): s_n = XOR_{i=1..L} c_i s_{n-i}."""
    C, B = 1, 1
    L, m = 0, 1
    w = 0
    for i, b in enumerate(bits):
        w = (w << 1) | b                       # bit j of w = s_{i-j}
        d = bin(C & w).count("1") & 1          # c_0 s_i + c_1 s_{i-1} + ...
        if d:
            T = C
            C ^= B << m
            if 2 * L <= i:
                L, B, m = i + 1 - L, T, 1
            else:
                m += 1
        else:
            m += 1
        if i >= 2 * DEG + 64:
            w &= (1 << (DEG + 2)) - 1          # keep the window bounded (only the low L+1 bits matter)
    return L, C


def pmulmod(a, b, phi):
    """a * b mod phi in GF(2)[x]."""
    r = 0
    while b:
        low = b & -b
        r ^= a << (low.bit_length() - 1)
        b ^= low
    return pmod(r, phi)


def pmod(a, phi):
    d = phi.bit_length() - 1
    while a.bit_length() - 1 >= d:
        a ^= phi << (a.bit_length() - 1 - d)
    return a


def psqmod(a, phi):
    # squaring in GF(2)[x] spreads the bits: bit i -> bit 2i
    s = int(bin(a)[2:].replace("0", "00").replace("1", "01"), 2) if a else 0
    return pmod(s, phi)


def ppowx(e, phi):
    """x^e mod phi."""
    r = 1
    for bit in bin(e)[2:]:
        r = psqmod(r, phi)
        if bit == "1":
            r = pmod(r << 1, phi)
    return r


def jump_words(seq, g):
    """seq: uint32 sequence s[0 .. DEG + 623]; returns s[J .. J + 623] = XOR_{g_i} s[i .. i + 623]."""
    acc = np.zeros(N, dtype=np.uint32)
    i = 0
    gg = g
    while gg:
        if gg & 1:
            acc ^= seq[i:i + N]
        gg >>= 1
        i += 1
    return acc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--verify-levels", type=int, default=6, help="levels checked against sequential generation")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                  "cdlrm_b200", "csrc", "mt_jump_table.h"))
    a = ap.parse_args()
    st = seed_state(5489)
    blocks = mt_blocks(st, 70)                              # 43 680 words of the sequence after the seed block
    seq = np.concatenate([st, blocks.reshape(-1)])          # s[0 ..]: seed block, then its successors
    # NOTE the seed block is an arbitrary 19968-bit vector; only its top 19937 bits are state, so the recurrence
    # relation is checked on s[624 ..] (a genuine orbit) below
    orbit = seq[N:]
    bits = [int(v) & 1 for v in orbit[:2 * DEG + 200]]
    L, C = berlekamp_massey(bits)
    assert L == DEG, L
    # characteristic polynomial = reciprocal of the connection polynomial: phi(x) = x^L C(1/x)
    phi = sum(((C >> i) & 1) << (L - i) for i in range(L + 1))
    assert phi.bit_length() - 1 == DEG and (phi & 1)
    # sum_i phi_i s[n+i] = 0 on every bit of the words
    idx = [i for i in range(DEG + 1) if (phi >> i) & 1]
    for n0 in (0, 1, 777):
        acc = np.uint32(0)
        for i in idx:
            acc ^= orbit[n0 + i]
        assert acc == 0, "phi does not annihilate the sequence"
    C_words = N * BLOCKS_PER_CHUNK
    polys = []
    g = ppowx(C_words, phi)
    for m in range(LEVELS):
        polys.append(g)
        g = psqmod(g, phi)                                   # x^(2J) = (x^J)^2
    # verification against sequential generation: the block that starts J words after block b0
    b0 = blocks[3]                                           # a genuine block (window s[n .. n+623])
    window = np.concatenate([b0, mt_blocks(b0, 33).reshape(-1)])     # s[n .. n + 624*34)
    for m in range(min(a.verify_levels, LEVELS)):
        jb = BLOCKS_PER_CHUNK << m                           # blocks to skip
        cur, left = b0.copy(), jb
        while left:
            step = min(left, 4096)
            cur = mt_blocks(cur, step)[-1]
            left -= step
        got = jump_words(window, polys[m])
        assert np.array_equal(got[1:], cur[1:]) and (int(got[0]) >> 31) == (int(cur[0]) >> 31), f"level {m} wrong"
        print(f"level {m}: jump by {jb} blocks verified", file=sys.stderr)
    with open(a.out, "w") as f:
        f.write("// GENERATED by tools/gen_mt_jump.py -- do not edit.  Jump-ahead polynomials of mt19937:\n")
        f.write("// g_m(x) = x^(MT_JUMP_CHUNK_WORDS * 2^m) mod phi(x), bit i of word w = coefficient of x^(32 w + i).\n")
        f.write("#pragma once\n#include <stdint.h>\n")
        f.write(f"#define MT_JUMP_BLOCKS_PER_CHUNK {BLOCKS_PER_CHUNK}\n#define MT_JUMP_CHUNK_WORDS {C_words}\n")
        f.write(f"#define MT_JUMP_LEVELS {LEVELS}\n#define MT_JUMP_POLY_WORDS {N}\n")
        f.write("#ifndef MT_JUMP_QUAL\n#define MT_JUMP_QUAL static const\n#endif\n")
        f.write("MT_JUMP_QUAL uint32_t mt_jump_poly[MT_JUMP_LEVELS][MT_JUMP_POLY_WORDS] = {\n")
        for g in polys:
            words = [(g >> (32 * w)) & 0xffffffff for w in range(N)]
            f.write("  {" + ",".join("0x%08xu" % w for w in words) + "},\n")
        f.write("};\n")
    print("wrote", a.out, file=sys.stderr)


if __name__ == "__main__":
    main()
