"""Offline generator of the mt19937 jump-ahead table used by pyg_lib_b200/csrc/mt19937_jump.cuh.

mt19937's raw (untempered) word stream is linear over GF(2): the window W_t = raw[t .. t+623] advances
by a fixed linear map A.  With phi(x) the minimal polynomial of A on the 19937-bit state and
g_p(x) = x^(p*S) mod (x * phi(x))  (the extra factor x makes the low 31 bits of the window's first word —
which are not part of the state but ARE an output word here — come out right),
    raw[t + p*S + j] = XOR_{i : bit i of g_p} raw[t + i + j],   j = 0..623,
so CTA p can start generating chunk p of the stream after one pass over ~20k already generated words.

Writes pyg_lib_b200/csrc/mt19937_jump.bin:  u32 magic, u32 S, u32 P, u32 words_per_poly(624), then
(P-1) polynomials g_1..g_{P-1}, each 624 little-endian u32 words (bit i of the polynomial = bit i%32 of word i/32).
Pure Python (big ints as GF(2)[x] elements); ~1 minute.
"""
import os
import struct
import sys
import time

import numpy as np

N, M = 624, 397
S_DEFAULT, P_DEFAULT = 1 << 17, 64


def raw_stream(seed: int, n_words: int) -> np.ndarray:
    st = np.zeros(N, dtype=np.uint64)
    st[0] = seed & 0xffffffff
    for j in range(1, N):
        st[j] = (1812433253 * (int(st[j - 1]) ^ (int(st[j - 1]) >> 30)) + j) & 0xffffffff
    raw = np.zeros(n_words + N, dtype=np.uint32)
    raw[:N] = st.astype(np.uint32)
    m = N
    while m < len(raw):  # raw[m] = raw[m-227] ^ T(raw[m-624], raw[m-623]), 227 words at a time
        n = min(N - M, len(raw) - m)
        u, v = raw[m - N:m - N + n], raw[m - N + 1:m - N + 1 + n]
        y = (u & np.uint32(0x80000000)) | (v & np.uint32(0x7fffffff))
        raw[m:m + n] = raw[m - (N - M):m - (N - M) + n] ^ (y >> np.uint32(1)) ^ np.where(v & np.uint32(1), np.uint32(0x9908b0df), np.uint32(0))
        m += n
    return raw


def berlekamp_massey(bits):
    """Connection polynomial C (int, bit i = c_i, c_0 = 1) and length L of the shortest LFSR for `bits`."""
    n = len(bits)
    C, B, L, m = 1, 1, 0, 1
    s = 0  # s holds bits[0..k] reversed so that the discrepancy is parity(C & window)
    for k in range(n):
        s = (s << 1) | bits[k]          # bit i of s = bits[k - i]
        d = bin(C & s).count('1') & 1   # sum_i c_i * bits[k-i]
        if d == 0:
            m += 1
        elif 2 * L <= k:
            T = C
            C ^= B << m
            L, B, m = k + 1 - L, T, 1
        else:
            C ^= B << m
            m += 1
    return C, L


def polymulmod(a: int, b: int, mod: int, deg: int) -> int:
    r = 0
    while a:
        low = a & -a
        r ^= b << (low.bit_length() - 1)
        a ^= low
    return polymod(r, mod, deg)


def polymod(r: int, mod: int, deg: int) -> int:
    while r.bit_length() - 1 >= deg:
        r ^= mod << (r.bit_length() - 1 - deg)
    return r


def polypowx(e: int, mod: int, deg: int) -> int:
    """x^e mod `mod`."""
    result, base = 1, 2
    while e:
        if e & 1:
            result = polymulmod(result, base, mod, deg)
        base = polymulmod(base, base, mod, deg)
        e >>= 1
    return result


def main(S=S_DEFAULT, P=P_DEFAULT):
    t0 = time.time()
    raw = raw_stream(5489, 2 * 19937 + 2000)
    bits = [int(x) & 1 for x in raw[:2 * 19937 + 200]]
    C, L = berlekamp_massey(bits)
    # The bit-0 sequence starts with a transient (the low 31 bits of state word 0 are not part of the
    # 19937-bit state), so the shortest LFSR has length 19938 = degree of x * phi(x): exactly the modulus
    # that reproduces whole windows including those bits.
    assert L == 19938 and not (C >> L) & 1, L
    # recurrence s[n] = sum_{i=1..L} c_i s[n-i]  ->  P(x) = sum_i c_i x^(L-i)  (c_0 = 1 -> x^L)
    mod = 0
    for i in range(L + 1):
        if (C >> i) & 1:
            mod |= 1 << (L - i)
    deg = L
    assert mod & 1 == 0 and (mod >> 1).bit_length() - 1 == 19937
    print(f'x*phi(x): degree {deg}, weight {bin(mod).count("1")}  ({time.time() - t0:.1f}s)')
    g1 = polypowx(S, mod, deg)
    polys, g = [], g1
    for p in range(1, P):
        polys.append(g)
        if p + 1 < P:
            g = polymulmod(g, g1, mod, deg)
    print(f'{len(polys)} jump polynomials for stride {S}  ({time.time() - t0:.1f}s)')
    # self-check on a different seed, all 32 bits of every window word (incl. word 0)
    chk = raw_stream(123456789, 3 * S + 40000)
    for p in (1, 2, 3):
        gp = polys[p - 1]
        idx = [i for i in range(deg) if (gp >> i) & 1]
        win = np.zeros(N, dtype=np.uint32)
        for i in idx:
            win ^= chk[i:i + N]
        assert np.array_equal(win, chk[p * S:p * S + N]), f'jump {p} wrong'
    print('self-check ok')
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'pyg_lib_b200', 'csrc', 'mt19937_jump.bin')
    with open(out, 'wb') as f:
        f.write(struct.pack('<4I', 0x4a54364d, S, P, N))
        for gp in polys:
            f.write(gp.to_bytes(N * 4, 'little'))
    print('wrote', out, os.path.getsize(out), 'bytes')


if __name__ == '__main__':
    main(*(int(a) for a in sys.argv[1:3]))
