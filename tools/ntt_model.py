#!/usr/bin/env python3
"""numpy model of the four-step / radix-16 NTT decomposition used by risc0_b200/csrc/ntt.cu.

Not shipped, not a test dependency: it exists so the index math (sub-NTT steps, inter-step twiddles, pass split,
expand-by-4 with skipped layers, zk shift) can be validated against the oracle on a CPU box before the CUDA
transcription runs on a GPU. Run: python tools/ntt_model.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import oracle_lib as O  # noqa: E402

P = O.P
L = O.lib()
ROU = {0: [int(O.decode(L.orc_rou_rev(k))) for k in range(28)], 1: [int(O.decode(L.orc_rou_fwd(k))) for k in range(28)]}


def brev(i, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def tw(d, lg, e):
    """w_{2^lg}^e in direction d (0 = REV/inverse, 1 = FWD)"""
    return pow(ROU[d][lg], e, P)


def steps_of(m):
    s = []
    while m > 0:
        a = min(4, m)
        s.append(a)
        m -= a
    return s


def radix_dif(v, a, d=0):
    """in-register 2^a-point DIF with constant twiddles; v: list of 2^a python ints"""
    n = 1 << a
    for s in range(a, 0, -1):
        half = 1 << (s - 1)
        for b in range(0, n, 2 * half):
            for i in range(half):
                x, y = v[b + i], v[b + i + half]
                v[b + i] = (x + y) % P
                v[b + i + half] = (x - y) * tw(d, s, i) % P
    return v


def radix_dit(v, a, skip=0, d=1):
    n = 1 << a
    for s in range(skip + 1, a + 1):
        half = 1 << (s - 1)
        for b in range(0, n, 2 * half):
            for i in range(half):
                x, y = v[b + i], v[b + i + half] * tw(d, s, i) % P
                v[b + i] = (x + y) % P
                v[b + i + half] = (x - y) % P
    return v


def standalone_dif(t, m):
    """t: list of 2^m ints, natural in -> bit-reversed out, twiddles ROU_REV, no scaling"""
    rem = m
    for a in steps_of(m):
        # sub-problems: contiguous blocks of 2^rem; within each, i = (i1: top a bits, r: low rem-a bits)
        sub = 1 << rem
        stride = 1 << (rem - a)
        for base in range(0, 1 << m, sub):
            for r in range(stride):
                v = [t[base + j * stride + r] for j in range(1 << a)]
                v = radix_dif(v, a)
                for j in range(1 << a):
                    if rem - a > 0:
                        v[j] = v[j] * tw(0, rem, r * brev(j, a)) % P
                    t[base + j * stride + r] = v[j]
        rem -= a
    return t


def standalone_dit(t, m, skip=0):
    """bit-reversed in -> natural out, twiddles ROU_FWD; first `skip` layers skipped"""
    steps = steps_of(m)[::-1]  # lowest bits first; steps_of gives e.g. [4,4,2] top-first -> low step is 2
    # the CUDA kernel wants the LOW step to be >= skip; reorder so that the largest steps come first from the bottom
    steps = sorted(steps, reverse=True)
    done = 0
    first = True
    for a in steps:
        sub = 1 << (done + a)
        stride = 1 << done
        for base in range(0, 1 << m, sub):
            for r in range(stride):
                v = [t[base + j * stride + r] for j in range(1 << a)]
                if done > 0:
                    for j in range(1 << a):
                        v[j] = v[j] * tw(1, done + a, brev(j, a) * r) % P
                v = radix_dit(v, a, skip if first else 0)
                for j in range(1 << a):
                    t[base + j * stride + r] = v[j]
        done += a
        first = False
    return t


def split(k):
    if k <= 12:
        return 0, k
    k2 = (k + 1) // 2
    return k - k2, k2


def model_intt(col, zk):
    n = len(col)
    k = n.bit_length() - 1
    k1, k2 = split(k)
    x = [int(v) for v in O.decode(col)]
    if k1:
        n2 = 1 << k2
        for Lo in range(n2):
            t = standalone_dif([x[H * n2 + Lo] for H in range(1 << k1)], k1)
            for Hs in range(1 << k1):
                x[Hs * n2 + Lo] = t[Hs] * tw(0, k, Lo * brev(Hs, k1)) % P
    ninv = pow(n, P - 2, P)
    for Hs in range(1 << k1):
        t = standalone_dif(x[Hs << k2:(Hs + 1) << k2], k2)
        for Ls in range(1 << k2):
            p = (Hs << k2) + Ls
            s = ninv * (pow(3, brev(p, k), P) if zk else 1) % P
            x[p] = t[Ls] * s % P
    return O.encode(np.array(x, dtype=np.uint64))


def model_lde(col, eb):
    n_in = len(col)
    k = (n_in << eb).bit_length() - 1
    k1, k2 = split(k)
    xin = [int(v) for v in O.decode(col)]
    x = [0] * (1 << k)
    for Hs in range(1 << k1):
        tile_in = xin[Hs << (k2 - eb):(Hs + 1) << (k2 - eb)]
        t = [tile_in[i >> eb] for i in range(1 << k2)]
        t = standalone_dit(t, k2, skip=eb)
        for kl in range(1 << k2):
            v = t[kl]
            if k1:
                v = v * tw(1, k, brev(Hs, k1) * kl) % P
            x[(Hs << k2) + kl] = v
    if k1:
        n2 = 1 << k2
        for kl in range(n2):
            t = standalone_dit([x[H * n2 + kl] for H in range(1 << k1)], k1)
            for H in range(1 << k1):
                x[H * n2 + kl] = t[H]
    return O.encode(np.array(x, dtype=np.uint64))


if __name__ == "__main__":
    rng = np.random.default_rng(5)
    for k in (1, 2, 3, 5, 9, 10, 13, 14):
        col = O.rand_elems(rng, 1 << k)
        for zk in (False, True):
            want = O.batch_interpolate_ntt(col, 1)
            if zk:
                want = O.zk_shift(want, 1)
            got = model_intt(col, zk)
            assert np.array_equal(want, got), ("intt", k, zk)
        for eb in (0, 2):
            want = O.batch_expand_into_evaluate_ntt(col, 1, eb)
            got = model_lde(col, eb)
            assert np.array_equal(want, got), ("lde", k, eb)
        print("k=%d ok" % k)
    print("ntt model matches oracle")
