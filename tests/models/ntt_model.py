"""Pure-Python model of csrc/ntt.cu's pass planning / index maps (CPU-side design check).

Mirrors ntt_run + ntt_pass_kernel statement by statement with Python ints so that the
tile/bit-reversal/twiddle-index logic can be validated against the naive DFT without a GPU."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..', 'oracle'))
from bn254 import R_MOD


def brev(x, bits):
    return int(format(x, '0%db' % bits)[::-1], 2) if bits else 0


def build_twiddles(w, logn):
    n = 1 << logn
    flat = [pow(w, i, R_MOD) for i in range(n // 2)]
    tab = [0] * (n - 1)
    for e in range(n - 1):
        t = (e + 1).bit_length() - 1
        j = e + 1 - (1 << t)
        tab[e] = flat[j << (logn - 1 - t)]
    return tab


def ntt_run(a, w, logn, max_s, LOGC, n_in=None, n_out=None, in_scale=None, out_scale=None, mod3=False):
    n = 1 << logn
    tw = build_twiddles(w, logn)
    n_in = n if n_in is None else n_in
    n_out = n if n_out is None else n_out
    npass = max(1, (logn + max_s - 1) // max_s)
    bits_left, hi = logn, logn
    cur_in = list(a) + [None] * (n - len(a))
    tmp = [None] * n
    out = [None] * n_out
    for p in range(npass):
        S = (bits_left + (npass - p) - 1) // (npass - p)
        lo = hi - S
        first, last = p == 0, p == npass - 1
        src = cur_in if first else tmp
        contiguous = last
        logc = (LOGC if logn - S >= LOGC else logn - S) if last else LOGC
        if not last:
            assert lo >= logc
        R, C = 1 << S, 1 << logc
        tiles = 1 << (logn - S - logc)
        dst_writes = []
        for tile in range(tiles):
            sm = [None] * (R * C)
            mid = 0
            if not contiguous:
                midbits = lo - logc
                mid = tile & ((1 << midbits) - 1)
                top = tile >> midbits
                base = (top << (lo + S)) | (mid << logc)
                for e in range(R * C):
                    c, j = e & (C - 1), e >> logc
                    idx = base | (j << lo) | c
                    v = src[idx] if (not first or idx < n_in) else 0
                    if first and in_scale and idx % 3 and idx < n_in:
                        v = v * in_scale[idx % 3] % R_MOD
                    sm[e] = v
            else:
                base = tile << S
                for e in range(R * C):
                    j, c = e & (R - 1), e >> S
                    idx = ((c << (logn - logc)) | base | j) if logc else (base | j)
                    v = src[idx] if (not first or idx < n_in) else 0
                    if first and in_scale and idx % 3 and idx < n_in:
                        v = v * in_scale[idx % 3] % R_MOD
                    sm[e] = v
            for m in range(S - 1, -1, -1):
                half = 1 << m
                t = lo + m
                T = (1 << t) - 1
                for b in range((R * C) >> 1):
                    if not contiguous:
                        c, jb = b & (C - 1), b >> logc
                    else:
                        jb, c = b & ((R >> 1) - 1), b >> (S - 1)
                    jl = jb & (half - 1)
                    ja = ((jb >> m) << (m + 1)) | jl
                    if not contiguous:
                        ea = (ja << logc) | c
                        eb = ea + (half << logc)
                        widx = (jl << lo) | (mid << logc) | c
                    else:
                        ea = (c << S) | ja
                        eb = ea + half
                        widx = jl
                    x, y = sm[ea], sm[eb]
                    s, d = (x + y) % R_MOD, (x - y) % R_MOD
                    if widx:
                        d = d * tw[T + widx] % R_MOD
                    sm[ea], sm[eb] = s, d
            if not contiguous:
                for e in range(R * C):
                    c, j = e & (C - 1), e >> logc
                    dst_writes.append((base | (j << lo) | c, sm[e]))
            else:
                restbits = logn - S - logc
                rest_rev = brev(tile, restbits)
                for e in range(R * C):
                    cr, jr = e & (C - 1), e >> logc
                    j, c = brev(jr, S), brev(cr, logc)
                    pos = (jr << (logn - S)) | (rest_rev << logc) | cr
                    if pos >= n_out:
                        continue
                    v = sm[(c << S) | j]
                    if out_scale:
                        v = v * out_scale[pos % 3 if mod3 else 0] % R_MOD
                    dst_writes.append((pos, v))
        dst = out if last else tmp
        for i, v in dst_writes:
            dst[i] = v
        hi -= S
        bits_left -= S
    return out
