"""Pure-Python model of the control structure of csrc/msm.cu + msm_tail.cu (CPU-side design check): chunked
shared-memory digit sort, segmented accumulation, partial folding, both bucket level-1 forms, mixed-basis batches.

Group elements are modelled as integers mod r (identity 0, add = +), which keeps every
index / run / slot decision of the kernels while making the expected answer trivial:
MSM(s, g) = sum s_i * g_i mod r.  Mirrors msm_run and its kernels statement by statement."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..', 'oracle'))
from bn254 import R_MOD

INVALID = 0xffffffff


def digit(s, w, c, carry):
    v = ((s >> (w * c)) & ((1 << c) - 1)) + carry
    if v > (1 << (c - 1)):
        return v - (1 << c), 1
    return v, 0


def serial_reduce(level0, keys, vals, pts_in, L, table, K, buckets, nthreads):
    pkeys = [None] * (2 * nthreads)
    ppts = [None] * (2 * nthreads)
    for t in range(nthreads):
        start = t * K
        if start >= L:
            pkeys[2 * t] = pkeys[2 * t + 1] = INVALID
            continue
        end = min(start + K, L)
        cur = keys[start]
        if not level0 and cur == INVALID:
            pkeys[2 * t] = pkeys[2 * t + 1] = INVALID
            continue
        acc, nruns = 0, 0
        for e in range(start, end):
            k = keys[e]
            if not level0 and k == INVALID:
                break
            if k != cur:
                if nruns == 0:
                    pkeys[2 * t], ppts[2 * t] = cur, acc
                else:
                    assert buckets[cur] == 0, "double direct write"
                    buckets[cur] = acc
                nruns += 1
                cur, acc = k, 0
            if level0:
                v = vals[e]
                tb, alt_tb, amask, nbk = table
                p = (alt_tb if (amask >> (k // nbk)) & 1 else tb)[v & 0x7fffffff]
                if v >> 31:
                    p = -p
                acc = (acc + p) % R_MOD
            else:
                acc = (acc + pts_in[e]) % R_MOD
        if nruns == 0:
            pkeys[2 * t], ppts[2 * t] = cur, acc
            pkeys[2 * t + 1], ppts[2 * t + 1] = cur, 0
        else:
            pkeys[2 * t + 1], ppts[2 * t + 1] = cur, acc
    return pkeys, ppts


def warp_reduce(keys, pts, n_in, buckets, nwarps, final):
    pkeys = [None] * (2 * nwarps)
    ppts = [None] * (2 * nwarps)
    for g in range(nwarps):
        key = [keys[g * 32 + l] if g * 32 + l < n_in else INVALID for l in range(32)]
        acc = [pts[g * 32 + l] if key[l] != INVALID else 0 for l in range(32)]
        d = 1
        while d < 32:
            new = list(acc)
            for l in range(32):
                if l + d < 32 and key[l + d] == key[l] and key[l] != INVALID:
                    new[l] = (acc[l] + acc[l + d]) % R_MOD
            acc = new
            d <<= 1
        for l in range(32):
            head = l == 0 or key[l - 1] != key[l]
            if final:
                if head and key[l] != INVALID:
                    assert buckets[key[l]] == 0
                    buckets[key[l]] = acc[l]
                continue
            if key[l] == INVALID:
                if l == 0:
                    pkeys[2 * g] = INVALID
                if l == 31:
                    pkeys[2 * g + 1] = INVALID
                continue
            if not head:
                continue
            touch_end = key[31] == key[l]
            if l == 0:
                pkeys[2 * g], ppts[2 * g] = key[l], acc[l]
                if touch_end:
                    pkeys[2 * g + 1], ppts[2 * g + 1] = key[l], 0
            elif touch_end:
                pkeys[2 * g + 1], ppts[2 * g + 1] = key[l], acc[l]
            else:
                assert buckets[key[l]] == 0
                buckets[key[l]] = acc[l]
    return pkeys, ppts


def warp_weighted(x):
    x = list(x)
    d = 1
    while d < 32:
        x = [(x[l] + x[l + d]) % R_MOD if l + d < 32 else x[l] for l in range(32)]
        d <<= 1
    s = x[0]
    y = [x[l] if l >= 1 else 0 for l in range(32)]
    d = 16
    while d >= 1:
        y = [(y[l] + y[l + d]) % R_MOD if l < d else y[l] for l in range(32)]
        d >>= 1
    return s, y[0]


def l1_serial_form(x):
    """msm_bucket_l1_kernel: four lanes per group, eight buckets each with the running-sum trick, two merge steps"""
    run, acc = [0] * 4, [0] * 4
    for sub in range(4):
        for l in range(7, 0, -1):
            run[sub] = (run[sub] + x[8 * sub + l]) % R_MOD
            acc[sub] = (acc[sub] + run[sub]) % R_MOD
        run[sub] = (run[sub] + x[8 * sub]) % R_MOD
    for step in range(2):
        d = 1 << step
        for sub in range(0, 4, 2 * d):
            so, to = run[sub + d], acc[sub + d]
            run[sub] = (run[sub] + so) % R_MOD
            acc[sub] = (acc[sub] + to + (so << (3 + step))) % R_MOD
    return run[0], acc[0]


def msm_model(scalars_list, g, c, K0=None, serial_l1_threshold=8192, J=None, alt=None, alt_mask=0, l1_serial=None):
    n = len(g)
    W = (255 + c - 1) // c
    NB = 1 << (c - 1)
    M = len(scalars_list)
    def window_table(base):
        t = [0] * (W * n)
        for w in range(W):
            for i in range(n):
                t[w * n + i] = (base[i] << (c * w)) % R_MOD
        return t
    table = window_table(g)
    alt_table = window_table(alt) if alt is not None else None   # mixed-basis batch: MSM m uses `alt` when bit m of alt_mask
    cnt = M * NB
    # msm_digits_kernel<false>: CTA (chunk j, MSM m) counts its digits per bucket -> H[m][j][b]
    if J is None:
        J = max(1, min((592 + M - 1) // M, (n + max(256, NB // 8) - 1) // max(256, NB // 8)))
    chunk = (n + J - 1) // J
    J = (n + chunk - 1) // chunk
    H = [[[0] * NB for _ in range(J)] for _ in range(M)]

    def digits_of(m, i):
        s = scalars_list[m][i]
        if s == 0:
            return
        carry = 0
        for w in range(W):
            d, carry = digit(s, w, c, carry)
            if d != 0:
                yield abs(d) - 1, (w * n + i) | ((1 if d < 0 else 0) << 31)
        assert carry == 0
    for m in range(M):
        for j in range(J):
            for i in range(j * chunk, min((j + 1) * chunk, n)):
                for b, _ in digits_of(m, i):
                    H[m][j][b] += 1
    # msm_hist_total_kernel + scan + msm_hist_offsets_kernel
    hist = [sum(H[m][j][b] for j in range(J)) for m in range(M) for b in range(NB)]
    offsets = [0] * (cnt + 1)
    for i in range(cnt):
        offsets[i + 1] = offsets[i] + hist[i]
    for m in range(M):
        for b in range(NB):
            run = offsets[m * NB + b]
            for j in range(J):
                v = H[m][j][b]
                H[m][j][b] = run
                run += v
    L = offsets[cnt]
    keys, vals = [None] * L, [None] * L
    # msm_digits_kernel<true>: the CTA reloads its slice of H as cursors
    for m in range(M):
        for j in range(J):
            cur = H[m][j]
            for i in range(j * chunk, min((j + 1) * chunk, n)):
                for b, v in digits_of(m, i):
                    assert keys[cur[b]] is None, "two digits scattered to one slot"
                    keys[cur[b]], vals[cur[b]] = m * NB + b, v
                    cur[b] += 1
    assert all(k is not None for k in keys)
    assert all(keys[e] <= keys[e + 1] for e in range(L - 1)), "entries not sorted by bucket"
    buckets = [0] * cnt
    L_max = n * W * M
    if K0 is None:
        K0 = 32 if L_max >= (2 << 20) else 16
    T0 = (L_max + K0 - 1) // K0
    pk, pp = serial_reduce(True, keys, vals, None, L, (table, alt_table, alt_mask, NB), K0, buckets, T0)
    slots = 2 * T0
    if slots > serial_l1_threshold:
        T1 = (slots + 7) // 8                                            # MSM_LEVEL1_K
        pk, pp = serial_reduce(False, pk, None, pp, slots, None, 8, buckets, T1)
        slots = 2 * T1
    while True:
        nwarps = (slots + 31) // 32
        fin = nwarps == 1
        pk, pp = warp_reduce(pk, pp, slots, buckets, nwarps, fin)
        if fin:
            break
        slots = 2 * nwarps
    n1 = (NB + 31) // 32
    out = []
    for m in range(M):
        s1, t1 = [], []
        serial = l1_serial if l1_serial is not None else (n1 * M > 2 * 592)
        for gidx in range(n1):
            x = [buckets[m * NB + gidx * 32 + l] if gidx * 32 + l < NB else 0 for l in range(32)]
            s, t = l1_serial_form(x) if serial else warp_weighted(x)
            s1.append(s); t1.append(t)
        nw = (n1 + 31) // 32
        S2, T2, U = [], [], []
        for w in range(nw):
            x = [s1[w * 32 + l] if w * 32 + l < n1 else 0 for l in range(32)]
            tt = [t1[w * 32 + l] if w * 32 + l < n1 else 0 for l in range(32)]
            s, t = warp_weighted(x)
            S2.append(s); T2.append(t); U.append(sum(tt) % R_MOD)
        x = [S2[l] if l < nw else 0 for l in range(32)]
        S, t3 = warp_weighted(x)
        r = (t3 * 32 + sum(T2)) % R_MOD
        r = (r * 32 + sum(U)) % R_MOD
        r = (r + S) % R_MOD
        out.append(r)
    return out
