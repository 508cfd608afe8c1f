"""Python emulation of fill_dialogue's chunking (csrc/graphify.cu) against the numpy oracle."""
import numpy as np, sys
sys.path.insert(0, '.')
from oracle import graph_np

def deg_prefix(L, A, Bk, k):
    t1 = L - A; a = min(k, t1)
    s1 = a * A + a * (a - 1) // 2 + (k - a) * (L - 1)
    b = max(k - 1 - Bk, 0)
    return s1 + k - b * (b + 1) // 2

def emu(L, wp, wf, spk, n):
    m = L - 1
    P = m if (wp < 0 or wp > m) else wp
    F = m if (wf < 0 or wf > m) else wf
    E = L * (P + F + 1) - P * (P + 1) // 2 - F * (F + 1) // 2
    col = np.full(E, -1); et = np.full(E, -1); inv = np.zeros(E, np.float32)
    tcol = np.full(E, -1); teid = np.full(E, -1)
    for ps in range(2):
        A, Bk = (P, F) if ps == 0 else (F, P)
        for k0 in range(0, L, 32):
            myS = [deg_prefix(L, A, Bk, min(k0 + l, L)) for l in range(32)]
            kend = min(k0 + 32, L)
            end = deg_prefix(L, A, Bk, kend)
            xb = myS[0]
            while xb < end:
                rowEnd = [end if l == 31 else myS[l + 1] for l in range(32)]
                rowok = [k0 + l < kend for l in range(32)]
                fits = []
                if any(rowok[l] and myS[l] == xb for l in range(32)):
                    fits = [l for l in range(32) if rowok[l] and myS[l] >= xb and rowEnd[l] <= xb + 32]
                if fits:
                    xe = rowEnd[max(fits)]; whole = True
                else:
                    cont = [l for l in range(32) if rowok[l] and myS[l] <= xb < rowEnd[l]]
                    re = rowEnd[cont[0]] if cont else rowEnd[31]
                    xe = min(xb + 32, re); whole = False
                keys = {}
                lanes = []
                for l in range(32):
                    x = xb + l; valid = x < xe; xs = x if valid else xb
                    lo, hi = 0, 31
                    for _ in range(5):
                        mid = (lo + hi + 1) >> 1
                        if myS[mid] <= xs: lo = mid
                        else: hi = mid - 1
                    row = k0 + lo; rlo = max(row - Bk, 0); other = rlo + (xs - myS[lo])
                    j, k = (other, row) if ps == 0 else (row, other)
                    sj, sk = spk[j], spk[k]
                    ty = (sj * n + sk) * 2 + (1 if j >= k else 0)
                    key = (lo, sj, j >= k) if valid else ('x', l)
                    keys[key] = keys.get(key, 0) + 1
                    lanes.append((valid, x, j, k, sj, ty, key, rlo))
                for (valid, x, j, k, sj, ty, key, rlo) in lanes:
                    if not valid: continue
                    if ps == 0:
                        assert col[x] == -1
                        col[x] = j; et[x] = ty
                        if whole: c = keys[key]
                        else:
                            rhi = min(k + P, L - 1)
                            c = sum(1 for jj in range(rlo, rhi + 1) if spk[jj] == sj and ((jj >= k) == (j >= k)))
                        inv[x] = np.float32(1.0) / np.float32(c)
                    else:
                        assert tcol[x] == -1
                        tcol[x] = k
                        klo = max(k - F, 0)
                        teid[x] = deg_prefix(L, P, F, k) + (j - klo)
                xb = xe
    return col, et, inv, tcol, teid

rng = np.random.default_rng(0)
cases = 0
for (wp, wf, n) in [(5, 5, 2), (10, 10, 2), (0, 0, 1), (-1, -1, 2), (-1, 3, 3), (4, -1, 9), (2, 7, 2), (200, 200, 2), (20, 15, 2), (31, 0, 2), (16, 16, 3)]:
    for L in list(range(1, 40)) + [63, 64, 65, 97, 110, 150]:
        spk = rng.integers(0, n, size=L)
        b = graph_np.batch_graphify_np(np.array([L]), spk[None, :], wp, wf, n)
        col, et, inv, tcol, teid = emu(L, wp, wf, spk, n)
        assert np.array_equal(col, b['col']), (L, wp, wf)
        assert np.array_equal(et, b['etype'])
        assert np.array_equal(inv, b['inv_cnt']), (L, wp, wf)
        assert np.array_equal(tcol, b['t_col'])
        assert np.array_equal(teid, b['t_eid'])
        cases += 1
print("ok", cases)
