"""The GEMM work list of a tree level on the CPU (tests/host/worklist_host.cpp drives the CUPPEN_HD functions the device
builder is made of): every element of every (merge, half) problem is covered exactly once, whole tiles come in L2-sized
super-columns with the n index fastest, and the tiles of an under-filled last wave -- or of a level shorter than half a
wave -- are emitted as two 64-column halves each."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def wl():
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "host")], check=True, stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(os.path.join(ROOT, "tests", "host", "_build", "libworklist_host.so"))
    ip = ctypes.POINTER(ctypes.c_int)
    lib.worklist_build.argtypes = [ctypes.c_int, ip, ip, ip, ip, ip] + [ctypes.c_int] * 6 + [ip, ip]
    lib.worklist_build.restype = ctypes.c_int
    return lib


def build(wl, n1, n2, k, ktop, kbot, p0=0, width=1 << 20, BMN=128, split_grid=148, supercol_mb=48, cap=1 << 16):
    nd = len(n1)
    arr = lambda v: np.ascontiguousarray(v, dtype=np.int32)
    a = [arr(x) for x in (n1, n2, k, ktop, kbot)]
    tiles = np.zeros(4 * cap, dtype=np.int32)
    probs = np.zeros(6 * nd, dtype=np.int32)
    ip = ctypes.POINTER(ctypes.c_int)
    cnt = wl.worklist_build(nd, *[x.ctypes.data_as(ip) for x in a], p0, width, BMN, split_grid, supercol_mb, cap,
                            tiles.ctypes.data_as(ip), probs.ctypes.data_as(ip))
    assert cnt >= 0
    return tiles[:4 * cnt].reshape(cnt, 4), probs.reshape(2 * nd, 3)


def check_cover(tiles, probs, BMN):
    """every (row, column) of every problem exactly once; returns the number of plain tiles the list stands for"""
    plain = 0
    for p, (M, N, K) in enumerate(probs.tolist()):
        mine = tiles[tiles[:, 0] == p]
        if M == 0 or N == 0:
            assert len(mine) == 0
            continue
        ntm, ntn = -(-M // BMN), -(-N // BMN)
        cover = np.zeros((ntm, -(-N // 64) if BMN == 128 else ntn), dtype=np.int32)      # 64-column granules
        for _, m0, n0, wdt in mine.tolist():
            assert m0 % BMN == 0 and 0 <= m0 < M and 0 <= n0 and wdt in (64, BMN)
            if BMN == 128:
                assert n0 % 64 == 0
                for g in range(n0 // 64, (n0 + wdt) // 64):
                    if g < cover.shape[1]:
                        cover[m0 // BMN, g] += 1
                    else:
                        assert n0 + 64 > N or g * 64 >= N       # a granule past a ragged right edge computes nothing
            else:
                cover[m0 // BMN, n0 // BMN] += 1
        assert (cover == 1).all(), (p, M, N, cover)
        plain += ntm * ntn
    return plain


@pytest.mark.parametrize("seed", range(12))
def test_random_levels_are_covered_once(wl, seed):
    rng = np.random.default_rng(seed)
    nd = int(rng.integers(1, 9))
    n1 = rng.integers(130, 2000, size=nd)
    n2 = n1 + rng.integers(0, 2, size=nd)
    k = np.array([int(rng.integers(0, a + b + 1)) for a, b in zip(n1, n2)])
    ktop = np.minimum(k, rng.integers(0, 2000, size=nd))
    kbot = np.minimum(k, rng.integers(0, 2000, size=nd))
    grid = int(rng.choice([0, 8, 148]))
    tiles, probs = build(wl, n1, n2, k, ktop, kbot, split_grid=grid, supercol_mb=int(rng.choice([1, 48])))
    plain = check_cover(tiles, probs, 128)
    halves = int((tiles[:, 3] == 64).sum())
    assert halves % 2 == 0 and len(tiles) == plain + halves // 2
    tail = plain % grid if grid else 0
    assert halves == (2 * tail if grid and 2 * tail <= grid else 0)
    # the half tiles are the last entries, in pairs (left half, right half) of the same tile
    if halves:
        h = tiles[len(tiles) - halves:]
        assert (h[:, 3] == 64).all() and (tiles[:len(tiles) - halves, 3] == 128).all()
        assert (h[0::2, :2] == h[1::2, :2]).all() and (h[1::2, 2] == h[0::2, 2] + 64).all()


def test_headline_shapes(wl):
    """GOE n=16384 on 8 ranks: 1024 rows of each half per rank at the top merge, k = 13473 -> 2 x 8 x 106 = 1696 tiles =
    11 waves of 148 + 68: the last 68 tiles become 136 half tiles.  On one rank: 13568 tiles, 100 left over: no split."""
    tiles, probs = build(wl, [1024], [1024], [13473], [6736], [6737])
    assert probs.tolist() == [[1024, 13473, 6736], [1024, 13473, 6737]]
    assert check_cover(tiles, probs, 128) == 1696 and len(tiles) == 1696 + 68 and int((tiles[:, 3] == 64).sum()) == 136
    tiles, probs = build(wl, [8192], [8192], [13473], [6736], [6737])
    assert check_cover(tiles, probs, 128) == 13568 and (tiles[:, 3] == 128).all()
    # super-columns: 48 MiB / (6736 * 8 * 128 B) = 7 n-tiles wide, n fastest inside a super-column
    first = tiles[tiles[:, 0] == 0][:16]
    assert first[:7, 2].tolist() == [128 * j for j in range(7)] and (first[:7, 1] == 0).all() and first[7].tolist()[1:3] == [128, 0]


def test_short_level_is_split_entirely(wl):
    """`-s 1 -n 4096`, top merge: k = 211 -> 2 x 16 x 2 = 64 tiles on 148 CTAs: every tile as two halves."""
    tiles, probs = build(wl, [2048], [2048], [211], [120], [100])
    assert check_cover(tiles, probs, 128) == 64 and len(tiles) == 128 and (tiles[:, 3] == 64).all()
    # a panel of root columns (p0, width) and the small-tile kernel (no split there)
    tiles, probs = build(wl, [300, 301], [300, 300], [500, 77], [260, 40], [250, 40], p0=128, width=256, BMN=64, split_grid=0)
    assert probs[:, 1].tolist() == [256, 256, 0, 0] and check_cover(tiles, probs, 64) == len(tiles)
