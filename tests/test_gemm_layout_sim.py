"""CPU replay of the index math of simspread.jl_b200/csrc/ss_gemm.cu (no GPU needed).

The DMMA kernel reads its operands from TMA SWIZZLE_128B tiles with permuted k / row assignments
so that every LDS.128 is bank-conflict free.  This test re-derives, in NumPy, (i) where TMA puts
every element, (ii) which shared-memory address every lane reads, (iii) what mma.m8n8k4 computes
from those fragments and (iv) where the epilogue stores each accumulator -- and checks the result
against A @ B, plus the conflict-freeness of each warp-wide LDS.128.  It pins the *design*; the
`-m gpu` tests pin the compiled kernel."""
import numpy as np
import pytest

BM = BN = 128
BK = 16
WARPS_N = 4
WM, WN = 64, 32
MT, NT = WM // 8, WN // 8


def swz(o):
    return o ^ (((o >> 7) & 7) << 4)


def rho(g):
    return (g >> 1) | ((g & 1) << 2)


def tma_kmajor(tile):
    """tile[row, k] (128 x 16) -> 16 KB smem image, box {16 k, 128 rows}, 128B swizzle."""
    sm = np.full(128 * 16, np.nan)
    for r in range(128):
        for k in range(16):
            sm[swz(r * 128 + k * 8) // 8] = tile[r, k]
    return sm


def tma_mmajor(tile, band=0, split=1):
    """tile[m, k] (128 x 16) -> 8 boxes {16 m, 16 k} of 2 KB each.  A row band (split 2 / 4) loads only its
    128 / split rows, into the boxes the whole tile would have put them (producer: a_off + b * 2048)."""
    sm = np.full(128 * 16, np.nan)
    per = 8 // split
    for b in range(band * per, (band + 1) * per):
        for k in range(16):
            for mi in range(16):
                o = k * 128 + mi * 8
                sm[(b * 2048 + swz(o)) // 8] = tile[b * 16 + mi, k]
    return sm


def lds128(sm, addr, log):
    assert addr % 16 == 0
    log.append(addr)
    return sm[addr // 8], sm[addr // 8 + 1]


def check_conflicts(addrs):
    """addrs: 32 lane byte addresses of one LDS.128; quarter-warps must hit 8 distinct 16B bank
    groups (or identical addresses)."""
    for q in range(4):
        seen = {}
        for a in addrs[8 * q:8 * q + 8]:
            bank = (a >> 4) & 7
            assert seen.setdefault(bank, a) == a, f"bank conflict in quarter {q}: {addrs}"


def mma_8x8x4(acc, a, b):
    """acc[lane][2], a[lane], b[lane] with lane = 4g+t: A[g][t], B[t][g(col)], C[g][2t+e]."""
    A = a.reshape(8, 4)
    B = b.reshape(8, 4).T  # B[t][col]
    C = A @ B  # 8 x 8
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        acc[lane][0] += C[g, 2 * t]
        acc[lane][1] += C[g, 2 * t + 1]


def run_warp(sA, sB, warp, a_mmajor, MT=MT, band=0):
    """MT = 8: whole tile; MT = 4 / 2: row band `band` of 64 / 32 rows (consume_unit<A_MMAJOR, MT_>)."""
    m_warp = band * (2 * MT * 8) + (warp // WARPS_N) * (MT * 8)
    n_warp = (warp % WARPS_N) * WN
    acc = np.zeros((MT, NT, 32, 2))
    for h in range(2):
        bf = np.zeros((NT, 32, 2))
        for j in range(NT):
            log = []
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                rg = rho(g)
                off = (n_warp + rg) * 128 + (((t + 4 * h) ^ rg) << 4)
                bf[j, lane] = lds128(sB, off + j * 8 * 128, log)
            check_conflicts(log)
        if a_mmajor:
            for s2 in range(2):
                s = 2 * h + s2
                af = np.zeros((MT // 2, 32, 2))
                for b in range(MT // 2):
                    log = []
                    for lane in range(32):
                        g, t = lane >> 2, lane & 3
                        k = 2 * t + (s & 1) + 8 * (s >> 1)
                        off = (m_warp >> 4) * 2048 + k * 128 + ((g ^ (k & 7)) << 4)
                        af[b, lane] = lds128(sA, off + b * 2048, log)
                    check_conflicts(log)
                for b in range(MT // 2):
                    for j in range(NT):
                        mma_8x8x4(acc[2 * b, j], af[b, :, 0], bf[j, :, s2])
                        mma_8x8x4(acc[2 * b + 1, j], af[b, :, 1], bf[j, :, s2])
        else:
            af = np.zeros((MT, 32, 2))
            for i in range(MT):
                log = []
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    rg = rho(g)
                    off = (m_warp + rg) * 128 + (((t + 4 * h) ^ rg) << 4)
                    af[i, lane] = lds128(sA, off + i * 8 * 128, log)
                check_conflicts(log)
            for s2 in range(2):
                for i in range(MT):
                    for j in range(NT):
                        mma_8x8x4(acc[i, j], af[i, :, s2], bf[j, :, s2])
    return acc


def epilogue(C, acc, warp, a_mmajor, MT=MT, band=0):
    m_warp = band * (2 * MT * 8) + (warp // WARPS_N) * (MT * 8)
    n_warp = (warp % WARPS_N) * WN
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for j in range(NT):
            for e in range(2):
                col = n_warp + 8 * j + t + 4 * e
                if a_mmajor:
                    for b in range(MT // 2):
                        row = m_warp + 16 * b + 2 * g
                        C[row, col] += acc[2 * b, j, lane, e]
                        C[row + 1, col] += acc[2 * b + 1, j, lane, e]
                else:
                    for i in range(MT):
                        C[m_warp + 8 * i + rho(g), col] += acc[i, j, lane, e]


@pytest.mark.parametrize("a_mmajor", [True, False])
def test_fragment_layout_reproduces_gemm(a_mmajor):
    rng = np.random.default_rng(3)
    K = 32
    A = rng.integers(-4, 5, size=(BM, K)).astype(float)
    B = rng.integers(-4, 5, size=(K, BN)).astype(float)
    C = np.zeros((BM, BN))
    for kb in range(K // BK):
        At = A[:, kb * BK:(kb + 1) * BK]
        Bt = B[kb * BK:(kb + 1) * BK, :].T  # [n, k]
        sA = tma_mmajor(At) if a_mmajor else tma_kmajor(At)
        sB = tma_kmajor(Bt)
        for warp in range(8):
            acc = run_warp(sA, sB, warp, a_mmajor)
            epilogue(C, acc, warp, a_mmajor)
    assert np.array_equal(C, A @ B)


@pytest.mark.parametrize("a_mmajor", [True, False])
@pytest.mark.parametrize("split", [2, 4])
def test_row_bands_reproduce_the_tile(a_mmajor, split):
    """A tile of a partial wave is computed as `split` row bands by the same 2 x 4 warp grid with a shorter warp
    tile; together the bands must write every entry of the tile exactly once, from the same k sequence."""
    rng = np.random.default_rng(5)
    K = 32
    A = rng.integers(-4, 5, size=(BM, K)).astype(float)
    B = rng.integers(-4, 5, size=(K, BN)).astype(float)
    C = np.zeros((BM, BN))
    writes = np.zeros((BM, BN), dtype=int)
    mt = 8 // split
    for band in range(split):
        for kb in range(K // BK):
            At = A[:, kb * BK:(kb + 1) * BK]
            Bt = B[kb * BK:(kb + 1) * BK, :].T
            sA = tma_mmajor(At, band, split) if a_mmajor else tma_kmajor(At)  # k-major A: the whole box is loaded
            sB = tma_kmajor(Bt)
            for warp in range(8):
                acc = run_warp(sA, sB, warp, a_mmajor, MT=mt, band=band)
                assert not np.isnan(acc).any()  # a band never reads rows the producer did not load
                epilogue(C, acc, warp, a_mmajor, MT=mt, band=band)
                if kb == 0:
                    W = np.zeros((BM, BN))
                    epilogue(W, np.ones_like(acc), warp, a_mmajor, MT=mt, band=band)
                    writes += (W != 0)
    assert np.array_equal(C, A @ B) and np.all(writes == 1)


def choose_split(M, N, sms=148, allow=4):
    """launch_gemm_f64_one's choice of the unit list: (full_tiles, split, reg_rows, short_bands, cost)."""
    tiles_m, tiles_n = -(-M // 128), -(-N // 128)
    total = tiles_m * tiles_n
    best, split, full, reg_rows, short_bands = float(-(-total // sms)), 1, total, tiles_m, 0
    valid_last = M - (tiles_m - 1) * 128
    s = 2
    while s <= allow and s <= 4:
        band_rows = 128 // s
        sb = -(-valid_last // band_rows)
        sb = sb if sb < s else 0
        rr = tiles_m - 1 if sb else tiles_m
        reg, shorts = rr * tiles_n, (tiles_n if sb else 0)
        rem = reg % sms
        for j in (0, 1):
            f = reg - rem - j * sms
            if f < 0 or (reg - f) * s + shorts * sb == 0:
                continue
            cost = f // sms + (-(-((reg - f) * s + shorts * sb) // sms)) / s * (1.04 if s == 2 else 1.08)
            if cost < best * 0.98:
                best, split, full, reg_rows, short_bands = cost, s, f, rr, sb
        s *= 2
    return full, split, reg_rows, short_bands, best


def coord(tile, tiles_m, tiles_n, GROUP_M=16):
    gs = GROUP_M * tiles_n
    gid = tile // gs
    first = gid * GROUP_M
    gm = min(tiles_m - first, GROUP_M)
    r = tile - gid * gs
    return first + r % gm, r // gm


def test_unit_enumeration_covers_every_valid_row_once():
    """Host-side choice of the split + unit_of: whole waves as tiles, the trailing tiles as row bands, a short last row
    tile as the bands that hold rows.  Every (row, column tile) with a valid row is computed by exactly one unit."""
    sms = 148
    shapes = [(1, 1), (600, 1250), (1650, 1660), (1790, 1900), (2400, 1000), (2040, 2048), (5000, 2000), (5000, 5000),
              (2000, 2000), (12500, 50000), (100000, 50000), (130, 130), (128 * 148, 128), (128 * 37 + 1, 128 * 4)]
    for M, N in shapes:
        tiles_m, tiles_n = -(-M // 128), -(-N // 128)
        full, split, reg_rows, sb, cost = choose_split(M, N, sms)
        reg_tiles = reg_rows * tiles_n
        units = full + (reg_tiles - full) * split + (tiles_n * sb if sb else 0)
        assert full % sms == 0 or split == 1
        cover = np.zeros((tiles_m * 128, tiles_n), dtype=int)
        for u in range(units):
            if u < full:
                tm, tn = coord(u, reg_rows, tiles_n)
                band, sp = 0, 1
            else:
                v = u - full
                nreg = (reg_tiles - full) * split
                if v < nreg:
                    tm, tn = coord(full + v // split, reg_rows, tiles_n)
                    band, sp = v % split, split
                else:
                    w = v - nreg
                    tm, tn, band, sp = tiles_m - 1, w // sb, w % sb, split
            rows = 128 // sp
            cover[tm * 128 + band * rows: tm * 128 + (band + 1) * rows, tn] += 1
        assert np.all(cover[:M] == 1), (M, N)
        assert cost <= -(-(tiles_m * tiles_n) // sms)
    assert choose_split(5000, 2000)[:4] == (592, 4, 39, 1)   # C3: 624 regular tiles + 16 tiles with 8 valid rows
    assert choose_split(2000, 2000)[1] == 4                   # 1 wave + 108 tiles -> everything in quarter tiles
    assert choose_split(100000, 50000)[1] == 1                # C4: 2 066 waves stay whole tiles
    assert choose_split(12500, 50000)[1] == 1


def test_tile_rasterisation_is_a_bijection():
    GROUP_M = 16

    def coord(tile, tiles_m, tiles_n):
        gs = GROUP_M * tiles_n
        gid = tile // gs
        first = gid * GROUP_M
        gm = min(tiles_m - first, GROUP_M)
        r = tile - gid * gs
        return first + r % gm, r // gm

    for tm, tn in [(1, 1), (1, 6), (7, 3), (16, 5), (37, 11), (782, 391)]:
        seen = {coord(t, tm, tn) for t in range(tm * tn)}
        assert len(seen) == tm * tn
        assert all(0 <= a < tm and 0 <= b < tn for a, b in seen)
