"""CPU replay of the threshold -> CSR build of simspread.jl_b200/csrc/ss_csr.cu (no GPU needed): the decomposition into
32-row blocks x column segments, the keep-mask words a lane builds from its 32 columns, the row-major order of the
(row, segment) counts whose exclusive scan yields row_ptr, and the mask replay of the fill (a warp prefix over the
popcounts of 32 words, up to four set bits per round).  Pins the *design*; the `-m gpu` tests pin the compiled kernels.
Reference semantics: an edge iff `cutoff(x, alpha, weighted) != 0` (src/core.jl:37-43, src/graphs.jl:10)."""
import numpy as np
import pytest


def keep_edge(x, alpha, weighted):
    return (x >= alpha) & (~weighted | (x != 0.0))


def choose_seg_words(rows, cols, sms=148):
    mask_words = -(-cols // 32)
    row_blocks = -(-rows // 32)
    seg = 32
    while seg > 4 and row_blocks * (-(-mask_words // seg)) < 8 * 64 * sms and -(-mask_words // (seg // 2)) <= 65535:
        seg >>= 1
    return seg, mask_words, row_blocks


def count_pass(S, alpha, weighted):
    rows, cols = S.shape
    seg_words, mask_words, row_blocks = choose_seg_words(rows, cols)
    nseg = -(-mask_words // seg_words)
    mask = np.zeros((rows, mask_words), dtype=np.uint32)
    seg_count = np.zeros((rows, nseg), dtype=np.int64)
    for rb in range(row_blocks):
        for seg in range(nseg):
            for lane in range(32):
                row = rb * 32 + lane
                if row >= rows:
                    continue
                for wd in range(seg * seg_words, min(mask_words, (seg + 1) * seg_words)):
                    bits = 0
                    for j in range(32):
                        c = wd * 32 + j
                        if c < cols and keep_edge(S[row, c], alpha, weighted):
                            bits |= 1 << j
                    mask[row, wd] = bits
                    seg_count[row, seg] += bin(bits).count("1")
    return mask, seg_count, nseg


def fill_pass(S, mask, row_ptr, weighted):
    rows, mask_words = mask.shape
    nnz = int(row_ptr[-1])
    col_idx = np.full(nnz, -1, dtype=np.int64)
    values = np.full(nnz, np.nan)
    for row in range(rows):
        run = int(row_ptr[row])
        for wb in range(0, mask_words, 32):
            words = [int(mask[row, wb + l]) if wb + l < mask_words else 0 for l in range(32)]
            cnt = [bin(w).count("1") for w in words]
            incl = np.cumsum(cnt)
            for lane in range(32):
                pos = run + int(incl[lane]) - cnt[lane]
                bits = words[lane]
                while bits:  # up to four set bits per round
                    taken = []
                    for _ in range(4):
                        if bits:
                            taken.append((bits & -bits).bit_length() - 1)
                            bits &= bits - 1
                    for u, b in enumerate(taken):
                        c = (wb + lane) * 32 + b
                        col_idx[pos + u] = c
                        values[pos + u] = S[row, c]
                    pos += len(taken)
            run += int(incl[31])
    return col_idx, values


@pytest.mark.parametrize("rows,cols", [(1, 1), (7, 5), (70, 130), (33, 1025)])
@pytest.mark.parametrize("weighted", [False, True])
def test_csr_build_replay(rows, cols, weighted):
    rng = np.random.default_rng(rows * 131 + cols)
    S = np.round(rng.random((rows, cols)), 2)
    S.flat[rng.integers(0, S.size, size=max(1, S.size // 20))] = np.nan
    S.flat[rng.integers(0, S.size, size=max(1, S.size // 20))] = -0.0
    for alpha in (0.9, 0.0, -0.5, 1.5):
        mask, seg_count, nseg = count_pass(S, alpha, np.bool_(weighted))
        flat = seg_count.reshape(-1)                       # (row, segment) row-major == CSR order
        offs = np.concatenate(([0], np.cumsum(flat)))
        row_ptr = np.concatenate((offs[:-1][::nseg], offs[-1:]))
        with np.errstate(invalid="ignore"):
            keep = keep_edge(S, alpha, np.bool_(weighted))
        wr, wc = np.nonzero(keep)
        assert np.array_equal(row_ptr, np.concatenate(([0], np.cumsum(np.bincount(wr, minlength=rows)))))
        col_idx, values = fill_pass(S, mask, row_ptr, weighted)
        assert np.array_equal(col_idx, wc)
        assert np.array_equal(values, S[wr, wc])


def test_segment_choice_limits():
    for rows, cols in [(100_000, 20_000), (5_000, 5_000), (300, 300), (32, 4_000_000), (2_000_000, 64)]:
        seg, mask_words, row_blocks = choose_seg_words(rows, cols)
        assert seg in (4, 8, 16, 32) and -(-mask_words // seg) <= 65535
    assert choose_seg_words(100_000, 20_000)[0] == 16     # C4: 3 125 row blocks x 40 segments
