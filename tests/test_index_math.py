"""CPU emulation of the index arithmetic the kernels rely on (no GPU, no compute calls): the multiply-high division used for
token -> pixel and item -> (row, column) maps (csrc/swin_elem.cu, csrc/nhwc_concat.cu: struct FastDiv), the concat kernel's
column -> source lookup, the 128-byte swizzle of the tcgen05 tiles (csrc/tc.cuh: sw128_offset) and the compact block-diagonal
P / dS tile layout of the attention backward (csrc/swin_attn_tc.cu: BCfg)."""
import random


def fastdiv_init(div):
    l = 0
    while (1 << l) < div:
        l += 1
    m = (((1 << 32) * ((1 << l) - div)) // div + 1) & 0xFFFFFFFF
    return m, min(l, 1), max(l - 1, 0)


def fastdiv(n, m, s1, s2):
    t = (m * n) >> 32                       # __umulhi
    return ((t + (((n - t) & 0xFFFFFFFF) >> s1)) & 0xFFFFFFFF) >> s2


def test_fastdiv_is_exact_for_every_divisor_the_kernels_use():
    rng = random.Random(0)
    divisors = list(range(1, 300)) + [384, 400, 420, 576, 640, 1600, 1764, 6400, 8400, 25600, 102400, 409600, 1 << 20, (1 << 31) - 1]
    for d in divisors:
        m, s1, s2 = fastdiv_init(d)
        ns = [0, 1, d - 1, d, d + 1, 2 * d - 1, (1 << 31) - 1, (1 << 32) - 1] + [rng.randrange(1 << 32) for _ in range(200)]
        for n in ns:
            assert fastdiv(n, m, s1, s2) == n // d, (n, d)


def test_concat_column_to_source_lookup():
    # nhwc_concat_kernel: j = number of k in 1..n-1 with col >= first[k]; unit = one 16-byte vector
    for widths in [(4, 4, 4), (8, 16), (2,) * 8, (8, 10), (1,), (3, 1, 5, 2)]:
        first, acc = [], 0
        for w in widths:
            first.append(acc)
            acc += w
        for col in range(acc):
            j = sum(1 for k in range(1, len(widths)) if col >= first[k])
            assert first[j] <= col < first[j] + widths[j]


def sw128_offset(row, chunk):
    return row * 128 + ((chunk ^ (row & 7)) << 4)


def test_sw128_swizzle_is_a_permutation_of_each_1kb_atom():
    for atom in range(4):
        seen = set()
        for r in range(atom * 8, atom * 8 + 8):
            for c in range(8):
                off = sw128_offset(r, c)
                assert atom * 1024 <= off < (atom + 1) * 1024 and off % 16 == 0
                seen.add(off)
        assert len(seen) == 64


def test_compact_block_diagonal_tile_layout():
    # attention backward: P / dS of a window pair stored as [item 0 rows: 8 KB][64 shared zero rows: 8 KB][item 1 rows: 8 KB];
    # key block kb starts at kb * 8 KB and spans 128 rows x 128 B: block 0 = (item 0, zeros), block 1 = (zeros, item 1)
    KBS = 64 * 128
    for r in (0, 1):                       # item
        for i in range(64):                # row inside the item
            for c in range(8):
                off = r * (2 * KBS) + sw128_offset(i, c)          # what the softmax warps write
                # the same element addressed through key block r as row 64*r + i of a 128-row tile
                assert off == r * KBS + sw128_offset(64 * r + i, c)
                assert not (KBS <= off < 2 * KBS)                 # never inside the shared zero rows
    # the other item's rows inside each key block are exactly the zero region
    for kb in (0, 1):
        other_rows = range(64, 128) if kb == 0 else range(0, 64)
        for row in other_rows:
            for c in range(8):
                off = kb * KBS + sw128_offset(row, c)
                assert KBS <= off < 2 * KBS
