"""Host-side check of the tcgen05 GEMM's work distribution (csrc/pw_gemm.cu:PieceMap, reached through the C ABI on the
CPU): whatever the shape and CTA count, the pieces must tile the output exactly once, and the split of the remainder
wave must only happen when it pays (remainder <= half a wave)."""
import ctypes as C

import numpy as np
import pytest

from epos_b200 import _lib


def pieces(m_tiles, N, block_n, tile_m, ctas):
    lib = _lib.lib()
    n = lib.epos_gemm_pieces(m_tiles, N, block_n, tile_m, ctas, None, 0)
    assert n > 0
    out = np.zeros((n, 3), np.int32)
    assert lib.epos_gemm_pieces(m_tiles, N, block_n, tile_m, ctas, out.ctypes.data, n) == n
    return out


@pytest.mark.parametrize('m_tiles,N,block_n,tile_m,ctas', [
    (300, 728, 256, 128, 148),      # middle flow at B = 8: 900 tiles = 6.08 waves -> remainder of 12 tiles is split
    (150, 728, 256, 256, 74),       # the same layer on CTA pairs
    (300, 1536, 256, 128, 148), (1200, 4032, 256, 128, 148), (1200, 22, 32, 128, 148), (1200, 48, 64, 128, 148),
    (38, 728, 256, 128, 148), (1, 64, 64, 128, 1), (5, 1000, 256, 128, 7), (300, 128, 128, 128, 148), (19, 2048, 256, 256, 74),
    # CTA pairs on every layer wider than 128 columns (round 2): decoder / heads at B = 8, ResNet, odd tile counts
    (600, 256, 256, 256, 74), (600, 1344, 256, 256, 74), (600, 4032, 256, 256, 74), (150, 1024, 256, 256, 74),
    (1200, 7680, 256, 256, 74), (2, 728, 256, 256, 74), (3, 136, 256, 256, 1)])
def test_pieces_tile_the_output_exactly_once(m_tiles, N, block_n, tile_m, ctas):
    p = pieces(m_tiles, N, block_n, tile_m, ctas)
    cover = np.zeros((m_tiles, N), np.int32)
    for m0, n0, nc in p:
        assert m0 % tile_m == 0 and 0 <= m0 // tile_m < m_tiles
        assert n0 % min(64, block_n) == 0 and 0 <= n0 < N and 0 < nc <= block_n
        cover[m0 // tile_m, n0:min(N, n0 + nc)] += 1
    assert cover.min() == 1 and cover.max() == 1
    n_tiles = -(-N // block_n)
    full = int((p[:, 2] == block_n).sum()) if block_n > 64 else len(p)
    rem = (m_tiles * n_tiles) % ctas
    if block_n > 64 and 0 < rem * 2 <= ctas:
        assert full == m_tiles * n_tiles - rem          # only the remainder wave is cut into 64-column blocks
        assert set(p[full:, 2].tolist()) == {64}
    else:
        assert len(p) == m_tiles * n_tiles               # plain tile grid


def test_pieces_rejects_bad_arguments():
    lib = _lib.lib()
    assert lib.epos_gemm_pieces(0, 10, 256, 128, 148, None, 0) < 0
    assert lib.epos_gemm_pieces(10, 10, 100, 128, 148, None, 0) < 0
    assert b'invalid argument' in lib.epos_last_error()
