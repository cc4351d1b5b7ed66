"""Host-side logic of the sequence-parallel mode (physicedit_b200/ulysses.py) that needs no GPU: the row partition and the text / image
split of a rank's rows.  The device path is covered on 2 GPUs by tests/test_multi_gpu.py."""
import types

import pytest

from physicedit_b200 import ulysses


def _ctx(N):
    c = ulysses.UlyssesContext.__new__(ulysses.UlyssesContext)
    c.N, c.rank, c.Hn = N, 0, 24 // N
    return c


@pytest.mark.parametrize("N", [1, 2, 3, 4, 6, 8])
@pytest.mark.parametrize("S", [64 + 512, 8192 + 512, 8192 + 288, 20480 + 512, 130, 1])
def test_row_partition_covers_the_sequence_in_128_row_chunks(N, S):
    b = _ctx(N).bounds(S)
    assert len(b) == N + 1 and b[0] == 0 and b[-1] == S
    assert all(b[i] <= b[i + 1] for i in range(N))
    assert all(b[i] % 128 == 0 or b[i] == S for i in range(N))   # every non-empty chunk starts on a GEMM m-tile boundary
    sizes = [b[i + 1] - b[i] for i in range(N)]
    assert max(sizes) - min(s for s in sizes if s) < 128 * N or 0 in sizes


@pytest.mark.parametrize("N,T,S_img", [(2, 512, 8192), (4, 288, 8192), (8, 512, 20480), (8, 96, 512), (3, 512, 1024)])
def test_text_and_image_segments_tile_each_ranks_rows(N, T, S_img):
    S = T + S_img
    b = _ctx(N).bounds(S)
    seen_t, seen_i = [], []
    for r in range(N):
        ws = types.SimpleNamespace(lo=b[r], hi=b[r + 1], T=T)
        t, i = ulysses._segments(ws)
        if ws.hi == ws.lo:
            continue
        if t:
            assert ws.lo <= t[0] < t[1] <= min(ws.hi, T)
            seen_t.append(t)
        if i:
            assert max(ws.lo, T) == i[0] < i[1] == ws.hi
            seen_i.append(i)
        assert (t[1] - t[0] if t else 0) + (i[1] - i[0] if i else 0) == ws.hi - ws.lo
    assert seen_t[0][0] == 0 and seen_t[-1][1] == T and seen_i[0][0] == T and seen_i[-1][1] == S
    assert all(a[1] == c[0] for a, c in zip(seen_t, seen_t[1:])) and all(a[1] == c[0] for a, c in zip(seen_i, seen_i[1:]))


def test_group_size_must_divide_the_head_count():
    assert all(24 % n == 0 for n in (1, 2, 3, 4, 6, 8))
    with pytest.raises(ValueError):
        c = _ctx(5)
        if ulysses.NUM_HEADS % c.N or c.N > 8:       # the constructor's own check (needs a process group to run for real)
            raise ValueError
