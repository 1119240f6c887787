"""The frame schedule of the temporal-sharing stem kernel (csrc/conv_stem.cu, exported host-only through
`fav_debug_stem_ts_schedule`) replayed on the CPU: every (output frame, temporal tap) pair of a tile must be issued exactly
once, from the input frame `st*t_o + kt - pt` (i3d.py:168-171 for KT = 7 / stride 2; torchvision BasicStem for KT = 3 /
stride 1), against the weight slot that holds that tap, and the first MMA that touches an accumulator must be the one that
overwrites it.  No GPU needed: the table is built by the same function the planner calls."""
import ctypes as C

import pytest

G = 4


def _schedule(KT, st, bn=64):
    from flickering_adversarial_video_b200 import _lib as L
    lib = L.load()
    nfr, nslot, ktmax = (C.c_int * 2)(), (C.c_int * 2)(), (C.c_int * 2)()
    tab = (C.c_uint32 * 16)()
    g = lib.fav_debug_stem_ts_schedule(KT, st, bn, nfr, nslot, ktmax, tab)
    assert g == G, L.last_error()
    return list(nfr), list(nslot), list(ktmax), [list(tab[0:8]), list(tab[8:16])]


@pytest.mark.parametrize("KT,st,pt", [(7, 2, 2), (7, 2, 3), (3, 1, 1)])
@pytest.mark.parametrize("To", [32, 45, 5, 2, 1])
def test_every_tap_of_every_output_frame_is_issued_once(KT, st, pt, To):
    bn = 64
    nfr, nslot, ktmax, tab = _schedule(KT, st, bn)
    assert sum(nfr[:st]) == st * (G - 1) + KT              # input frames a tile reads
    assert sum(nslot[:st]) == KT                           # every tap lives in exactly one class
    for t0 in range(0, To, G):
        geff = min(G, To - t0)                             # output frames of this tile that exist
        touched, cover = set(), {}
        for c in range(st):                                # the kernel's loop: class -> kh -> frame -> K half
            for kh in range(2):
                for f in range(nfr[c]):
                    w = tab[c][f]
                    jlo, jtop, dcol, boff = w & 15, (w >> 4) & 15, (w >> 8) & 1023, w >> 18
                    n = min(jtop, geff - 1) - jlo + 1
                    if n <= 0:
                        continue
                    dd = c + f * st                        # frame index inside the tile
                    t_in = st * t0 - pt + dd
                    assert dcol == jlo * bn
                    assert boff % (bn * 4) == 0
                    slot0 = boff // (bn * 4)
                    fresh_top = c == 0 and kh == 0 and jtop <= geff - 1
                    for half in range(2):
                        for i in range(n):
                            j, slot = jlo + i, slot0 + i
                            assert 0 <= slot < nslot[c]
                            kt = ktmax[c] - slot * st      # slot s holds tap ktmax - s*st
                            assert kt % st == c and 0 <= kt < KT
                            assert t_in == st * (t0 + j) + kt - pt
                            overwrite = half == 0 and fresh_top and i == n - 1
                            assert overwrite == (j not in touched), (t0, c, kh, f, half, j)
                            touched.add(j)
                            cover[(j, kt, kh, half)] = cover.get((j, kt, kh, half), 0) + 1
        assert touched == set(range(geff))
        assert set(cover.values()) == {1}
        assert len(cover) == geff * KT * 2 * 2


def test_rejects_unsupported_stems():
    from flickering_adversarial_video_b200 import _lib as L
    lib = L.load()
    a, b, c = (C.c_int * 2)(), (C.c_int * 2)(), (C.c_int * 2)()
    tab = (C.c_uint32 * 16)()
    assert lib.fav_debug_stem_ts_schedule(9, 2, 64, a, b, c, tab) < 0     # more than 7 taps
    assert lib.fav_debug_stem_ts_schedule(7, 3, 64, a, b, c, tab) < 0     # temporal stride 3
    assert lib.fav_debug_stem_ts_schedule(7, 2, 64, None, b, c, tab) < 0
