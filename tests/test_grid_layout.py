"""The device grid's memory layout, checked without a device through b200geo_grid_plan — the same function
b200geo_grid_create / _create_uniform lay their buffers out with. Default layout: every row of every member starts on a
128-byte boundary with interior x = 0 exactly 128 bytes in. Uniform element layout (generic SoA path): one lead-in and
pitch in ELEMENTS for all members, member m at member_stride x (bytes of the members before it) — LibFlatArray's
addressing contract (lib/libflatarray/include/libflatarray/macros.hpp:327-349)."""
import ctypes

import numpy as np
import pytest

from libgeodecomp_b200 import capi


def plan(dim, ghost, widths, stride=0, modes=None):
    desc = capi.GridDesc()
    for i in range(3):
        desc.dim[i], desc.ghost[i] = dim[i], ghost[i]
        for s in range(2):
            desc.ghost_mode[i][s] = modes[i][s] if modes else capi.GHOST_EDGE
    desc.n_members = len(widths)
    for m, w in enumerate(widths):
        desc.member_bytes[m] = w
    out = (ctypes.c_int64 * (7 * len(widths)))()
    total = ctypes.c_int64()
    capi.check(capi.lib().b200geo_grid_plan(ctypes.byref(desc), stride, out, ctypes.byref(total)))
    rows = np.array(list(out), dtype=np.int64).reshape(len(widths), 7)
    return [dict(zip(("elem", "lead", "pitch", "plane", "origin", "bytes", "offset"), r)) for r in rows], total.value, desc


def min_stride(desc):
    least = ctypes.c_int64()
    capi.check(capi.lib().b200geo_grid_uniform_min_stride(ctypes.byref(desc), ctypes.byref(least)))
    return least.value


def cases():
    rng = np.random.default_rng(20260117)
    for _ in range(200):
        dim = [int(rng.integers(1, 300)), int(rng.integers(1, 40)), int(rng.integers(1, 20))]
        ghost = [int(rng.integers(0, 5)), int(rng.integers(0, 5)), int(rng.integers(0, 5))]
        widths = [int(w) for w in rng.choice([1, 2, 4, 8], size=int(rng.integers(1, 9)))]
        yield dim, ghost, widths


def test_default_layout_rows_are_128_byte_aligned():
    for dim, ghost, widths in cases():
        layout, total, _ = plan(dim, ghost, widths)
        end = 0
        for L, w in zip(layout, widths):
            assert L["elem"] == w and L["lead"] * w == 128
            assert (L["pitch"] * w) % 128 == 0 and L["pitch"] >= L["lead"] + dim[0] + ghost[0]
            assert L["plane"] == L["pitch"] * (dim[1] + 2 * ghost[1])
            assert L["origin"] == ghost[2] * L["plane"] + ghost[1] * L["pitch"] + L["lead"]
            assert L["offset"] == end and L["offset"] % 256 == 0
            assert L["bytes"] >= L["plane"] * (dim[2] + 2 * ghost[2]) * w + 128
            end += L["bytes"]
        assert total == end


def test_known_layouts_of_the_bench_grids():
    """the numbers DESIGN.md §4 quotes"""
    (L,), total, _ = plan((1024, 1024, 1024), (1, 1, 1), [8])
    assert (L["lead"], L["pitch"]) == (16, 1056) and abs(total - 8.9e9) < 0.1e9
    layout, total, _ = plan((512, 512, 512), (1, 1, 1), [4] * 24)
    assert layout[0]["pitch"] == 576 and abs(total - 14.6e9) < 0.1e9
    (L,), total, _ = plan((16384, 16384, 1), (1, 1, 0), [1])
    assert L["lead"] == 128 and L["pitch"] == 16640 and abs(total - 0.27e9) < 0.01e9


def test_uniform_layout_is_libflatarrays_addressing_contract():
    for dim, ghost, widths in cases():
        _, _, desc = plan(dim, ghost, widths)
        least = min_stride(desc)
        assert least % 256 == 0
        for stride in (least, least + 256, 1 << 20 if (1 << 20) >= least else least):
            layout, total, _ = plan(dim, ghost, widths, stride)
            lead, pitch = layout[0]["lead"], layout[0]["pitch"]
            assert lead * min(widths) == 128 and pitch % lead == 0 and pitch >= lead + dim[0] + ghost[0]
            before = 0
            for L, w in zip(layout, widths):
                # one element index for all members
                assert (L["lead"], L["pitch"], L["plane"], L["origin"]) == (lead, pitch, layout[0]["plane"], layout[0]["origin"])
                # member m starts DIM_PROD x offset<CELL, m> bytes into the buffer, DIM_PROD = member_stride
                assert L["offset"] == stride * before and L["bytes"] == stride * w
                # rows stay 128-byte aligned for every member width, the member array itself 256-byte aligned
                assert (L["pitch"] * w) % 128 == 0 and (L["lead"] * w) % 128 == 0 and L["offset"] % 256 == 0
                # the padded array (plus the slack behind its last row) fits
                assert L["plane"] * (dim[2] + 2 * ghost[2]) + lead <= stride
                before += w
            assert total == stride * sum(widths)


def test_single_width_uniform_layout_equals_the_default_layout():
    """a cell whose members all have the same width (Jacobi, LBM) is laid out row for row as in the default layout: the
    hand-written kernels' alignment assumptions hold on it too"""
    for widths in ([8], [4] * 5, [1, 1]):
        default, _, desc = plan((100, 7, 5), (2, 1, 1), widths)
        uniform, _, _ = plan((100, 7, 5), (2, 1, 1), widths, min_stride(desc))
        for a, b in zip(default, uniform):
            assert [a[k] for k in ("elem", "lead", "pitch", "plane", "origin")] == [b[k] for k in ("elem", "lead", "pitch", "plane", "origin")]


def test_bad_descriptions_are_rejected_without_a_device():
    with pytest.raises(ValueError):
        plan((10, 10, 10), (1, 1, 1), [3])                       # member width
    with pytest.raises(ValueError):
        plan((10, 10, 10), (1, 1, 1), [8], stride=256)           # stride smaller than the padded grid
    with pytest.raises(ValueError):
        plan((10, 10, 10), (1, 1, 1), [8], stride=1000)          # not a multiple of 256
    with pytest.raises(ValueError):
        plan((10, 10, 10), (17, 1, 1), [8])                      # x ghost wider than the lead-in
    wrap_one_side = [[capi.GHOST_WRAP, capi.GHOST_EDGE]] + [[capi.GHOST_EDGE] * 2] * 2
    with pytest.raises(ValueError):
        plan((10, 10, 10), (1, 1, 1), [8], modes=wrap_one_side)
    peer_on_x = [[capi.GHOST_PEER, capi.GHOST_PEER]] + [[capi.GHOST_EDGE] * 2] * 2
    with pytest.raises(capi.LogicError):
        plan((10, 10, 10), (1, 1, 1), [8], modes=peer_on_x)
