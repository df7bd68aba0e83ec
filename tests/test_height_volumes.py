"""Row N1 completion: height-volume polygons in the distance field (Shaders/DistanceField.fx,
LightingRenderer.DistanceField.cs:185-260) and the incremental slice updates of RenderDistanceFieldPartition (:415-464,
MaximumFieldUpdatesPerFrame).  CPU: closed forms of the restated polygon distance and the slice bookkeeping; GPU: parity with the
oracle at +-1 LSB of the UNORM16 field, incremental == all-at-once, the static / dynamic partition with volumes."""
import numpy as np
import pytest

import illuminant_b200 as ib
from illuminant_b200 import scenes
from illuminant_b200.distance_field import DistanceField, DynamicDistanceField, LightObstruction, LightObstructionType, SimpleHeightVolume

SQUARE = [(100.0, 100.0), (200.0, 100.0), (200.0, 200.0), (100.0, 200.0)]
L_SHAPE = [(0.0, 0.0), (90.0, 0.0), (90.0, 30.0), (30.0, 30.0), (30.0, 90.0), (0.0, 90.0)]


def test_polygon_distance_closed_forms(oracle):
    """finalEval(z, zRange, sdPolygon(xy)) with PolygonXyBias = 1.5 (DistanceField.fx:13, :57-74)."""
    hv = SimpleHeightVolume(SQUARE, ZBase=0.0, Height=50.0)
    d = lambda x, y, z, v=hv: oracle.height_volume_distance(v, x, y, z)
    assert d(150, 150, 25) == pytest.approx(-48.5 - 25.0)          # inside on all axes: xy distance (biased) + z distance
    assert d(250, 150, 25) == pytest.approx(51.5)                   # outside in xy, inside in z: never negative
    assert d(150, 150, 80) == pytest.approx(30.0)                   # inside in xy, above the volume: the z distance
    assert d(250, 150, 80) == pytest.approx(51.5 + 30.0)            # outside on both
    assert d(150, 150, -10) == pytest.approx(10.0)                  # below ZBase
    assert d(230, 240, 25) == pytest.approx(50.0 + 1.5)             # nearest feature is the corner (200, 200): 30-40-50 triangle
    assert d(199.0, 150, 25) == pytest.approx(-1.0 + 1.5)           # the bias pulls the boundary 1.5 px inwards ...
    assert d(198.0, 150, 0.0) == pytest.approx(-2.0 + 1.5 + 0.0)    # ... and on the z boundary the z term is 0
    # winding does not matter, nor does the starting vertex
    rev = SimpleHeightVolume(list(reversed(SQUARE)), ZBase=0.0, Height=50.0)
    rot = SimpleHeightVolume(SQUARE[2:] + SQUARE[:2], ZBase=0.0, Height=50.0)
    for p in ((150, 150, 25), (250, 150, 25), (90, 90, 60), (101, 199, 10)):
        assert d(*p) == d(*p, v=rev) == d(*p, v=rot)
    # a concave polygon: the notch of the L is outside
    L = SimpleHeightVolume(L_SHAPE, ZBase=0.0, Height=10.0)
    assert d(60, 60, 5, v=L) == pytest.approx(30.0 + 1.5)           # 30 px from both inner edges of the notch
    assert d(15, 60, 5, v=L) == pytest.approx(-15.0 + 1.5 - 5.0)    # inside the vertical arm
    assert d(60, 15, 5, v=L) == pytest.approx(-15.0 + 1.5 - 5.0)    # inside the horizontal arm


def test_incremental_update_bookkeeping_without_a_device():
    """RenderDistanceFieldPartition: min(MaximumFieldUpdatesPerFrame, invalid slices) slices per frame, a triplet per iteration
    (`slicesToUpdate -= 3`), slices validated in triplets (LightingRenderer.DistanceField.cs:137-147, :415-464)."""
    calls = []

    def patch(df):
        df._ensure_atlas = lambda: None
        df._render_slices = lambda handle, static, first, count, obs, vols: calls.append((first, count, len(obs), len(vols), static is not None))

    df = DistanceField(None, 64, 64, 32.0, 9)
    patch(df)
    assert df.SliceCount == 9 and df.SliceInfo.InvalidSlices == list(range(9)) and df.NeedsRasterize and not df.IsFullyGenerated
    assert df.RenderDistanceField([], [], 1) == 1 and calls == [(0, 1, 0, 0, False)]            # the default budget: one triplet
    assert df.SliceInfo.InvalidSlices == [3, 4, 5, 6, 7, 8] and df.ValidSliceCount == 3
    assert df.RenderDistanceField([], [], 4) == 2 and calls[1:] == [(1, 2, 0, 0, False)]        # 4 -> 1 -> -2: two triplets, one launch
    assert not df.NeedsRasterize and df.IsFullyGenerated and df.RenderDistanceField([], [], 4) == 0
    df.Invalidate()
    assert df.SliceInfo.InvalidSlices == list(range(9)) and not df.IsFullyGenerated

    # dynamic field: the static partition first; a dynamic slice only validates once its static slice has (DistanceField.cs:276-282)
    calls.clear()
    ddf = DynamicDistanceField(None, 64, 64, 32.0, 6)
    patch(ddf)
    ddf.static_handle, ddf.handle = 1, 2
    st = LightObstruction(LightObstructionType.Box, (10, 10, 5), (4, 4, 4), IsDynamic=False)
    dy = LightObstruction(LightObstructionType.Box, (30, 30, 5), (4, 4, 4), IsDynamic=True)
    hv_s = SimpleHeightVolume(SQUARE, 0, 10, IsDynamic=False)
    hv_d = SimpleHeightVolume(SQUARE, 0, 10, IsDynamic=True)
    n = ddf.RenderDistanceField([st, dy], [hv_s, hv_d, hv_d], 1)
    assert n == 2 and calls == [(0, 1, 1, 1, False), (0, 1, 1, 2, True)]        # static slice 0 (its items), then dynamic slice 0 over it
    assert ddf.StaticSliceInfo.InvalidSlices == [3, 4, 5] and ddf.SliceInfo.InvalidSlices == [3, 4, 5]
    ddf.RenderDistanceField([st, dy], [hv_s, hv_d], 1)
    assert ddf.IsFullyGenerated and ddf.ValidSliceCount == 6
    ddf.Invalidate(False)                                                       # the per-frame case: dynamic items moved
    calls.clear()
    ddf.RenderDistanceField([st, dy], [hv_s, hv_d], 6)
    assert calls == [(0, 1, 1, 1, True), (1, 1, 1, 1, True)] and ddf.IsFullyGenerated


# ---------------------------------------------------------------------------------------------------------------- GPU
def _scene(ctx, cls=DistanceField, res=1.0):
    rs = np.random.RandomState(21)
    W, H = 320, 224
    df = cls(ctx, W, H, 96.0, 9, res, 128)
    obs = [LightObstruction(LightObstructionType.Box, (60.0, 50.0, 20.0), (20.0, 14.0, 20.0), IsDynamic=False),
           LightObstruction(LightObstructionType.Ellipsoid, (250.0, 160.0, 30.0), (30.0, 22.0, 30.0), IsDynamic=True)]
    star = [(160 + (40 if k % 2 == 0 else 16) * np.cos(k * np.pi / 5), 110 + (40 if k % 2 == 0 else 16) * np.sin(k * np.pi / 5)) for k in range(10)]
    vols = [SimpleHeightVolume([(float(np.float32(x)), float(np.float32(y))) for x, y in star], ZBase=0.0, Height=40.0, IsDynamic=False),
            SimpleHeightVolume([(20.0, 150.0), (90.0, 140.0), (110.0, 200.0), (40.0, 215.0)], ZBase=10.0, Height=60.0, IsDynamic=True),
            SimpleHeightVolume([(float(rs.uniform(200, 300)), float(rs.uniform(10, 90))) for _ in range(3)], ZBase=0.0, Height=90.0, IsDynamic=False)]
    return df, obs, vols


def _lsb_diff(a, b):
    return np.abs(a.astype(np.int32) - b.astype(np.int32))


@pytest.mark.gpu
@pytest.mark.parametrize("res", [1.0, 0.5])
def test_height_volume_field_matches_oracle(ctx, oracle, res):
    df, obs, vols = _scene(ctx, res=res)
    df.Rasterize(obs, vols)
    gpu = df.Save()
    ref = np.zeros_like(gpu)
    oracle.update_distance_field_slices(ref, df, obs, vols, 0, df.PhysicalSliceCount)
    d = _lsb_diff(gpu, ref)
    assert d.max() <= 1, f"max difference {d.max()} LSB"
    assert (d != 0).mean() < 0.02
    # the volumes really are in the field: the star's centre is inside at z = 0 (encoded above the zero level 192/255)
    zero = 192.0 / 255.0 * 65535.0
    cx, cy = int(160 * df.SliceWidth / 320), int(110 * df.SliceHeight / 224)
    assert gpu[cy, cx, 0] > zero + 1000
    without = DistanceField(ctx, 320, 224, 96.0, 9, res, 128)
    without.Rasterize(obs, [])
    assert without.Save()[cy, cx, 0] < zero


@pytest.mark.gpu
def test_incremental_slice_updates_converge_to_the_full_field(ctx, oracle):
    df, obs, vols = _scene(ctx)
    df.Rasterize(obs, vols)
    full = df.Save()
    inc = DistanceField(ctx, 320, 224, 96.0, 9, 1.0, 128)
    frames = 0
    while inc.NeedsRasterize:
        before = inc.ValidSliceCount
        assert inc.RenderDistanceField(obs, vols, maximumFieldUpdatesPerFrame=1) == 1      # the reference default: one triplet per frame
        frames += 1
        assert inc.ValidSliceCount == before + 3
        if inc.ValidSliceCount < inc.SliceCount:
            with pytest.raises(ib.IlluminantError):                                        # DistanceField.Save: "must be fully valid"
                inc.Save()
            part = np.empty_like(full)
            ctx.check(ctx.lib.ilb_df_download(inc.handle, part.ctypes.data_as(__import__("ctypes").c_void_p), part.nbytes))
            done = inc.ValidSliceCount // 3
            for p in range(inc.PhysicalSliceCount):
                ox, oy = (p % inc.ColumnCount) * inc.SliceWidth, (p // inc.ColumnCount) * inc.SliceHeight
                cell, want = part[oy:oy + inc.SliceHeight, ox:ox + inc.SliceWidth], full[oy:oy + inc.SliceHeight, ox:ox + inc.SliceWidth]
                assert np.array_equal(cell, want) if p < done else not cell.any()           # untouched slices are still cleared
    assert frames == 3 and np.array_equal(inc.Save(), full)
    # re-rasterising one slice after an obstruction moved leaves the others alone
    moved = [LightObstruction(o.Type, (o.Center[0] + 25.0, o.Center[1], o.Center[2]), o.Size, IsDynamic=o.IsDynamic) for o in obs]
    inc.SliceInfo.InvalidSlices[:] = [3, 4, 5]
    inc.RenderDistanceField(moved, vols, 3)
    after = inc.Save()
    ref = full.copy()
    oracle.update_distance_field_slices(ref, inc, moved, vols, 1, 1)
    assert _lsb_diff(after, ref).max() <= 1
    ox, oy = (0 % inc.ColumnCount) * inc.SliceWidth, 0
    assert np.array_equal(after[oy:oy + inc.SliceHeight, ox:ox + inc.SliceWidth], full[oy:oy + inc.SliceHeight, ox:ox + inc.SliceWidth])


@pytest.mark.gpu
def test_dynamic_field_with_height_volumes(ctx, oracle):
    ddf, obs, vols = _scene(ctx, DynamicDistanceField)
    ddf.Rasterize(obs, vols)
    static_ref = np.zeros((ddf.TextureHeight, ddf.TextureWidth, 4), np.uint16)
    oracle.update_distance_field_slices(static_ref, ddf, [o for o in obs if not o.IsDynamic], [v for v in vols if not v.IsDynamic], 0, ddf.PhysicalSliceCount)
    assert _lsb_diff(ddf.SaveStatic(), static_ref).max() <= 1
    static_gpu = ddf.SaveStatic()
    ref = static_gpu.copy()
    oracle.update_distance_field_slices(ref, ddf, [o for o in obs if o.IsDynamic], [v for v in vols if v.IsDynamic], 0, ddf.PhysicalSliceCount, base=static_gpu)
    assert _lsb_diff(ddf.Save(), ref).max() <= 1
    # per-frame path: the dynamic volume moves, the static field is reused
    moved = [SimpleHeightVolume([(x + 40.0, y - 20.0) for x, y in v.Polygon], v.ZBase, v.Height, v.IsDynamic) if v.IsDynamic else v for v in vols]
    ddf.RasterizeDynamic(obs, moved)
    ref2 = static_gpu.copy()
    oracle.update_distance_field_slices(ref2, ddf, [o for o in obs if o.IsDynamic], [v for v in moved if v.IsDynamic], 0, ddf.PhysicalSliceCount, base=static_gpu)
    assert _lsb_diff(ddf.Save(), ref2).max() <= 1 and not np.array_equal(ref2, ref)
    assert np.array_equal(ddf.SaveStatic(), static_gpu)
