"""SURVEY.md 8f row 4 -- wire / on-disk formats: PointCloud2 views and the KITTI .bin reader."""
import numpy as np
import pytest

from msf_loam_b200 import io as mio
from msf_loam_b200 import synth as S
from msf_loam_b200._lib import MsflError, NO_FIELD


def test_kitti_bin_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(1000, 4)).astype(np.float32)
    f = tmp_path / "000000.bin"
    pts.tofile(f)
    got = mio.read_kitti_bin(str(f))
    assert got.dtype == np.float32 and np.array_equal(got, pts)
    (tmp_path / "trunc.bin").write_bytes(pts.tobytes()[:-5])   # trailing partial point is dropped
    assert mio.read_kitti_bin(str(tmp_path / "trunc.bin")).shape == (999, 4)
    ring = mio.rings_from_elevation(np.array([[10, 0, 10 * np.tan(np.radians(2.0))], [10, 0, 10 * np.tan(np.radians(-24.8))]]))
    assert list(ring) == [63, 0]


def test_pointcloud2_view_field_mapping_and_errors():
    xyzi, ring = S.raycast_scan(S.make_scene(), "vlp16", S.trajectory(1)[0], seed=3)
    data, fields = mio.make_pointcloud2(xyzi[:100], ring[:100])
    v = mio.pointcloud2_view(data, 100, 1, 22, 2200, fields)
    assert (v.cloud.n, v.cloud.stride, v.cloud.off_xyz, v.cloud.off_intensity, v.cloud.off_ring) == (100, 22, 0, 12, 16)
    no_ring = [f for f in fields if f[0] != "ring"]
    assert mio.pointcloud2_view(data, 100, 1, 22, 2200, no_ring).cloud.off_ring == NO_FIELD
    with pytest.raises(MsflError):   # big-endian
        mio.pointcloud2_view(data, 100, 1, 22, 2200, fields, is_bigendian=True)
    with pytest.raises(MsflError):   # padded rows
        mio.pointcloud2_view(data, 100, 1, 22, 2208, fields)
    with pytest.raises(MsflError):   # x must be FLOAT32
        mio.pointcloud2_view(data, 100, 1, 22, 2200, [("x", 0, 8, 1)] + fields[1:])
    with pytest.raises(MsflError):   # field outside the point
        mio.pointcloud2_view(data, 100, 1, 22, 2200, fields[:4] + [("ring", 21, mio.PC2_UINT16, 1)])


@pytest.mark.gpu
def test_extraction_from_pointcloud2_buffer_equals_arrays():
    from msf_loam_b200 import Engine
    xyzi, ring = S.raycast_scan(S.make_scene(), "vlp16", S.trajectory(1)[0], seed=4)
    e = Engine()
    ref = e.extract_features(xyzi, ring)
    for step, offs in ((22, (0, 4, 8, 12, 16)), (32, (0, 4, 8, 16, 20)), (27, (3, 7, 11, 17, 23))):
        data, fields = mio.make_pointcloud2(xyzi, ring, point_step=step, offsets=offs)
        v = mio.pointcloud2_view(data, len(xyzi), 1, step, step * len(xyzi), fields)
        got = e.extract_features(v)
        for k in ("full", "ring", "curvature", "idx_sharp", "idx_less_sharp", "idx_flat", "idx_less_flat"):
            assert np.array_equal(got[k], ref[k]), (step, k)
    e.close()
