"""Host-side pieces of scipnp.data that need no device."""
import numpy as np
import pytest


def test_load_mat_v5_roundtrip(tmp_path):
    import scipy.io as sio
    from scipnp.data import load_mat
    rng = np.random.default_rng(0)
    orig = rng.random((6, 5, 4)) * 255
    mask = (rng.random((6, 5, 2)) > 0.5).astype(np.uint8)
    p = str(tmp_path / "scene.mat")
    sio.savemat(p, {"orig": orig, "mask": mask, "meas": np.zeros((6, 5, 2))})
    d = load_mat(p)
    assert set(d) == {"orig", "mask"}
    assert d["orig"].dtype == np.float32 and d["mask"].dtype == np.float32      # pnp_sci_test_orig.py:99-100
    np.testing.assert_array_equal(d["orig"], np.float32(orig))
    np.testing.assert_array_equal(d["mask"], np.float32(mask))
    assert list(load_mat(p, names=("meas",))) == ["meas"]
    with pytest.raises(KeyError):
        load_mat(p, names=("nothing",))


def test_load_mat_v73_needs_h5py(tmp_path):
    from scipnp.data import load_mat
    p = tmp_path / "v73.mat"
    p.write_bytes(b"MATLAB 7.3 MAT-file, Platform: GLNXA64" + b" " * 90 + b"\x89HDF\r\n\x1a\n")
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            load_mat(str(p))
