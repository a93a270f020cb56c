"""lungmask_b200/io.py: DICOM series / NIfTI / MetaImage ingest and mask write-back (reference behaviour:
lungmask/utils.py:132-269 read_dicoms + load_input_image, tests/test_utils.py:66-121).  The DICOM files are written by
a small encoder in this test (explicit and implicit VR, an undefined-length sequence in front of the pixel data)."""
import gzip
import os
import struct

import numpy as np
import pytest

from lungmask_b200 import io as lio
from lungmask_b200 import orient


from _dicom_writer import _elem, _write_dicom  # noqa: E402


@pytest.mark.parametrize("explicit", [True, False])
def test_dicom_series_sorted_deduplicated_filtered(tmp_path, explicit):
    rng = np.random.default_rng(1)
    vol = rng.integers(-1024, 1500, size=(5, 12, 10)).astype(np.int16)
    order = [3, 0, 4, 1, 2]
    os.makedirs(tmp_path / "a" / "b")
    for k in order:                                           # shuffled names, nested directories
        _write_dicom(tmp_path / "a" / ("s%d.dcm" % (9 - k)), vol[k], "1.2.3.4", (0, 0, 2.5 * k), explicit=explicit)
    _write_dicom(tmp_path / "a" / "b" / "dup.dcm", vol[1], "1.2.3.4", (0, 0, 2.5), explicit=explicit)        # a duplicate slice
    _write_dicom(tmp_path / "a" / "loc.dcm", vol[0], "1.2.3.4", (0, 0, 99), image_type="ORIGINAL\\PRIMARY\\LOCALIZER", explicit=explicit)
    _write_dicom(tmp_path / "a" / "der.dcm", vol[0], "1.2.3.5", (0, 0, 0), image_type="DERIVED\\SECONDARY", explicit=explicit)
    (tmp_path / "a" / "notes.txt").write_text("not a dicom file")
    vols = lio.read_dicoms(str(tmp_path))
    assert len(vols) == 1                                     # the derived series and the localizer are filtered (utils.py:152-170)
    v = vols[0]
    assert v.GetSize() == (10, 12, 5) and v.array.dtype == np.int16
    assert np.array_equal(v.array, vol)
    assert v.spacing == (0.75, 0.5, 2.5) and v.origin == (0.0, 0.0, 0.0)
    assert v.GetDirection() == (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)
    # load_input_image keeps every image type and takes the largest series (utils.py:253-268)
    got = lio.load_input_image(str(tmp_path))
    assert got.GetSize() == (10, 12, 5) or got.GetSize() == (10, 12, 6)
    assert got.meta["SeriesInstanceUID"] == "1.2.3.4"


def test_dicom_rescale_unsigned_and_orientation(tmp_path):
    rng = np.random.default_rng(2)
    stored = rng.integers(0, 3000, size=(3, 8, 8)).astype(np.uint16)
    for k in range(3):                                        # slices stacked against the in-plane normal, rows run right-to-left
        _write_dicom(tmp_path / ("%d.dcm" % k), stored[k], "9.9", (10, 20, -3.0 * k), intercept=-1024, signed=False,
                     iop="-1\\0\\0\\0\\1\\0")
    v = lio.load_input_image(str(tmp_path))
    assert v.array.dtype == np.int16
    assert np.array_equal(v.array, stored[::-1].astype(np.int32) - 1024)     # sorted by ascending z (utils.py:211-214)
    assert v.origin == (10.0, 20.0, -6.0) and v.spacing[2] == 3.0
    D = np.asarray(v.GetDirection()).reshape(3, 3)
    assert np.array_equal(D[:, 0], [-1, 0, 0]) and np.array_equal(D[:, 1], [0, 1, 0]) and np.array_equal(D[:, 2], [0, 0, 1])
    code = orient.orientation_from_direction(v.GetDirection())
    assert code == "RPS"          # x runs towards the patient's right


def test_dicom_rejects_compressed(tmp_path):
    p = tmp_path / "c.dcm"
    meta = _elem((0x0002, 0x0010), "UI", "1.2.840.10008.1.2.4.70", True)
    p.write_bytes(b"\x00" * 128 + b"DICM" + meta + _elem((0x0008, 0x0008), "CS", "ORIGINAL\\PRIMARY", True))
    with pytest.raises(lio.DicomError):
        lio.read_dicom_file(str(p))
    with pytest.raises(SystemExit):
        lio.load_input_image(str(tmp_path))                   # "No dicoms found!" (utils.py:259-260)


@pytest.mark.parametrize("ext", [".nii", ".nii.gz", ".mha"])
def test_mask_roundtrip_keeps_geometry(tmp_path, ext):
    rng = np.random.default_rng(3)
    th = 0.3
    R = np.asarray([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]]) @ np.diag([1.0, -1.0, 1.0])
    like = lio.Volume(rng.integers(-1000, 1000, size=(4, 6, 5)).astype(np.int16), (0.7, 0.8, 2.5), (-100.0, 30.5, 12.0), tuple(R.ravel()))
    mask = rng.integers(0, 3, size=like.array.shape).astype(np.uint8)
    path = str(tmp_path / ("m" + ext))
    lio.save_mask(path, mask, like)
    back = lio.load_input_image(path)
    assert back.array.dtype == np.uint8 and np.array_equal(back.array, mask)
    assert np.allclose(back.spacing, like.spacing, atol=1e-5) and np.allclose(back.origin, like.origin, atol=1e-4)
    assert np.allclose(back.GetDirection(), like.GetDirection(), atol=1e-6)
    lio.save_mask(str(tmp_path / "v" + ext) if False else str(tmp_path / ("v" + ext)), like.array, like)   # int16 volumes too
    assert np.array_equal(lio.load_input_image(str(tmp_path / ("v" + ext))).array, like.array)


def test_nifti_header_is_standard(tmp_path):
    """Field offsets of the NIfTI-1 header written here (so that other readers accept it): sizeof_hdr, dim, datatype,
    bitpix, vox_offset, sform rows in RAS, magic."""
    like = lio.Volume(np.zeros((2, 3, 4), np.int16), (0.5, 0.6, 0.7), (1.0, 2.0, 3.0))
    p = str(tmp_path / "h.nii.gz")
    lio.save_mask(p, np.zeros((2, 3, 4), np.uint8), like)
    raw = gzip.open(p, "rb").read()
    assert struct.unpack_from("<i", raw, 0)[0] == 348 and raw[344:348] == b"n+1\x00"
    assert struct.unpack_from("<8h", raw, 40)[:4] == (3, 4, 3, 2)
    assert struct.unpack_from("<hh", raw, 70) == (2, 8)
    assert struct.unpack_from("<f", raw, 108)[0] == 352.0 and len(raw) == 352 + 24
    assert np.allclose(struct.unpack_from("<4f", raw, 280), (-0.5, 0, 0, -1.0))       # LPS -> RAS: x and y negated
    assert np.allclose(struct.unpack_from("<4f", raw, 296), (0, -0.6, 0, -2.0))
    assert np.allclose(struct.unpack_from("<4f", raw, 312), (0, 0, 0.7, 3.0))


def test_nifti_qform_only(tmp_path):
    hdr = bytearray(352)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, 3, 3, 2, 2, 1, 1, 1, 1)
    struct.pack_into("<hh", hdr, 70, 4, 16)
    struct.pack_into("<8f", hdr, 76, -1.0, 2.0, 3.0, 4.0, 0, 0, 0, 0)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<2h", hdr, 252, 1, 0)
    struct.pack_into("<3f", hdr, 256, 0.0, 0.0, 0.0)                                   # identity rotation, qfac = -1
    struct.pack_into("<3f", hdr, 268, 5.0, 6.0, 7.0)
    hdr[344:348] = b"n+1\x00"
    arr = np.arange(12, dtype="<i2").reshape(2, 2, 3)
    p = tmp_path / "q.nii"
    p.write_bytes(bytes(hdr) + arr.tobytes())
    v = lio.load_input_image(str(p))
    assert np.array_equal(v.array, arr) and v.spacing == (2.0, 3.0, 4.0)
    assert v.origin == (-5.0, -6.0, 7.0)
    assert np.allclose(np.asarray(v.GetDirection()).reshape(3, 3), np.diag([-1.0, -1.0, -1.0]))


def test_reference_names_reexported():
    import lungmask.utils as u
    assert u.load_input_image is lio.load_input_image and u.read_dicoms is lio.read_dicoms
