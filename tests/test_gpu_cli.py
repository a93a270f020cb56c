"""`python -m lungmask IN OUT` end to end on the GPU (reference: lungmask/__main__.py:78-144, tests/test_cli.py:11-20 -
there with the released weights, here with the synthetic model written to a .pth file): DICOM series directory and NIfTI
in, NIfTI / MetaImage out with the input's geometry; a re-oriented copy of the volume gives the same mask in its own
frame (mask.py:157-164,189-197)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def models():
    from oracle import synth
    return {3: synth.random_state_dict(3, seed=13, head_gain=0.3)}


def test_cli_dicom_and_reoriented_nifti(tmp_path, models):
    import torch
    from lungmask_b200 import LMInferer, io as lio
    from lungmask_b200.__main__ import main
    from oracle import synth
    from _dicom_writer import _write_dicom

    p = str(tmp_path / "w3.pth")
    torch.save(models[3], p)
    vol = synth.phantom(5, 160, 176, seed=21)
    ddir = tmp_path / "series"
    ddir.mkdir()
    for k in range(vol.shape[0]):                      # written in reverse so the reader has to sort by position
        _write_dicom(ddir / ("im%02d.dcm" % (9 - k)), vol[k], "1.2.840.1", (-80.0, -90.0, 1.5 * k))
    out_nii = str(tmp_path / "mask.nii.gz")
    main([str(ddir), out_nii, "--modelpath", p, "--noprogress", "--batchsize", "4"])
    got = lio.load_input_image(out_nii)
    want = LMInferer(modelpath=p, tqdm_disable=True, batch_size=4).apply(vol)
    assert got.array.dtype == np.uint8 and np.array_equal(got.array, want)
    assert np.allclose(got.spacing, (0.75, 0.5, 1.5)) and np.allclose(got.origin, (-80.0, -90.0, 0.0), atol=1e-4)
    assert want.max() > 0

    # the same volume stored with flipped x and swapped y / z axes: direction cosines say so, the mask comes back in that frame
    arr = np.ascontiguousarray(vol[:, :, ::-1].transpose(1, 0, 2))       # array axes (z', y', x') = (y, z, -x)
    D = np.zeros((3, 3))
    D[0, 0] = -1.0      # image x' runs along -X
    D[2, 1] = 1.0       # image y' runs along Z
    D[1, 2] = 1.0       # image z' runs along Y
    src = lio.Volume(arr, (0.75, 1.5, 0.5), (10.0, 0.0, 0.0), tuple(D.ravel()))
    in_mha = str(tmp_path / "rot.mha")
    lio.save_mask(in_mha, arr, src)
    out_mha = str(tmp_path / "rot_mask.mha")
    main([in_mha, out_mha, "--modelpath", p, "--noprogress"])
    got2 = lio.load_input_image(out_mha)
    assert got2.array.shape == arr.shape
    assert np.array_equal(got2.array.transpose(1, 0, 2)[:, :, ::-1], want)
    assert np.allclose(got2.GetDirection(), src.GetDirection())
