"""Validates the oracle against the reference ITSELF (imported verbatim from /root/reference with the
stand-ins of oracle/standins.py).  Only runs where the reference tree exists (the build container)."""
import numpy as np
import pytest
import torch

from oracle import ref_loader, restate, synth

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


def test_reference_known_answer_tests_through_standins(ref):
    """tests/test_utils.py:58-63,73-107,124-159 of the reference, verbatim assertions."""
    u = ref.utils
    m = np.zeros((10, 10, 10), dtype=np.uint8)
    m[2:8, 3:7, 4:6] = 1
    assert tuple(u.bbox_3D(m, margin=2)) == (0, 10, 1, 9, 2, 8)
    img = np.full((10, 10), dtype=np.int16, fill_value=-1000)
    img[2:8, 3:7] = 1
    img[9, 9] = 1
    assert np.sum(u.simple_bodymask(img)) == 24
    cropped, bb = u.crop_and_resize(img, width=20, height=20)
    assert tuple(bb) == (2, 3, 8, 7) and cropped.shape == (20, 20) and np.sum(cropped) == 400
    out = u.reshape_mask(np.full((10, 10), dtype=np.uint8, fill_value=1), (2, 2, 22, 22), origsize=(30, 30))
    assert out.shape == (30, 30) and np.sum(out) == 400
    li = np.zeros((1, 6, 6), dtype=np.uint8)
    li[0] = np.asarray([[0, 0, 0, 0, 0, 0], [0, 1, 1, 2, 2, 0], [0, 2, 0, 3, 1, 0], [0, 4, 4, 4, 0, 0], [0, 4, 0, 4, 0, 0], [0, 4, 4, 4, 0, 0]])
    gt = [[0, 0, 0, 0, 0, 0], [0, 1, 1, 2, 2, 0], [0, 1, 0, 3, 2, 0], [0, 4, 4, 4, 0, 0], [0, 4, 0, 4, 0, 0], [0, 4, 4, 4, 0, 0]]
    vol = np.tile(li, (2, 1, 1))
    assert np.all(u.postprocessing(vol, spare=[], disable_tqdm=True, skip_below=1)[0] == gt)
    assert u.postprocessing(vol, spare=[3], disable_tqdm=True, skip_below=1)[0][2, 3] == 2
    assert u.postprocessing(vol, spare=[3], disable_tqdm=True, skip_below=3)[0][2, 1] == 0
    # and the same KATs hold for the restatement
    assert np.sum(restate.simple_bodymask(img)) == 24
    c2, b2 = restate.crop_and_resize(img, width=20, height=20)
    assert tuple(b2) == (2, 3, 8, 7) and np.sum(c2) == 400
    assert np.all(restate.postprocessing(vol, spare=[], skip_below=1)[0] == gt)


def test_state_dict_schema_is_the_reference_layout(ref):
    for K in (3, 6):
        m = ref.resunet.UNet(n_classes=K, padding=True, depth=5, up_mode="upsample", batch_norm=True, residual=False)
        rsd = m.state_dict()
        sch = synth.schema(K)
        assert [k for k, _, _ in sch] == list(rsd.keys())
        assert all(tuple(rsd[k].shape) == tuple(s) for k, s, _ in sch)
        assert len(sch) == 227


@pytest.mark.parametrize("shape,seed", [((3, 256, 256), 0), ((2, 300, 414), 1), ((2, 97, 200), 2), ((2, 40, 52), 3)])
def test_preprocess_equals_reference(ref, shape, seed):
    vol = synth.phantom(*shape, seed=seed)
    a, ba = ref.utils.preprocess(vol, resolution=[256, 256])
    r, br = restate.preprocess(vol, resolution=[256, 256])
    assert np.array_equal(a, r) and np.array_equal(np.asarray(ba), np.asarray(br))


@pytest.mark.parametrize("S,K,seed", [(8, 3, 0), (5, 6, 1), (1, 3, 2)])
def test_postprocessing_equals_reference(ref, S, K, seed):
    lab = synth.label_noise_volume(S, K, seed=seed, speckle=2e-3)
    assert np.array_equal(ref.utils.postprocessing(lab, disable_tqdm=True), restate.postprocessing(lab))
    assert np.array_equal(ref.utils.postprocessing(lab, spare=[K - 1], disable_tqdm=True), restate.postprocessing(lab, spare=[K - 1]))


def test_forward_and_apply_equal_reference(ref, tmp_path):
    sd = synth.random_state_dict(3, seed=13)
    p = str(tmp_path / "w.pth")
    torch.save(sd, p)
    inf = ref.mask.LMInferer(modelname="R231", modelpath=p, force_cpu=True, batch_size=2, tqdm_disable=True)
    vol = synth.phantom(3, 200, 216, seed=5)
    taps = {}
    assert np.array_equal(inf.apply(vol), restate.inference(vol, sd, batch_size=2, taps=taps))
    x = torch.as_tensor(restate.normalise(taps["resized"])[:, None], dtype=torch.float32)
    with torch.inference_mode():
        assert torch.equal(inf.model(x), torch.as_tensor(taps["scores"]))


def test_preprocess_equals_reference_on_ragged_and_noisy_volumes(ref):
    """slices smaller than the 128 x 128 thumbnail, non-square and odd sizes, volumes without a clear body"""
    shapes = [(2, 17, 23), (1, 128, 128), (2, 129, 127), (1, 64, 300), (2, 511, 513), (1, 33, 33), (2, 200, 100), (1, 12, 12),
              (2, 256, 255), (1, 150, 400)]
    vols = [synth.phantom(*sh, seed=50 + i) for i, sh in enumerate(shapes)]
    rng = np.random.default_rng(0)
    vols += [rng.normal(-400, 400, size=(2, 90 + 7 * i, 110 + 5 * i)).astype(np.int16) for i in range(6)]
    for vol in vols:
        a, ba = ref.utils.preprocess(vol, resolution=[256, 256])
        r, br = restate.preprocess(vol, resolution=[256, 256])
        assert np.array_equal(a, r) and np.array_equal(np.asarray(ba), np.asarray(br)), vol.shape


def test_postprocessing_equals_reference_sweep(ref):
    """random sizes, class counts and speckle levels x the spare / skip_below combinations of utils.py:272"""
    rng = np.random.default_rng(0)
    for seed in range(10, 22):
        S, K = int(rng.integers(1, 7)), int(rng.choice([3, 6]))
        lab = synth.label_noise_volume(S, K, seed=seed, speckle=float(rng.choice([5e-4, 2e-3, 1e-2])), H=int(rng.choice([48, 64, 96])),
                                       W=int(rng.choice([48, 80, 128])))
        for spare, skip in (([], 3), ([K - 1], 3), ([], 1), ([1], 2)):
            a = ref.utils.postprocessing(lab.copy(), spare=list(spare), disable_tqdm=True, skip_below=skip)
            assert np.array_equal(a, restate.postprocessing(lab.copy(), spare=list(spare), skip_below=skip)), (seed, S, K, spare, skip)
