"""-m gpu: tmae_generate_scores (csrc/scores.cu) through the C ABI against the oracle and the reference-generated goldens:
byte work, so everything is asserted bit-exact (segmented image, both maps, fp32 scores incl. NaN placement)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import ref_scores

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


def _run(gray_np, side=224):
    from textmae_image_compression_b200.scores import generate_scores
    g = torch.from_numpy(np.ascontiguousarray(gray_np)).cuda()
    sc, s_map, t_map, seg = generate_scores(g, side, return_maps=True)
    torch.cuda.synchronize()
    return sc.cpu().numpy(), s_map.cpu().numpy(), t_map.cpu().numpy(), seg.cpu().numpy()


def test_kodak_scores_equal_reference_generated(cuda_dev):
    gray = np.load(GOLDEN / "kodak_gray6.npz")
    gold = torch.load(GOLDEN / "kodak_scores.pt").numpy()
    names = sorted(gray.files)
    for shape in {gray[n].shape for n in names}:                 # one batched launch per image size
        sel = [i for i, n in enumerate(names) if gray[n].shape == shape]
        sc, s_map, t_map, seg = _run(np.stack([gray[names[i]] for i in sel]))
        for b, i in enumerate(sel):
            o_sc, o_s, o_t, o_seg = ref_scores.generate_scores(gray[names[i]], return_maps=True)
            assert np.array_equal(seg[b], o_seg), names[i]
            assert np.array_equal(s_map[b], o_s) and np.array_equal(t_map[b], o_t), names[i]
            assert np.array_equal(sc[b], gold[i]), names[i]       # the reference generator's own output


def test_synthetic_cases_equal_reference_goldens(cuda_dev):
    z = np.load(GOLDEN / "scores_refexec.npz")
    for k, (kind, h, w, seed) in enumerate(z["cases"]):
        img = ref_scores.synthetic_gray(int(kind), int(h), int(w), int(seed))
        sc, s_map, t_map, _ = _run(img)
        assert np.array_equal(s_map, z[f"case{k}_s_map"]), k
        assert np.array_equal(t_map, z[f"case{k}_t_map"]), k
        assert np.array_equal(sc, z[f"case{k}_scores"], equal_nan=True), k


@pytest.mark.parametrize("shape,side", [((1080, 2048), 224), ((1080, 2048), 512), ((8, 8), 16), ((9, 301), 224), ((224, 224), 224),
                                        ((100, 90), 224), ((2047, 1025), 64)])
def test_geometry_sweep_equals_oracle(cuda_dev, shape, side):
    """DIV2K-sized, tiny, up-scaled, odd sizes (uncovered last rows / columns), other output sides."""
    for kind in (0, 1, 2, 4):
        img = ref_scores.synthetic_gray(kind, shape[0], shape[1], 31 * kind + shape[0])
        sc, s_map, t_map, seg = _run(img, side)
        o_sc, o_s, o_t, o_seg = ref_scores.generate_scores(img, side, return_maps=True)
        assert np.array_equal(seg, o_seg), kind
        assert np.array_equal(s_map, o_s) and np.array_equal(t_map, o_t), kind
        assert np.array_equal(sc, o_sc, equal_nan=True), kind


def test_batch_and_feeds_mask_select(cuda_dev):
    """A batch of images in one call == the images one by one; the scores drive the mask kernel like reference scores do."""
    from textmae_image_compression_b200.scores import generate_scores
    imgs = np.stack([ref_scores.synthetic_gray(k % 5, 384, 512, 100 + k) for k in range(9)])
    batch = generate_scores(torch.from_numpy(imgs).cuda())
    single = torch.stack([generate_scores(torch.from_numpy(imgs[k]).cuda()) for k in range(9)])
    assert torch.equal(batch.nan_to_num(-1), single.nan_to_num(-1))
    o = np.stack([ref_scores.generate_scores(imgs[k]) for k in range(9)])
    assert np.array_equal(batch.cpu().numpy(), o, equal_nan=True)


def test_rejects_bad_geometry_and_cpu_tensors(cuda_dev):
    from textmae_image_compression_b200.scores import generate_scores
    with pytest.raises(ValueError):
        generate_scores(torch.zeros(1, 4, 300, dtype=torch.uint8, device="cuda"))
    with pytest.raises(ValueError):
        generate_scores(torch.zeros(1, 64, 64, dtype=torch.uint8, device="cuda"), out_side=100)
    with pytest.raises(RuntimeError):
        generate_scores(torch.zeros(1, 64, 64, dtype=torch.uint8))
    with pytest.raises(TypeError):
        generate_scores(torch.zeros(1, 64, 64, device="cuda"))


def test_preprocess_image_scores_keeps_the_reference_file_contract(cuda_dev, tmp_path):
    """generate_scores_file.py:13-36: sorted rglob over the dataset folder, one row per image, torch.save of the stack; images
    of different sizes in one folder (batched per size on the GPU)."""
    cv2 = pytest.importorskip("cv2")
    from textmae_image_compression_b200.scores import preprocess_image_scores
    imgs = {"b.png": ref_scores.synthetic_gray(0, 300, 411, 1), "a.png": ref_scores.synthetic_gray(4, 256, 384, 2),
            "sub/c.png": ref_scores.synthetic_gray(2, 300, 411, 3)}
    (tmp_path / "data" / "sub").mkdir(parents=True)
    for name, im in imgs.items():
        assert cv2.imwrite(str(tmp_path / "data" / name), im)
    out_file = tmp_path / "scores.pt"
    got = preprocess_image_scores(tmp_path / "data", out_file)
    saved = torch.load(out_file)
    assert torch.equal(saved.nan_to_num(-1), got.nan_to_num(-1)) and saved.shape == (3, 196) and saved.dtype == torch.float32
    order = sorted(imgs)                                       # sorted(Path.rglob) == lexicographic full paths: a.png, b.png, sub/c.png
    for row, name in zip(saved, order):
        assert np.array_equal(row.numpy(), ref_scores.generate_scores(imgs[name]), equal_nan=True), name


def test_generated_scores_drive_the_model_like_reference_scores(cuda_dev, kodak):
    """grey Kodak image -> GPU scores -> mask kernel + forward: identical to the run on the reference-generated scores."""
    from textmae_image_compression_b200 import MCM, make_state_dict, vit_base
    from textmae_image_compression_b200.scores import generate_scores
    gray = np.load(GOLDEN / "kodak_gray6.npz")
    names = sorted(gray.files)
    sel = [i for i, n in enumerate(names) if gray[n].shape == (512, 768)][:4]
    sc = generate_scores(torch.from_numpy(np.stack([gray[names[i]] for i in sel])).cuda())
    imgs, gold_scores = kodak
    assert torch.equal(sc.cpu(), gold_scores[sel])
    cfg = vit_base(64)
    m = MCM(num_keep_patches=64, softmax_isa=16)
    m.load_state_dict(make_state_dict(cfg, seed=0))
    m.cuda().eval()
    a = m(imgs[sel].cuda(), sc, need_recon=False)
    b = m(imgs[sel].cuda(), gold_scores[sel].cuda(), need_recon=False)
    assert torch.equal(a["ids_restore"], b["ids_restore"]) and torch.equal(a["latents"]["y_sym"], b["latents"]["y_sym"])
