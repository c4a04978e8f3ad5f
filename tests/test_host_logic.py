"""Host-side logic that needs no GPU: the C-ABI library loads and exports every declared symbol, channel layouts,
parameter packing, Keras weight ordering, tiling / stitching, the no-CPU-fallback guarantee."""
import ctypes
import math
import os
import random
import re

import numpy as np
import pytest
import torch

import sem_b200
from sem_b200 import _lib as L
from sem_b200.engine import Engine, Layout, ParamSpec
from sem_b200.nets import KerasGraph, UNetBuilder
from sem_b200 import HelperFunctions as HF
from oracle import unet as OU

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_of_the_header():
    hdr = open(os.path.join(ROOT, "include", "semb200.h")).read()
    declared = set(re.findall(r"\b(semb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = ctypes.CDLL(L.lib_path())
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/semb200.h but not exported"
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    assert L.load().semb_version() == 100


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(L.SembError):
        sem_b200.UNetModel((32, 32, 1))
    e = Engine(1, "f32", dry=True)
    UNetBuilder(e, 32, 32, 16)
    e.finalize()
    with pytest.raises(L.SembError):
        e.forward(False)
    src = "".join(open(os.path.join(ROOT, "automatic-sem-image-segmentation_b200", f)).read()
                  for f in os.listdir(os.path.join(ROOT, "automatic-sem-image-segmentation_b200")) if f.endswith(".py"))
    assert "import oracle" not in src and "from oracle" not in src      # the product never touches the checker


def test_layout_and_param_packing_roundtrip():
    lay = Layout.concat(Layout.simple(8), Layout.simple(17), Layout.simple(26))
    assert (lay.logical, lay.phys) == (51, 64)
    assert lay.index_map().tolist() == list(range(8)) + list(range(8, 25)) + list(range(32, 58))
    spec = ParamSpec("w", "conv_kernel", (3, 3, 51, 5), (3, 3, 64, 8), {2: lay.index_map(), 3: np.arange(5)}, True)
    w = np.random.default_rng(0).standard_normal((3, 3, 51, 5)).astype(np.float32)
    p = spec.to_phys(w)
    assert p.shape == (3, 3, 64, 8) and p[:, :, 25:32].any() == False and p[..., 5:].any() == False
    assert np.array_equal(spec.to_logical(p), w)


def test_unet_builder_matches_oracle_spec_and_packs_conv_transpose():
    e = Engine(1, "bf16", dry=True)
    b = UNetBuilder(e, 48, 32, 16)
    e.finalize()
    spec = OU.UNetSpec(16)
    assert b.creation_names == spec.names()
    p0 = spec.init_params(0)
    for n in spec.names():
        e.set_param(n, p0[n].numpy())
        assert np.array_equal(e.get_param(n), p0[n].numpy()), n
    # Conv2DTranspose kernel (2,2,Cout,Cin) is stored as a 1x1 kernel with 4*Cout outputs
    k = p0["conv2d_transpose_1/kernel"].numpy()
    phys = e.params.get("conv2d_transpose_1/kernel").numpy().reshape(e.specs["conv2d_transpose_1/kernel"].phys_shape)
    assert phys.shape == (1, 1, 432, 512)
    assert phys[0, 0, 3, 2 * 128 + 7] == k[1, 0, 7, 3]
    names = b.keras_weight_names()
    assert sorted(names) == sorted(spec.names()) and len(set(names)) == 348


def test_keras_layer_order_rule():
    """depth-sorted (deepest first), ties broken by pre-order DFS from the output over the call-order inputs [K3.5]."""
    g = KerasGraph()
    i = g.layer("input", [])
    a = g.layer("a", [i], ["a/w"])          # long branch: a -> b
    b = g.layer("b", [a], ["b/w"])
    c = g.layer("c", [i], ["c/w"])          # short branch
    o = g.layer("add", [c, b])              # call order: short branch first
    assert [g.names[k] for k in g.layer_order(o)] == ["input", "a", "c", "b", "add"]
    assert g.weight_names(o) == ["a/w", "c/w", "b/w"]


@pytest.mark.parametrize("h,w,th,tw", [(712, 1024, 256, 256), (712, 1024, 384, 384), (768, 1024, 256, 256), (100, 90, 128, 128)])
def test_tile_and_stitch_roundtrip(h, w, th, tw):
    rng = np.random.default_rng(0)
    img = rng.random((h, w, 1)).astype(np.float32)
    tiles = HF.tile_image(img, tw, th, min_overlap=2)
    if (h, w, th) == (712, 1024, 256):
        assert tiles.shape[0] == 15          # 5 x 3
    if (h, w, th) == (712, 1024, 384):
        assert tiles.shape[0] == 6
    if (h, w, th) == (768, 1024, 256):
        assert tiles.shape[0] == 20
    for mode in (0, 1, 2):
        back = HF.stitch_image(tiles, w, h, min_overlap=2, manage_overlap_mode=mode)
        assert np.allclose(back, img, atol=1e-6), mode


def test_otsu_threshold_separates_bimodal():
    rng = np.random.default_rng(0)
    img = np.concatenate([rng.normal(60, 8, 5000), rng.normal(190, 10, 3000)]).clip(0, 255).astype(np.uint8)
    t = HF.threshold_otsu(img)
    assert 80 < t < 170
    assert HF.segment(img.reshape(100, 80), threshold=-1).dtype == np.uint8


def test_gan_builders_route_strided_and_7x7_convs_to_the_tensor_core_paths():
    """Host logic only (dry engines): in bf16 mode every stride-2 conv of the CycleGAN nets is a space-to-depth op with a
    virtual (3,3,4*Cin,Cout) kernel and the 7x7 stem / head are tap-folded; in fp32 parity mode none of them is."""
    from sem_b200.engine import ConvOp
    from sem_b200.gan_nets import DiscriminatorBuilder, GeneratorBuilder
    for dtype, expect in (("bf16", True), ("f32", False)):
        e = Engine(1, dtype, dry=True)
        GeneratorBuilder(e, 32, 32, filters=8, n_res=1)
        e.finalize()
        convs = [op for op in e.ops if isinstance(op, ConvOp)]
        strided = [op for op in convs if op.geom.stride == 2]
        seven = [op for op in convs if op.geom.R == 7]
        assert len(strided) == 6 and len(seven) == 2
        assert all((op.s2d is not None) == expect for op in strided)
        assert all((op.tapfold is not None) == expect for op in seven)
        if expect:
            assert [op.tapfold["kind"] for op in seven] == ["stem", "head"]      # decided by LOGICAL channel counts (filters = 8 here)
            down = strided[0]
            assert down.s2d["k"] == 3 and (down.s2d["pt"], down.s2d["pl"]) == (0, 0) and down.geom_s.Cin == 4 * down.geom.Cin
            up = strided[-1]
            assert up.transposed and (up.s2d["pt"], up.s2d["pl"]) == (1, 1) and up.geom_t.pad_t == 1
            keys = [pk["w"] for pk in e.tc_packs if pk.get("vw")]
            # forward + flipped image per virtual kernel; the stem reads the (gradient-free) input image: no flipped image
            assert len(keys) == 2 * (len(strided) + len(seven)) - 1
        d = Engine(1, dtype, dry=True)
        DiscriminatorBuilder(d, 64, 64, filters=16)
        d.finalize()
        dconvs = [op for op in d.ops if isinstance(op, ConvOp)]
        assert [op.geom.stride for op in dconvs] == [2, 2, 2, 1]
        assert all((op.s2d is not None) == expect for op in dconvs[:3]) and dconvs[3].s2d is None
        if expect:
            assert dconvs[1].geom.H == 31 and dconvs[1].s2d_hw == (16, 16)        # odd size: zero-extended space-to-depth image


def test_image_pool_swap_branch_matches_oracle_pool():
    """ImagePool.query past pool_size entries (CycleGAN.py:927-964): same random.Random stream -> same swaps, same outputs
    as the oracle pool, including the batch_size truncation quirk (pool built with batch 2, fed batches of 5)."""
    import random
    from oracle import cyclegan as OC
    from sem_b200.cyclegan_model import ImagePool
    for pool_batch, feed in ((2, 5), (4, 4)):
        ours, ref = ImagePool(pool_batch, 6, random.Random(7)), OC.ImagePool(pool_batch, 6, random.Random(7))
        g = torch.Generator().manual_seed(0)
        swaps = 0
        for step in range(40):
            imgs = torch.rand(feed, 4, 4, 1, generator=g)
            a, b = ours.query(imgs.clone()), ref.query(imgs.clone())
            assert a.shape == b.shape == (min(pool_batch, feed), 4, 4, 1)
            assert torch.equal(a, b), step
            swaps += int(not torch.equal(a, imgs[:a.shape[0]]))
        assert ours.num_imgs == ref.num_imgs == 6 and swaps > 5
        for x, y in zip(ours.images, ref.images):
            assert torch.equal(x, y)
    assert torch.equal(ImagePool(2, 0).query(imgs), imgs)


def test_cyclegan_builders_options_create_the_reference_variables():
    from sem_b200.gan_nets import DiscriminatorBuilder, GeneratorBuilder
    from oracle import cyclegan as OC
    for opts in ({"use_skip_connection": True}, {"use_resize_convolution": True}, {}):
        e = Engine(1, "bf16", dry=True)
        b = GeneratorBuilder(e, 32, 32, 8, n_res=1, **opts)
        e.finalize()
        spec = OC.generator_spec(8, n_res=1, **opts)
        assert b.creation_names == [n for n, _, _ in spec]
        for n, shape, _ in spec:
            assert e.specs[n].logical_shape == tuple(shape), n
    e = Engine(1, "bf16", dry=True)
    b = DiscriminatorBuilder(e, 64, 64, 16, gaussian_noise=0.15)
    assert len(b.noise_ops) == 4 and b.creation_names == [n for n, _, _ in OC.discriminator_spec(16)]
    with pytest.raises(ValueError):
        GeneratorBuilder(Engine(1, "bf16", dry=True), 36, 36, 8, n_res=1, use_skip_connection=True)


def test_keras_archive_roundtrip_and_pb_reader(tmp_path):
    """keras_io: `.keras` zip (metadata.json, config.json, model.weights.npz with Keras layer paths) round trip, and the
    product's own frozen-graph reader against the oracle's (when the reference tree is present)."""
    import zipfile
    from sem_b200 import keras_io
    got = [keras_io.keras_layer_name(n) for n in ("conv2d_1", "conv2d_2", "batch_normalization_85", "stem")]
    assert got == ["conv2d", "conv2d_1", "batch_normalization_84", "stem"]
    order = ["conv2d_1/kernel", "batch_normalization_1/beta", "batch_normalization_1/moving_mean", "conv2d_transpose_1/kernel",
             "conv2d_transpose_1/bias"]
    rng = np.random.default_rng(0)
    named = {n: rng.standard_normal((3, 2)).astype(np.float32) for n in order}
    p = str(tmp_path / "model.keras")
    keras_io.save_keras(p, {"class": "MultiResUNet", "filters": 16}, named, order)
    with zipfile.ZipFile(p) as z:
        assert set(z.namelist()) == {"metadata.json", "config.json", "model.weights.npz"}
    with np.load(__import__("io").BytesIO(zipfile.ZipFile(p).read("model.weights.npz")), allow_pickle=True) as w:
        assert set(w.files) == {"layers/conv2d/vars", "layers/batch_normalization/vars", "layers/conv2d_transpose/vars"}
        assert list(w["layers/batch_normalization/vars"].item().keys()) == ["0", "1"]
    cfg, back = keras_io.load_keras(p)
    assert cfg["filters"] == 16 and all(np.array_equal(back[n], named[n]) for n in order)
    pb = "/root/reference/ImageJ Plugin/SEM_Particle_Segmentation_Models/TiO2_UNet_Masks_GAN.pb"
    if os.path.exists(pb):
        from oracle import pb_reader
        spec = OU.UNetSpec(16)
        ours = keras_io.read_pb_weights(pb, spec.names())
        ref = pb_reader.unet_params_from_pb(pb)
        assert all(np.array_equal(ours[n], ref[n]) for n in spec.names())


def test_classical_postprocessing_and_scores():
    """Measurements.Measure.segment (Otsu -> EDT -> smoothing -> local maxima -> watershed with lines), the 8 -> 4
    connectivity fix, Li threshold, and the Calculate_Scores.py metrics (whole-image / instance IoU, ROC, threshold sweep with
    and without the reference's index quirk)."""
    from scipy import ndimage
    from sem_b200 import Measurements as M, Scores
    yy, xx = np.mgrid[0:80, 0:120]
    two = (((yy - 40) ** 2 + (xx - 40) ** 2 < 22 ** 2) | ((yy - 40) ** 2 + (xx - 78) ** 2 < 22 ** 2))
    img = two.astype(np.uint8) * 200
    assert M.threshold_otsu(img) in (0.0, 99.0, 100.0) or 0 <= M.threshold_otsu(img) < 200
    seg = HF.segment(img, threshold=-1, watershed_lines=True, min_distance=9, use_four_connectivity=True)
    assert seg.dtype == np.uint8 and set(np.unique(seg)) <= {0, 255}
    assert ndimage.label(img > 0)[1] == 1 and ndimage.label(seg > 0)[1] == 2           # touching discs are split ...
    line = (img > 0) & (seg == 0)
    assert 15 <= int(line.sum()) <= 40 and not np.any((seg > 0) & (img == 0))          # ... by a thin line, nothing leaves the mask
    nosplit = HF.segment(img, threshold=-1, watershed_lines=False)
    assert np.array_equal(nosplit > 0, img > 0)
    # 8 -> 4 connectivity: a diagonal contact is broken
    d = np.zeros((6, 6), dtype=np.uint8)
    d[1:3, 1:3] = 255
    d[3:5, 3:5] = 255
    f = HF.eight_to_four_connected(d.copy())
    assert ndimage.label(f > 0)[1] == 2 and ndimage.label(f > 0, structure=np.ones((3, 3)))[1] == 2
    # Li threshold separates a bimodal image between the modes
    bi = np.concatenate([np.full(500, 20.0), np.full(300, 180.0)]) + np.random.default_rng(0).normal(0, 3, 800)
    assert 40 < M.threshold_li(bi) < 160
    # scores
    gt = two.astype(np.uint8)
    assert Scores.calculateWholeImageIoU(gt, gt) == 1.0
    half = gt.copy()
    half[:, 60:] = 0
    iou = Scores.calculateWholeImageIoU(half, gt)
    assert abs(iou - half.sum() / gt.sum()) < 1e-12
    assert Scores.calculateInstanceIoU(gt, gt) > 0.99
    tpr, tnr, fpr, fnr = Scores.ROC(half, gt)
    assert abs(tpr - iou) < 1e-12 and tnr == 1.0 and fpr == 0.0 and abs(fnr - (1 - iou)) < 1e-12
    pred = gt.astype(np.float32) * 0.65
    r = Scores.sweep_iou([pred], [gt])
    assert r["best_whole_image"][0] == 1.0 and r["whole_image"][7] == 0.0 and r["whole_image"][6] == 1.0
    q = Scores.sweep_iou([pred], [gt], reference_indexing=True)
    assert q["whole_image"][5] == 1.0 and q["whole_image"][6] == 0.0          # threshold t/10 stored at index t-1 (Calculate_Scores.py:252)


def test_workflow_steps_0_and_5_and_step_functions(tmp_path):
    from PIL import Image
    from sem_b200 import StartProcess as SP
    rng = np.random.default_rng(0)
    d = str(tmp_path)
    inp = os.path.join(d, "Input_Images")
    os.makedirs(inp)
    for i in range(2):
        img = (rng.random((200, 260)) * 80).astype(np.uint8)
        img[40:120, 60:200] += 150
        Image.fromarray(img).save(os.path.join(inp, f"im{i}.tif"))
    HF.initialize_directories(d, os.path.join(d, "oc"), os.path.join(d, "ou"))
    for sub in ("1_WGAN/Models", "2_CycleGAN/data/trainB", "2_CycleGAN/generate_images/Synthetic_Masks_Filtered", "3_UNet/Models"):
        assert os.path.isdir(os.path.join(d, sub))
    HF.prepare_images_cycle_gan(d, inp, 96, 96, num_simulated_masks=12)
    assert len(os.listdir(os.path.join(d, "2_CycleGAN", "data", "trainA"))) == 12
    assert len(os.listdir(os.path.join(d, "2_CycleGAN", "data", "testA"))) == 5
    msk = np.zeros((200, 260), np.uint8)
    msk[50:100, 80:150] = 255          # over the bright particle: kept
    msk[150:180, 10:40] = 255          # over background: filtered out
    mdir = os.path.join(d, "msk")
    os.makedirs(mdir)
    Image.fromarray(msk).save(os.path.join(mdir, "im0.tif"))
    HF.filter_gan_masks(inp, mdir, os.path.join(d, "out"), do_watershed_and_four_connectivity=False)
    o = np.array(Image.open(os.path.join(d, "out", "im0.tif")))
    assert o[75, 100] == 255 and o[165, 25] == 0
    for name in ("start_step_0", "start_step_1", "start_step_2", "start_step_3", "start_step_4", "start_step_5", "start_step_6a", "start_step_6b"):
        assert callable(getattr(SP, name))


def test_res_path_units_merge_into_pair_convs_only_in_bf16_mode(monkeypatch):
    """Host logic only (dry engines): in bf16 tensor-core mode the res_path units whose merged weight image stays resident
    (rp1 x 4 at full resolution, rp2 units 2-3) run their 3x3 conv and 1x1 shortcut as one PairConvOp; parameter names,
    creation order and the Keras layer order are identical to the unmerged build; fp32 parity mode never merges."""
    from sem_b200.engine import ConvOp, PairConvOp

    def build(dtype):
        e = Engine(1, dtype, dry=True)
        b = UNetBuilder(e, 64, 64, 16)
        e.finalize()
        return e, b

    e16, b16 = build("bf16")
    e32, b32 = build("f32")
    pairs = [op for op in e16.ops if isinstance(op, PairConvOp)]
    assert len(pairs) == 6 and not any(isinstance(op, PairConvOp) for op in e32.ops)
    assert [(op.geom.Cin, op.geom.Cout) for op in pairs] == [(32, 32), (16, 32), (16, 32), (16, 32), (32, 64), (32, 64)]
    assert b16.creation_names == b32.creation_names and b16.keras_weight_names() == b32.keras_weight_names()
    n_conv = lambda e: sum(isinstance(op, ConvOp) for op in e.ops)
    assert n_conv(e32) - n_conv(e16) == 6
    monkeypatch.setenv("SEMB_NO_PAIR_CONV", "1")
    e_off, _ = build("bf16")
    assert not any(isinstance(op, PairConvOp) for op in e_off.ops)


def test_res_path_lanes_are_planned_around_their_producers_and_consumers():
    """Host logic only: every res_path is a side lane; forward joins at the decoder block that reads the skip concat; the
    backward order moves the lane right behind that block's backward (not behind the whole deep part of the net) and the
    op that accumulates into the lane's input gradient (the max-pool of the level) joins it."""
    from sem_b200.engine import ConvOp, PoolOp
    e = Engine(1, "bf16", dry=True)
    UNetBuilder(e, 64, 64, 16)
    e.finalize()
    lanes = sorted({op.lane for op in e.ops if op.lane})
    assert lanes == [1, 2, 3, 4]
    order = e.bwd_order
    assert sorted(map(id, order)) == sorted(map(id, e.ops))
    pos = {id(op): i for i, op in enumerate(order)}
    fpos = {id(op): i for i, op in enumerate(e.ops)}
    for l in lanes:
        group = [op for op in e.ops if op.lane == l]
        joiner = [op for op in e.ops if l in op.join_fwd]
        assert len(joiner) == 1 and isinstance(joiner[0], ConvOp) and joiner[0].lane == 0
        assert fpos[id(joiner[0])] > max(fpos[id(op)] for op in group)          # forward: consumer after the lane
        gp = sorted(pos[id(op)] for op in group)
        assert gp == list(range(gp[0], gp[0] + len(group)))                      # contiguous in the backward order ...
        assert gp[0] == pos[id(joiner[0])] + 1                                   # ... right behind the consumer's backward
        assert [pos[id(op)] for op in reversed(group)] == gp                     # ... in reversed forward order
        bj = [op for op in e.ops if l in op.join_bwd]
        assert len(bj) == 1 and isinstance(bj[0], PoolOp) and pos[id(bj[0])] > gp[-1]
    # ops that share a gradient buffer keep their planned relative order: the lane's first conv writes d(m_k) before the pool adds to it
    main = [op for op in order if op.lane == 0]
    assert main == [op for op in reversed(e.ops) if op.lane == 0]


def test_simplex_noise_and_mask_grid_helpers():
    """Host pieces of WGAN.simulate_masks (WassersteinGAN.py:375-545): the restated 2-D OpenSimplex noise is deterministic per
    seed, smooth and inside [-1, 1]; the jittered hexagonal grid offsets odd rows by half a cell and keeps the reference's unfilled
    trailing slots at the origin."""
    from sem_b200 import simplex_noise as sn
    from sem_b200.WassersteinGAN import WGAN
    sn.seed(7)
    x, y = np.arange(0, 4, 4 / 300), np.arange(0, 4, 4 / 200)
    a = sn.noise2array(x, y)
    assert a.shape == (200, 300) and -1.0 <= a.min() < -0.3 and 0.3 < a.max() <= 1.0
    assert np.abs(np.diff(a, axis=0)).max() < 0.05 and np.abs(np.diff(a, axis=1)).max() < 0.05        # continuous
    sn.seed(7)
    assert np.array_equal(a, sn.noise2array(x, y))
    sn.seed(8)
    assert np.abs(a - sn.noise2array(x, y)).max() > 0.1
    assert abs(float(sn.noise2array(np.array([0.0]), np.array([0.0]))[0, 0])) < 1e-12                  # lattice points carry zero noise
    np.random.seed(0)
    px, py = WGAN._grid_positions('HEXAGONAL', 200, 100, 8.0, 8.0, 1, 1)
    assert px.size == py.size == math.ceil(100 / 8) * math.ceil(200 / 8) + 1 and px.min() >= 0 and px.max() <= 200 and py.max() <= 100
    rows = {}
    for xx, yy in zip(px, py):
        rows.setdefault(int(round(yy / 8)), []).append(int(xx))
    assert np.median(np.asarray(sorted(rows[1])) % 8) in (3, 4, 5)               # odd rows: shifted by half of the 8-pixel cell (+- jitter 1)
    cx, cy = WGAN._grid_positions('CUBIC', 64, 64, 8.0, 8.0, 1, 1)
    assert cx.size == 64


def test_simulate_masks_places_separated_particles(tmp_path):
    """WGAN.simulate_masks (WassersteinGAN.py:375-545) with the generator replaced by discs: hexagonal grid + noise clustering +
    overlap limit (the default path), and the free-position path with noise-driven rotation and normally distributed sizes."""
    from PIL import Image
    from scipy import ndimage
    from sem_b200.WassersteinGAN import WGAN
    root = str(tmp_path)
    os.makedirs(os.path.join(root, "Input_Masks"))
    yy, xx = np.mgrid[0:32, 0:32]
    disc = ((yy - 16) ** 2 + (xx - 16) ** 2 < 100).astype(np.uint8) * 255
    Image.fromarray(disc).save(os.path.join(root, "Input_Masks", "p.tif"))
    wg = WGAN(root_dir=root)
    assert wg.train_images.shape == (4, 32, 32, 1)
    wg.model = object()                                            # never called: the particles come from the stub below
    wg._generate_particles = lambda count: np.repeat(disc[None], count, 0)
    np.random.seed(1)
    random.seed(1)
    wg.simulate_masks(no_of_images=2, img_width=128, img_height=96)
    files = sorted(os.listdir(wg.generate_dir))
    assert files == ["00000.tif", "00001.tif"] and sorted(os.listdir(os.path.join(root, "2_CycleGAN", "data", "testB"))) == files
    for f in files:
        m = np.array(Image.open(os.path.join(wg.generate_dir, f)))
        assert m.shape == (96, 128) and set(np.unique(m)) <= {0, 255} and 0.02 < (m > 0).mean() < 0.9
        lab, n = ndimage.label(m > 0)
        sizes = ndimage.sum(m > 0, lab, range(1, n + 1))
        # an eroded disc of radius 10 has ~200 pixels; overlap <= 1 % keeps the particles apart (clipped ones at the border are smaller)
        assert n >= 3 and sizes.max() < 2.2 * 210
    wg.generate_dir = os.path.join(root, "free")
    wg.simulate_masks(no_of_images=1, min_no_of_particles=20, max_no_of_particles=30, use_normal_distribution=True, max_overlap=None,
                      use_random_rotation='PERLIN', img_width=96, img_height=96)
    m = np.array(Image.open(os.path.join(root, "free", "00000.tif")))
    assert m.shape == (96, 96) and (m > 0).any()


def test_wgan_dataset_preparation_and_archive_layout(tmp_path):
    """WGAN.__init__ (WassersteinGAN.py:288-357) without a device: masks thresholded to {-1, 1}, four flips each, zero-padded
    (value 0, the middle of the [-1, 1] range -- the reference's choice) to a common size divisible by 16, with the reference's
    width rule (the running HEIGHT maximum enters the width maximum, :333)."""
    from PIL import Image
    from sem_b200.WassersteinGAN import WGAN, GANMonitor
    root = str(tmp_path)
    os.makedirs(os.path.join(root, "Input_Masks"))
    a = np.zeros((20, 30), np.uint8); a[5:15, 8:22] = 200
    b = np.zeros((40, 18), np.uint8); b[10:30, 4:14] = 255
    Image.fromarray(a).save(os.path.join(root, "Input_Masks", "a.tif"))
    Image.fromarray(b).save(os.path.join(root, "Input_Masks", "b.tif"))
    wg = WGAN(root_dir=root)
    # heights 20, 40 -> 48; widths: max(20, 30) = 30, then max(40, 18) = 40 (the height leaks in) -> 48
    assert wg.train_images.shape == (8, 48, 48, 1) and wg.train_images.dtype == np.float32
    assert set(np.unique(wg.train_images)) == {-1.0, 0.0, 1.0}
    first = wg.train_images[0, :, :, 0]
    top, left = (48 - 20) // 2, (48 - 30) // 2
    assert np.array_equal(first[top:top + 20, left:left + 30] > 0, a > 127) and float(np.abs(first[:top]).max()) == 0.0
    assert np.array_equal(wg.train_images[1, :, :, 0][top:top + 20, left:left + 30], np.fliplr(first[top:top + 20, left:left + 30]))
    assert (wg.batch_size, wg.epochs, wg.n_z) == (64, 1000, 128)
    assert wg.generate_dir.endswith(os.path.join("2_CycleGAN", "data", "trainB")) and wg.model_dir.endswith(os.path.join("1_WGAN", "Models"))
    assert WGAN.discriminator_loss(np.array([1.0, 3.0]), np.array([0.0, 1.0])) == -1.5 and WGAN.generator_loss(np.array([2.0, 4.0])) == -3.0

    class _Stub:                                             # GANMonitor only needs `n` and a callable model
        n = 4
        def __call__(self, z, training=False):
            return np.tile(np.linspace(-1, 1, 48, dtype=np.float32)[None, :, None, None], (z.shape[0], 1, 48, 1))
    mon = GANMonitor(os.path.join(root, "mosaics"), num_img=9, latent_dim=16, output_epochs=20)
    mon.set_model(_Stub())
    assert mon.on_epoch_end(1) is None                       # only every 20th epoch
    sheet = np.array(Image.open(mon.on_epoch_end(20)))
    assert sheet.shape == (2 * 48, 3 * 48) and sheet[0, 0] == 0 and sheet[47, 0] == 255


def test_weight_pack_sizes_follow_the_chunk_plan():
    """semb_pack_weights_tc(dst=NULL) is a host-side size query (no device work): the packed bf16 image is
    [n-chunk][k-chunk][tap][K chunk][N chunk], and layers whose weights are STREAMED per K chunk are planned with N <= 128
    (the two-tiles-per-weight-chunk mode of conv_tma.cu, tc_common.cuh::tc_plan)."""
    lib = L.load()
    size = lambda k, ci, co, flip=0: lib.semb_pack_weights_tc(None, k, k, ci, co, flip, None, None)
    assert size(3, 16, 16) == 9 * 16 * 16 * 2                                   # resident, one chunk
    assert size(3, 8, 32) == 9 * 16 * 32 * 2                                    # K padded to one 16-channel MMA step
    assert size(3, 512, 512) == 4 * 16 * 9 * 32 * 128 * 2 == 9 * 512 * 512 * 2  # CycleGAN residual conv: 4 N chunks of 128, 16 K chunks of 32
    assert size(3, 144, 216) == 2 * 5 * 9 * 32 * 112 * 2                        # re-planned from one N chunk of 224 to two of 112
    assert size(3, 120, 64) == 1 * 4 * 9 * 32 * 64 * 2                          # K chunk sized for the 18 x 18 pair halo (32, not 64)
    assert size(1, 216, 1024, 0) == 8 * 4 * 64 * 128 * 2 and size(1, 216, 1024, 1) == 2 * 16 * 64 * 112 * 2      # forward / flipped (data gradient)
    assert lib.semb_pack_weights_tc(None, 5, 5, 16, 16, 0, None, None) < 0      # only 1x1 and 3x3 images exist (5x5 runs via space-to-depth)
