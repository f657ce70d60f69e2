"""CPU-side host logic: config keys, registries, and the N > 1 gradient-synchronisation logic with world_size-2 gloo processes."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config_keys_match_reference_names():
    from afigan.config import get_cfg
    cfg = get_cfg()
    for k in ("GUIDE_ARCHITECTURE", "GUIDE_WEIGHTS", "AFI_GEN_WEIGHTS", "AFI_DIS_WEIGHTS", "AF_EXTRACTOR_WEIGHTS", "AFI_FREEZE"):
        assert k in cfg.MODEL, k                                   # reference afigan/config/defaults.py:5-11
    assert cfg.MODEL.GUIDE_BACKBONE.NAME == "build_resnet_fpn_backbone" and cfg.MODEL.GUIDE_BACKBONE.FREEZE_AT == 2
    assert cfg.MODEL.BIFPN.FPN_REPEAT == 3 and cfg.MODEL.BIFPN.NORM == "SyncBN"
    assert cfg.MODEL.SWINT.DEPTHS == [2, 2, 6, 2]
    assert cfg.MODEL.RESNETS.RADIX == 1 and cfg.SOLVER.OPTIMIZER == "SGD" and cfg.SOLVER.AMP.ENABLED is False
    cfg.merge_from_list(["MODEL.AFI_FREEZE", True, "MODEL.AFI_GEN_WEIGHTS", "/tmp/g.pth"])
    assert cfg.MODEL.AFI_FREEZE is True
    with pytest.raises(KeyError):
        cfg.merge_from_list(["MODEL.SRF_FREEZE", True])            # the key one shipped yaml uses by mistake stays an error


def test_backbone_registry_names():
    from afigan._compat import BACKBONE_REGISTRY
    import afigan.modeling  # noqa: F401
    for name in ("build_resnet_fpn_sr_backbone", "build_resnest_fpn_sr_backbone", "build_resnet_pafpn_sr_backbone",
                 "build_resnest_pafpn_sr_backbone", "build_swint_bifpn_sr_backbone"):
        assert callable(BACKBONE_REGISTRY.get(name))
    from afigan.config import get_cfg
    from afigan.modeling import GUIDE_ARCH_REGISTRY, build_guide_model
    cfg = get_cfg()
    cfg.MODEL.GUIDE_ARCHITECTURE = "RCNN_FPN_only"                 # configs/step1_*.yaml:5
    assert callable(GUIDE_ARCH_REGISTRY.get("RCNN_FPN_only"))
    assert callable(BACKBONE_REGISTRY.get(cfg.MODEL.GUIDE_BACKBONE.NAME))      # build_resnet_fpn_backbone (detectron2's, or the stand-in)


def test_guide_model_and_paired_scale_mapper_produce_the_config1_shape_pairs():
    """an 800x1333 image and its int(0.5 x) copy, padded to 32, give HR levels 200x336 .. 13x21 and LR levels
    104x168 .. 7x11 (reference rcnn_only.py:34-44, dataset_mapper.py:70-125, transform_gen.py:540-552).  Runs the guide ResNet-50-FPN on
    the CPU at 1/4 of the image size (the shape arithmetic is the same) and checks the full-size arithmetic without running it."""
    import numpy as np
    from afigan.config import get_cfg
    from afigan.engine import PairedScaleMapper, shortest_edge_size
    from afigan.modeling import build_guide_model
    from afigan.modeling.meta_arch import pad_to_divisible
    assert shortest_edge_size(480, 640, 800, 1333) == (800, 1067) and shortest_edge_size(600, 1000, 800, 1333) == (800, 1333)
    rng = np.random.default_rng(0)
    img = rng.integers(0, 255, size=(150, 250, 3), dtype=np.uint8)
    annos = [{"bbox": [10.0, 20.0, 110.0, 70.0], "bbox_mode": "XYXY_ABS"}]
    flipped = plain = None
    mapper = PairedScaleMapper(min_size=(200,), max_size=333, rng=rng)
    for _ in range(20):
        d = mapper({"image_array": img, "annotations": annos})
        assert tuple(d["image"].shape) == (3, 200, 333) and tuple(d["image_x0.5"].shape) == (3, 100, 166)      # int(0.5 * 333) = 166
        b, b2 = d["instances"]["gt_boxes"][0], d["instances_x0.5"]["gt_boxes"][0]
        assert torch.allclose(b2 * torch.tensor([333 / 166, 2.0, 333 / 166, 2.0]), b, atol=1e-3)                 # same flip at both scales
        if float(b[0]) > 150:
            flipped = d
        else:
            plain = d
    assert flipped is not None and plain is not None
    assert torch.equal(flipped["image"], plain["image"].flip(2)) and torch.equal(flipped["image_x0.5"], plain["image_x0.5"].flip(2))
    # full-size arithmetic (not executed): 800x1333 -> pad 32 -> 800x1344; LR 400x666 -> 416x672
    for (h, w), want in (((800, 1333), (800, 1344)), ((400, 666), (416, 672))):
        assert tuple(pad_to_divisible([torch.zeros(1, h, w)], 32).shape[2:]) == want
    cfg = get_cfg()
    cfg.MODEL.DEVICE = "cpu"
    cfg.MODEL.GUIDE_ARCHITECTURE = "RCNN_FPN_only"
    torch.manual_seed(0)
    guide = build_guide_model(cfg).eval()
    assert guide.backbone.size_divisibility == 32
    keys = set(guide.state_dict())
    assert {"backbone.bottom_up.stem.conv1.weight", "backbone.bottom_up.res2.0.shortcut.norm.running_var", "backbone.bottom_up.res5.2.conv3.weight",
            "backbone.fpn_lateral5.weight", "backbone.fpn_output2.bias"} <= keys                                  # detectron2 R-50-FPN key names
    out = guide([plain], "image")[0]["features"]
    assert [tuple(out[k].shape[1:]) for k in ("p2", "p3", "p4", "p5", "p6")] == [(256, 56, 88), (256, 28, 44), (256, 14, 22), (256, 7, 11), (256, 4, 6)]
    lr_f, hr_f = guide.extract_pair([plain, flipped])
    assert [tuple(t.shape) for t in lr_f] == [(2, 256, 32, 48), (2, 256, 16, 24), (2, 256, 8, 12), (2, 256, 4, 6), (2, 256, 2, 3)]
    assert all(not t.requires_grad and t.dtype == torch.float32 for t in lr_f + hr_f)
    assert torch.allclose(hr_f[1][0], out["p3"][0], atol=1e-5)


def test_neck_parameter_names_on_cpu():
    from afigan._compat import ShapeSpec
    from afigan.modeling import FPN_AFIGAN, PAFPN_AFIGAN

    class BU(torch.nn.Module):
        def output_shape(self):
            return {f"res{i + 2}": ShapeSpec(channels=c, stride=2 ** (i + 2)) for i, c in enumerate((256, 512, 1024, 2048))}

    feats = ["res2", "res3", "res4", "res5"]
    f = FPN_AFIGAN(BU(), feats, 256)
    p = PAFPN_AFIGAN(BU(), feats, 256)
    fn, pn = set(dict(f.named_parameters())), set(dict(p.named_parameters()))
    assert {"fpn_lateral2.weight", "fpn_output5.bias", "srf_module.Generators.0.3.0.weight"} <= fn
    assert {"fpn_lateral3.weight", "pafpn_output2.weight", "pafpn_downsample5.weight", "srf_module.Generators.0.1.RDBs.2.conv5.weight"} <= pn
    assert "pafpn_downsample2.weight" not in pn
    assert f.fpn_lateral5.weight.shape == (256, 2048, 1, 1)
    assert sum(q.numel() for q in f.srf_module.parameters()) == 7_834_624


def _ddp_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "afi-gan_b200"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from afigan.engine.sync import FlatGradSync
    from oracle import afigan_oracle as O
    torch.set_num_threads(2)
    g_sd, d_sd = O.init_states(0)
    lr_shapes, hr_shapes = ((4, 6),), ((7, 11),)
    lr_f, hr_f = O.synthetic_features(1, rank, lr_shapes, hr_shapes)             # each rank its own shard (seed 1234 + rank)
    res = O.stage1_step(dict(g_sd), {k: v.clone() for k, v in d_sd.items()}, lr_f, hr_f, lr=None)
    keys = O.generator_param_keys()
    params = [torch.nn.Parameter(g_sd[k].clone()) for k in keys]
    sync = FlatGradSync(params)
    for p, k in zip(params, keys):
        assert p.grad.data_ptr() == sync.views[keys.index(k)].data_ptr()          # .grad are views into ONE flat buffer
        p.grad.copy_(res["g_grads"][k])
    sync.all_reduce()
    assert sync.grad_scale == 1.0 / world
    avg = {k: (p.grad * sync.grad_scale).clone() for p, k in zip(params, keys)}
    if rank == 0:
        torch.save({"avg": avg, "own": res["g_grads"]}, out)
    else:
        torch.save({"own": res["g_grads"]}, out + ".1")
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo(tmp_path):
    """Intended DDP semantics: after the flat all-reduce every rank holds the mean of the per-shard oracle gradients."""
    out = str(tmp_path / "r0.pt")
    mp.spawn(_ddp_worker, args=(2, 29631, out), nprocs=2, join=True)
    r0, r1 = torch.load(out), torch.load(out + ".1")
    for k, v in r0["avg"].items():
        ref = 0.5 * (r0["own"][k] + r1["own"][k])
        assert float((v - ref).norm()) <= 1e-6 * float(ref.norm()), k
        assert float((r0["own"][k] - r1["own"][k]).norm()) > 1e-2 * float(ref.norm())     # the shards really differ


def _bcast_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "afi-gan_b200"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from afigan.engine.sync import FlatGradSync
    torch.manual_seed(100 + rank)                                       # detectron2 seeds each rank differently
    params = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7))]
    before = [p.detach().clone() for p in params]
    mom = [torch.randn(5, 3), torch.randn(7)]
    sync = FlatGradSync(params)                                         # broadcasts rank 0's parameters, like DDP's constructor
    sync.broadcast_parameters(mom)                                      # ... and on resume the momentum buffers
    sync.flat.fill_(float(rank + 1))
    sync.start()                                                        # asynchronous all-reduce
    sync.finish()
    torch.save({"before": before, "after": [p.detach().clone() for p in params], "mom": mom, "flat": sync.flat.clone()}, f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_initial_state_is_broadcast_from_rank0_world2_gloo(tmp_path):
    """replicas that start from different seeds must be made identical before averaged gradients are applied (DDP's constructor
    broadcast, reference stage1_trainer.py:80-89); the asynchronous all-reduce sums like the blocking one."""
    out = str(tmp_path / "b")
    mp.spawn(_bcast_worker, args=(2, 29633, out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    for a, b, c0, c1 in zip(r0["after"], r1["after"], r0["before"], r1["before"]):
        assert torch.equal(a, b) and torch.equal(a, c0) and not torch.equal(c0, c1)
    for a, b in zip(r0["mom"], r1["mom"]):
        assert torch.equal(a, b)
    assert torch.equal(r0["flat"], torch.full_like(r0["flat"], 3.0)) and torch.equal(r1["flat"], r0["flat"])


def test_checkpoint_key_remap_and_lr_schedule():
    from afigan._compat import ShapeSpec
    from afigan.engine import (convert_afi_names, load_extractor_into_detector, load_generator_into_extractor, remain_only_afi_names,
                               warmup_multistep_lr)
    from afigan.modeling import FPN_AFIGAN, Generator

    class BU(torch.nn.Module):
        def output_shape(self):
            return {f"res{i + 2}": ShapeSpec(channels=c, stride=2 ** (i + 2)) for i, c in enumerate((256, 512, 1024, 2048))}

    class Detector(torch.nn.Module):          # stands for GeneralizedRCNN_AFExtractor: the neck lives under `backbone.`
        def __init__(self):
            super().__init__()
            self.backbone = FPN_AFIGAN(BU(), ["res2", "res3", "res4", "res5"], 256)

    torch.manual_seed(7)
    G = Generator(n_residual_dense_blocks=3)
    ckpt = {"module." + k: v.clone() for k, v in G.state_dict().items()}          # saved from a DDP-wrapped stage-1 model
    renamed, back = convert_afi_names({k[7:]: v for k, v in ckpt.items()})
    assert "backbone.srf_module.Generators.0.3.0.weight" in renamed and back["backbone.srf_module.Generators.0.0.0.bias"] == "Generators.0.0.0.bias"
    det = Detector()
    assert not torch.equal(det.backbone.srf_module.Generators[0][0][0].weight, G.Generators[0][0][0].weight)
    assert load_generator_into_extractor(det, ckpt) == len(G.state_dict())          # stage 1 -> stage 2
    for k, v in G.state_dict().items():
        assert torch.equal(det.state_dict()["backbone.srf_module." + k], v), k
    kept, _ = remain_only_afi_names(det.state_dict())
    assert len(kept) == len(G.state_dict()) and all("srf_module" in k for k in kept)
    det3 = Detector()
    lateral_before = det3.backbone.fpn_lateral2.weight.clone()
    assert load_extractor_into_detector(det3, det.state_dict()) == len(G.state_dict())   # stage 2 -> stage 3: only the interpolator moves
    assert torch.equal(det3.backbone.fpn_lateral2.weight, lateral_before)
    assert torch.equal(det3.backbone.srf_module.Generators[0][4][0].weight, G.Generators[0][4][0].weight)
    # WarmupMultiStepLR of the stage-1 recipe
    assert abs(warmup_multistep_lr(0) - 1e-6) < 1e-12 and abs(warmup_multistep_lr(500) - 1e-3 * (0.001 * 0.5 + 0.5)) < 1e-12
    assert warmup_multistep_lr(1000) == 1e-3 and warmup_multistep_lr(269999) == 1e-3 and abs(warmup_multistep_lr(270000) - 1e-4) < 1e-12


def test_bench_flop_accounting_matches_the_survey():
    """bench.py's op counts: G forward 19 206 144 FLOP per input pixel, D forward 30 689 280 per pixel, a config-1 stage-1 step
    2.332e13 FLOP per image; the step that evaluates G(lr) once executes exactly one generator forward less."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    assert bench.G_FWD_FLOP_PER_INPUT_PX == 2 * 1_179_648 + 3 * 2_469_888 + 4_718_592 + 4_718_592 == 19_206_144
    assert bench.D_FWD_FLOP_PER_PX == 2_359_296 + 9_437_184 + 18_874_368 + 18_432 == 30_689_280
    lr_px, hr_px = 2 * 23_282, 2 * 89_523
    ref = bench.stage1_step_flops(lr_px, hr_px)
    assert abs(ref / 2 - 2.332e13) < 0.001e13                         # per image (batch 2)
    assert ref - bench.stage1_step_flops(lr_px, hr_px, 1) == bench.G_FWD_FLOP_PER_INPUT_PX * lr_px
    assert sum(h * w for h, w in bench.C1_LR_SHAPES) == 23_282 and sum(h * w for h, w in bench.C1_HR_SHAPES) == 89_523
