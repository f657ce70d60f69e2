"""Checkpoint interop on REAL files (reference afigan/engine/checkpoint.py:64-125): weights trained by this package load
into the UNMODIFIED reference modules and vice versa, through `torch.save({"model": ...})` files, incl. the stage 1 -> 2 -> 3 key remaps."""
import os
import pickle
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_modules():
    sys.path.insert(0, ROOT)
    from oracle import build_ref, ref_runner
    build_ref.build(verbose=False)                      # stages oracle/_ref when /root/reference is present (dev container)
    if not ref_runner.available():
        pytest.skip("oracle/_ref is not staged (needs /root/reference once)")
    return ref_runner.load_modules()


def test_generator_and_discriminator_files_round_trip_with_the_reference_modules(tmp_path):
    from afigan.engine import AFCheckpointer
    from afigan.modeling import Discriminator, Generator
    gen_mod, dis_mod = _ref_modules()
    # ---- ours -> file -> reference (strict): stage-1 checkpoints G_0/model_final.pth, D_0/model_final.pth (stage1_trainer.py:129-150)
    torch.manual_seed(3)
    G, D = Generator(n_residual_dense_blocks=3), Discriminator()
    D.Discriminators[0][1][0].norm.running_mean.normal_()
    D.Discriminators[0][2][0].norm.num_batches_tracked.fill_(20)
    sgd = torch.optim.SGD(G.parameters(), lr=0.1, momentum=0.9)
    pg = AFCheckpointer(torch.nn.parallel.DataParallel(G) if False else G, str(tmp_path / "G_0"), optimizer=sgd).save("model_final", iteration=299999)
    pd = AFCheckpointer(D, str(tmp_path / "D_0")).save("model_final", iteration=299999)
    ck = torch.load(pg, weights_only=False)
    assert set(ck) == {"model", "optimizer", "iteration"} and ck["iteration"] == 299999
    rG, rD = gen_mod.Generator(n_residual_dense_blocks=3), dis_mod.Discriminator()
    rG.load_state_dict(torch.load(pg, weights_only=False)["model"], strict=True)
    rD.load_state_dict(torch.load(pd, weights_only=False)["model"], strict=True)
    for (k, a), (k2, b) in zip(G.state_dict().items(), rG.state_dict().items()):
        assert k == k2 and torch.equal(a, b)
    for (k, a), (k2, b) in zip(D.state_dict().items(), rD.state_dict().items()):
        assert k == k2 and torch.equal(a, b), k
    assert int(rD.state_dict()["Discriminators.0.2.0.norm.num_batches_tracked"]) == 20
    # ---- reference -> file (DDP-style `module.` prefix, as fvcore would meet it) -> ours
    torch.manual_seed(4)
    rG2, rD2 = gen_mod.Generator(n_residual_dense_blocks=3), dis_mod.Discriminator()
    torch.save({"model": {"module." + k: v for k, v in rG2.state_dict().items()}, "iteration": 7}, tmp_path / "ref_g.pth")
    torch.save({"model": rD2.state_dict()}, tmp_path / "ref_d.pth")
    G2, D2 = Generator(n_residual_dense_blocks=3), Discriminator()
    cg = AFCheckpointer(G2)
    rest = cg.load(str(tmp_path / "ref_g.pth"))
    assert rest == {"iteration": 7} and cg.last_incompatible == {"missing_keys": [], "unexpected_keys": [], "incorrect_shapes": []}
    AFCheckpointer(D2).load(str(tmp_path / "ref_d.pth"))
    for k, v in rG2.state_dict().items():
        assert torch.equal(G2.state_dict()[k], v), k
    for k, v in rD2.state_dict().items():
        assert torch.equal(D2.state_dict()[k], v), k
    # ---- resume_or_load: last_checkpoint bookkeeping
    c = AFCheckpointer(Generator(n_residual_dense_blocks=3), str(tmp_path / "G_0"))
    assert c.has_checkpoint() and c.get_checkpoint_file().endswith("model_final.pth")
    assert c.resume_or_load("", resume=True)["iteration"] == 299999
    assert torch.equal(c.model.state_dict()["Generators.0.4.0.weight"], G.state_dict()["Generators.0.4.0.weight"])


def test_stage_1_to_2_to_3_remaps_on_files(tmp_path):
    """`_load_AFExtractor_weights_file` (Generators.* -> backbone.srf_module.Generators.*, suffix-matched) and `_load_TargetDetector_weights_file`
    (keep only srf_module) on files, starting from a checkpoint of the REFERENCE generator; model-zoo style .pkl input too."""
    from afigan._compat import ShapeSpec
    from afigan.engine import AFCheckpointer
    from afigan.modeling import FPN_AFIGAN, PAFPN_AFIGAN
    gen_mod, _ = _ref_modules()

    class BU(torch.nn.Module):
        def output_shape(self):
            return {f"res{i + 2}": ShapeSpec(channels=c, stride=2 ** (i + 2)) for i, c in enumerate((256, 512, 1024, 2048))}

    class Detector(torch.nn.Module):
        def __init__(self, neck):
            super().__init__()
            self.backbone = neck(BU(), ["res2", "res3", "res4", "res5"], 256)

    torch.manual_seed(5)
    rG = gen_mod.Generator(n_residual_dense_blocks=3)
    torch.save({"model": rG.state_dict(), "iteration": 1}, tmp_path / "stage1_G.pth")
    with open(tmp_path / "stage1_G.pkl", "wb") as f:
        pickle.dump({"model": {k: v.numpy() for k, v in rG.state_dict().items()}, "__author__": "test"}, f)
    for src in ("stage1_G.pth", "stage1_G.pkl"):
        ext = Detector(FPN_AFIGAN)                                               # stage-2 AF extractor (FPN neck)
        n = AFCheckpointer(ext)._load_AFExtractor_weights_file(str(tmp_path / src))
        assert n == len(rG.state_dict())
        for k, v in rG.state_dict().items():
            assert torch.equal(ext.state_dict()["backbone.srf_module." + k], v), k
    AFCheckpointer(ext, str(tmp_path / "stage2")).save("model_final")
    det = Detector(PAFPN_AFIGAN)                                                 # stage-3 target detector with ANOTHER neck
    lateral_before = det.backbone.fpn_lateral3.weight.clone()
    n = AFCheckpointer(det)._load_TargetDetector_weights_file(str(tmp_path / "stage2" / "model_final.pth"))
    assert n == len(rG.state_dict())
    assert torch.equal(det.backbone.fpn_lateral3.weight, lateral_before)          # only the interpolator moved
    for k, v in rG.state_dict().items():
        assert torch.equal(det.state_dict()["backbone.srf_module." + k], v), k
