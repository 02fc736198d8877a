"""`python -m probabilistic_domain_adaptation_b200.run <script>` makes an UNCHANGED script bind this package's model, the
fused trainer helpers and the fused prediction functions.  A miniature stand-in tree (tests/standin.py) plays the
reference checkout: its my_models package raises on import and its prediction functions raise when called, to prove they
are bypassed; its trainer classes carry the reference's method names."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import standin  # noqa: E402


def _run(ref, script, *argv, timeout=600):
    # (the checkout itself is on PYTHONPATH only so that a test's sitecustomize.py is picked up at interpreter start)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, str(ref), os.environ.get("PYTHONPATH", "")]))
    return subprocess.run([sys.executable, "-m", "probabilistic_domain_adaptation_b200.run", script, *argv],
                          cwd=str(ref), env=env, capture_output=True, text=True, timeout=timeout)


def test_unchanged_script_binds_the_sm100_model_and_helpers(tmp_path):
    ref = standin.build(tmp_path / "checkout")
    standin._write(os.path.join(ref, "LIVECell", "livecell_mt.py"), """
        import sys
        from prob_utils.my_models import ProbabilisticUnet, l2_regularisation, clean_folder
        from prob_utils.my_trainer import MeanTeacherTrainer, AdaMTTrainer, FixMatchTrainer, AdaMatchTrainer, PUNetTrainer
        from prob_utils.my_predictions import punet_prediction, punet_pseudo_prediction
        import prob_utils.my_predictions.punet_predictions as pp
        import probabilistic_domain_adaptation_b200 as pkg
        assert ProbabilisticUnet is pkg.ProbabilisticUnet, ProbabilisticUnet.__module__
        assert sys.argv[1:] == ["--train", "--consensus"], sys.argv
        m = ProbabilisticUnet(input_channels=1, num_classes=1, num_filters=[64, 128, 256, 512], latent_dim=6,
                              no_convs_fcomb=3, beta=1.0, consensus_masking=True, rl_swap=True)
        assert len(m.state_dict()) == 100
        for cls, helper in ((MeanTeacherTrainer, "sample_from_teacher"), (AdaMTTrainer, "_momentum_update"),
                            (FixMatchTrainer, "sample_from_weak_model"), (AdaMatchTrainer, "sample_from_model")):
            fn = getattr(cls, helper)
            assert fn.__module__ == "probabilistic_domain_adaptation_b200.trainer_mixins", (cls, helper, fn.__module__)
        assert MeanTeacherTrainer._train_epoch_impl(None) == "reference step body"     # step bodies stay reference code
        assert MeanTeacherTrainer.momentum == 0.5                                      # reference defaults are kept
        assert AdaMTTrainer._current_momentum.__qualname__.startswith("FusedAdaMTMixin")
        # punet_trainer.py:15-17 and punet_predictions.py:15-63, 66-136 are on the fused path as well
        assert PUNetTrainer._sample.__module__ == "probabilistic_domain_adaptation_b200.predictions"
        for fn in (punet_prediction, punet_pseudo_prediction, pp.punet_prediction, pp.punet_pseudo_prediction):
            assert fn.__module__ == "probabilistic_domain_adaptation_b200.predictions", fn.__module__
        print("SHIM-OK")
        """)
    out = _run(ref, "LIVECell/livecell_mt.py", "--train", "--consensus", timeout=300)
    assert out.returncode == 0 and "SHIM-OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_unchanged_predict_scripts_run_on_the_fused_path(tmp_path):
    """`lung_punet.py --predict` (Lung-XRay/lung_punet.py:91-127) and a `--get_pseudo_labels` call
    (LIVECell/livecell_punet_target.py:45-53), unchanged, through run.py: the outputs equal what the package computes
    directly for the same RNG seed, the mask is uint8 {0,1}, and every block batch costs ONE fused Fcomb launch."""
    import numpy as np
    import torch
    from oracle import punet_oracle as po
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    ref = standin.build(tmp_path / "checkout")
    sd = po.make_state_dict(0, last_layer_gain=8.0)
    os.makedirs(os.path.join(ref, "checkpoints", "punet-source-lung-jsrt1"))
    os.makedirs(os.path.join(ref, "checkpoints", "punet-source-livecell-A172"))
    torch.save({"model_state": sd}, os.path.join(ref, "checkpoints", "punet-source-lung-jsrt1", "best.pt"))
    torch.save({"model_state": sd}, os.path.join(ref, "checkpoints", "punet-source-livecell-A172", "best.pt"))
    g = torch.Generator().manual_seed(4)
    lung = [(torch.randn(256, 256, generator=g) * 40 + 120).numpy().astype("float32"),
            (torch.randn(520, 704, generator=g) * 25 + 90).numpy().astype("float32")]     # 1 block and 2 x 2 blocks
    os.makedirs(os.path.join(ref, "data", "jsrt2", "org_test"))
    for i, im in enumerate(lung):
        np.save(open(os.path.join(ref, "data", "jsrt2", "org_test", f"img{i}.png"), "wb"), im)
    cells = (torch.randn(64, 96, generator=g) * 10 + 128).numpy().astype("float32")
    os.makedirs(os.path.join(ref, "data", "images", "livecell_train_val_images"))
    np.save(open(os.path.join(ref, "data", "images", "livecell_train_val_images", "A172_Phase_C7_1.tif"), "wb"), cells)
    # the scripts themselves are untouched; a sitecustomize seeds the RNG and counts fused launches around them
    standin._write(os.path.join(ref, "sitecustomize.py"), """
        import atexit, json, os, torch
        torch.manual_seed(1234)
        from probabilistic_domain_adaptation_b200 import ops
        ops.PROFILE = []
        def _dump():
            kinds = {}
            for k, *_ in ops.PROFILE:
                kinds[k] = kinds.get(k, 0) + 1
            json.dump(kinds, open(os.path.join(os.path.dirname(__file__), "launch_kinds.json"), "w"))
        atexit.register(_dump)
        """)
    import json

    out = _run(ref, "Lung-XRay/lung_punet.py", "--predict")
    assert out.returncode == 0 and "PREDICT-OK" in out.stdout, out.stdout + out.stderr
    kinds = json.load(open(os.path.join(ref, "launch_kinds.json")))
    # img0: one 256 x 256 block -> 1 batch; img1: blocks of two different outer shapes... every batch = ONE fused launch
    from probabilistic_domain_adaptation_b200 import ProbabilisticUnet, consensus, predictions, tiled
    n_batches = 0
    for im in lung:
        groups = {}
        for outer, _ in tiled.blocking(im.shape, (384, 384), (64, 64)):
            groups[(outer[2], outer[3])] = groups.get((outer[2], outer[3]), 0) + 1
        n_batches += sum((n + 7) // 8 for n in groups.values())
    assert kinds.get("fcomb_mc") == n_batches, (kinds, n_batches)
    dev = torch.device("cuda:0")
    model = ProbabilisticUnet(1, 1, [64, 128, 256, 512], 6, 3, 1.0).to(dev).eval()
    model.load_state_dict(sd)
    torch.manual_seed(1234)
    pred_dir = os.path.join(ref, "pred", "punet_source", "source-jsrt1-target-jsrt2")
    from glob import glob
    for path in glob(os.path.join(ref, "data", "jsrt2", "org_test", "*")):   # the script's glob: same listing order
        i = int(os.path.basename(path)[3])
        im = lung[i]
        want = tiled.predict_with_halo(torch.from_numpy(im), model, prior_samples=8).cpu().numpy()
        got = np.load(open(os.path.join(pred_dir, f"img{i}.tif"), "rb"))
        assert got.dtype == np.float64 and got.shape == im.shape
        assert np.array_equal(got.astype(np.float32), want), np.abs(got - want).max()
        assert 0.0 <= got.min() and got.max() <= 1.0 and got.std() > 0

    out = _run(ref, "LIVECell/livecell_pseudo.py", "--get_pseudo_labels")
    assert out.returncode == 0 and "PSEUDO-OK" in out.stdout, out.stdout + out.stderr
    kinds = json.load(open(os.path.join(ref, "launch_kinds.json")))
    assert kinds.get("fcomb_mc") == 1, kinds
    ann = np.load(open(os.path.join(ref, "pseudo", "annotations", "train", "A172", "A172_Phase_C7_1.tif"), "rb"))
    msk = np.load(open(os.path.join(ref, "pseudo", "consensus", "train", "A172", "A172_Phase_C7_1.tif"), "rb"))
    assert ann.dtype == np.float32 and msk.dtype == np.uint8 and ann.shape == msk.shape == cells.shape
    assert set(np.unique(msk).tolist()) <= {0, 1}
    torch.manual_seed(1234)
    patch = predictions.standardize_image(cells, dev)
    mean, mask = consensus.punet_pseudo_labels(model, patch, 16)
    assert np.array_equal(ann, mean.cpu().numpy().squeeze()) and np.array_equal(msk, mask.cpu().numpy().squeeze())
