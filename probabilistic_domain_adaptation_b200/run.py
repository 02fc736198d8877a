"""Runs an UNCHANGED experiment script of the reference on libpda_b200:

    cd <checkout of Probabilistic-Domain-Adaptation>
    python -m probabilistic_domain_adaptation_b200.run LIVECell/livecell_mt.py --train --consensus --masking ...

Before the script starts, `install()`
  * makes `prob_utils.my_models` (and its sub-modules `probabilistic_unet`, `unet`, `unet_blocks`, `utils`) resolve to
    this package's mirror, so `from prob_utils.my_models import ProbabilisticUnet, l2_regularisation, clean_folder`
    (every script and trainer, e.g. LIVECell/livecell_mt.py:8, prob_utils/my_trainer/mean_teacher_trainer.py:12) binds
    the sm_100a implementation -- the reference's own my_models package is never imported;
  * puts the fused helpers on the reference's trainer classes in place (`sample_from_teacher`, `sample_from_weak_model`,
    `sample_from_model`, `_momentum_update`, `PUNetTrainer._sample`: INTEGRATION.md section 2); their step bodies,
    loggers and checkpoints stay reference code on torch_em;
  * replaces `prob_utils.my_predictions.punet_prediction` / `punet_pseudo_prediction` (punet_predictions.py:15-63,
    66-136: what every `--predict` / `--get_pseudo_labels` runs) by `predictions.py`: same signatures and files, the
    per-image work goes through `tiled.predict_with_halo` / `consensus.punet_pseudo_labels` (one fused
    Fcomb + consensus launch per block batch instead of S `sample()` calls per block).
It is the zero-edit alternative to the two `__init__.py` edits described in INTEGRATION.md.
"""
import importlib
import os
import runpy
import sys
import warnings

_MIXINS = {
    "MeanTeacherTrainer": "FusedMeanTeacherMixin",
    "AdaMTTrainer": "FusedAdaMTMixin",
    "FixMatchTrainer": "FusedFixMatchMixin",
    "AdaMatchTrainer": "FusedFixMatchMixin",
}


def _alias_models():
    from . import my_models
    prob_utils = importlib.import_module("prob_utils")          # the reference's (empty) package __init__
    sys.modules["prob_utils.my_models"] = my_models
    for sub in ("probabilistic_unet", "unet", "unet_blocks", "utils"):
        sys.modules[f"prob_utils.my_models.{sub}"] = importlib.import_module(f"{my_models.__name__}.{sub}")
    prob_utils.my_models = my_models
    return my_models


def patch_trainers(trainer_module):
    """Copies the mixins' helper methods onto the reference trainer classes (in place: scripts import the classes by
    name from prob_utils.my_trainer).  Returns the names of the classes that were patched."""
    from . import trainer_mixins
    done = []
    punet_cls = getattr(trainer_module, "PUNetTrainer", None)
    if punet_cls is not None:
        from . import predictions
        punet_cls._sample = predictions.punet_trainer_sample      # punet_trainer.py:15-17
        done.append("PUNetTrainer")
    for cls_name, mixin_name in _MIXINS.items():
        cls = getattr(trainer_module, cls_name, None)
        if cls is None:
            continue
        mixin = getattr(trainer_mixins, mixin_name)
        for klass in reversed(mixin.__mro__[:-1]):               # base mixins first, most derived last
            for name, attr in vars(klass).items():
                if name.startswith("__"):
                    continue
                if name in ("n_samples", "do_consensus_masking", "momentum") and hasattr(cls, name):
                    continue                                     # class-level defaults never shadow the reference's
                setattr(cls, name, attr)
        done.append(cls_name)
    return done


def patch_predictions(pred_module):
    """Binds predictions.punet_prediction / punet_pseudo_prediction into prob_utils.my_predictions (and its
    punet_predictions sub-module, for callers that import from there).  Returns the patched names."""
    from . import predictions
    done = []
    sub = sys.modules.get(pred_module.__name__ + ".punet_predictions")
    for name in ("punet_prediction", "punet_pseudo_prediction"):
        if hasattr(pred_module, name) or (sub is not None and hasattr(sub, name)):
            setattr(pred_module, name, getattr(predictions, name))
            if sub is not None:
                setattr(sub, name, getattr(predictions, name))
            done.append(name)
    return done


def install(reference_root=None, patch=True):
    """Aliases the model package and patches the trainers.  `reference_root`: the directory that contains `prob_utils/`
    (default: the current directory)."""
    root = os.path.abspath(reference_root or os.getcwd())
    if not os.path.isdir(os.path.join(root, "prob_utils")):
        raise FileNotFoundError(f"no prob_utils/ under {root}: run from the reference checkout or pass --reference-root")
    if root not in sys.path:
        sys.path.insert(0, root)
    _alias_models()
    patched = []
    if patch:
        try:
            patched = patch_trainers(importlib.import_module("prob_utils.my_trainer"))
        except ImportError as e:                                 # e.g. torch_em missing: prediction-only environments
            warnings.warn(f"prob_utils.my_trainer not importable ({e}); trainers left unpatched")
        try:
            patched += patch_predictions(importlib.import_module("prob_utils.my_predictions"))
        except ImportError as e:                                 # imageio / torch_em missing
            warnings.warn(f"prob_utils.my_predictions not importable ({e}); prediction functions left unpatched")
    return patched


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    root = None
    if argv and argv[0] == "--reference-root":
        root, argv = argv[1], argv[2:]
    if not argv:
        raise SystemExit(__doc__)
    script = argv[0]
    install(root)
    sys.argv = argv
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
