"""CPU: `python -m probabilistic_domain_adaptation_b200.run <script>` makes an UNCHANGED script bind this package's
model and the fused trainer helpers.  A miniature stand-in tree plays the reference checkout (its my_models package raises
on import, to prove it is bypassed; its trainer classes carry the reference's method names)."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write(path, text):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fh:
        fh.write(textwrap.dedent(text))


def test_unchanged_script_binds_the_sm100_model_and_helpers(tmp_path):
    ref = tmp_path / "checkout"
    _write(str(ref / "prob_utils" / "__init__.py"), "")
    _write(str(ref / "prob_utils" / "my_models" / "__init__.py"), "raise ImportError('reference my_models imported')\n")
    _write(str(ref / "prob_utils" / "my_trainer" / "__init__.py"), """
        from prob_utils.my_models import l2_regularisation          # as mean_teacher_trainer.py:12
        class _Base:
            n_samples = 16
        class MeanTeacherTrainer(_Base):
            momentum = 0.5
            def sample_from_teacher(self, x): return "reference"
            def sample_from_model(self): return "reference"
            def _momentum_update(self): return "reference"
            def _train_epoch_impl(self): return "reference step body"
        class AdaMTTrainer(MeanTeacherTrainer): pass
        class FixMatchTrainer(_Base):
            def sample_from_weak_model(self, x): return "reference"
        class AdaMatchTrainer(FixMatchTrainer): pass
        class PUNetTrainer(_Base): pass
        """)
    _write(str(ref / "LIVECell" / "livecell_mt.py"), """
        import sys
        from prob_utils.my_models import ProbabilisticUnet, l2_regularisation, clean_folder
        from prob_utils.my_trainer import MeanTeacherTrainer, AdaMTTrainer, FixMatchTrainer, AdaMatchTrainer
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
        print("SHIM-OK")
        """)
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    out = subprocess.run([sys.executable, "-m", "probabilistic_domain_adaptation_b200.run", "LIVECell/livecell_mt.py",
                          "--train", "--consensus"], cwd=str(ref), env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "SHIM-OK" in out.stdout, out.stdout + out.stderr
