"""Mixins that put the reference's trainer classes on the fused kernels WITHOUT editing their step bodies.

The reference trainers (prob_utils/my_trainer/*.py) subclass torch_em's DefaultTrainer and call three helper
methods from `_train_epoch_impl`: `sample_from_teacher` / `sample_from_weak_model`, `sample_from_model` and
`_momentum_update`.  Listing one of these mixins BEFORE the reference class in a subclass' bases overrides exactly
those helpers (same signatures, same return types), e.g.

    class MeanTeacherTrainer(FusedMeanTeacherMixin, prob_utils.my_trainer.MeanTeacherTrainer): pass

Attributes read are the ones the reference trainers already own: model, teacher, n_samples, do_consensus_masking,
momentum, _iteration (mean_teacher_trainer.py:19-50, fixmatch_trainer.py:19-35, adamt_trainer.py:20-43).
"""
from . import consensus


class _FusedSamplingMixin:
    n_samples = 16
    do_consensus_masking = False

    def sample_from_model(self):
        """mean_teacher_trainer.py:90-93 (+ copies): mean of n_samples sigmoid samples of the forwarded model."""
        return consensus.sample_from_model(self.model, self.n_samples)


class FusedMeanTeacherMixin(_FusedSamplingMixin):
    """MeanTeacherTrainer (mean_teacher_trainer.py:52-55, 72-93)."""
    momentum = 0.999
    _pda_ema = None

    def sample_from_teacher(self, teacher_inputs, upper_thres=0.9, lower_thres=0.1):
        return consensus.sample_from_teacher(self.teacher, teacher_inputs, self.n_samples, upper_thres, lower_thres,
                                             self.do_consensus_masking)

    def _current_momentum(self):
        return self.momentum

    def _momentum_update(self):
        if self._pda_ema is None or self._pda_ema.model is not self.model or self._pda_ema.teacher is not self.teacher:
            self._pda_ema = consensus.MomentumUpdater(self.model, self.teacher)
        self._pda_ema.step(self._current_momentum())


class FusedAdaMTMixin(FusedMeanTeacherMixin):
    """AdaMTTrainer: warm-up momentum min(1 - 1/(it+1), m) (adamt_trainer.py:40-43, 60-76)."""

    def _current_momentum(self):
        return consensus.adamt_momentum(self._iteration, self.momentum)


class FusedFixMatchMixin(_FusedSamplingMixin):
    """FixMatchTrainer / AdaMatchTrainer (fixmatch_trainer.py:37-59, adamatch_trainer.py:33-54)."""

    def sample_from_weak_model(self, weak_inputs, upper_thres=0.9, lower_thres=0.1):
        return consensus.sample_from_weak_model(self.model, weak_inputs, self.n_samples, upper_thres, lower_thres,
                                                self.do_consensus_masking)
