"""A miniature stand-in for a checkout of the reference repository (test infrastructure).

The real reference cannot be imported on the GPU box (/root/reference does not travel) and its trainers / prediction
modules need torch_em and imageio, which are not installed anywhere here.  The stand-in carries what `run.install()`
touches, with the reference's names and call shapes:

  prob_utils/my_models/__init__.py        raises on import           (proves the reference model is bypassed)
  prob_utils/my_trainer/__init__.py       trainer classes with the reference's helper-method names
  prob_utils/my_predictions/__init__.py   punet_prediction / punet_pseudo_prediction that raise when called
                                          (proves the reference prediction path is bypassed), imported the way
                                          prob_utils/my_predictions/__init__.py:1-2 does
  imageio/v3.py                           imread / imwrite on .npy payloads (imageio is not installed)
  Lung-XRay/lung_punet.py                 the `--predict` branch of Lung-XRay/lung_punet.py:91-127, argument for argument
  LIVECell/livecell_pseudo.py             a `--get_pseudo_labels` call as LIVECell/livecell_punet_target.py makes it
"""
import os
import textwrap


def _write(path, text):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fh:
        fh.write(textwrap.dedent(text))


def build(root):
    root = str(root)
    _write(os.path.join(root, "prob_utils", "__init__.py"), "")
    _write(os.path.join(root, "prob_utils", "my_models", "__init__.py"),
           "raise ImportError('reference my_models imported')\n")
    _write(os.path.join(root, "prob_utils", "my_trainer", "__init__.py"), """
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
        class PUNetTrainer(_Base):
            def _sample(self, n_samples=16): return "reference"
        """)
    _write(os.path.join(root, "prob_utils", "my_predictions", "punet_predictions.py"), """
        from prob_utils.my_models import clean_folder                # as punet_predictions.py:12
        def punet_prediction(input_image_path, output_pred_path, model, prior_samples=8, device="cpu", mysig=None):
            raise RuntimeError("reference punet_prediction called")
        def punet_pseudo_prediction(input_image_path, output_pred_path, model, prior_samples=8, device="cpu",
                                    cellname_=None, split_name=None):
            raise RuntimeError("reference punet_pseudo_prediction called")
        """)
    _write(os.path.join(root, "prob_utils", "my_predictions", "__init__.py"),
           "from .punet_predictions import punet_prediction, punet_pseudo_prediction\n")
    _write(os.path.join(root, "imageio", "__init__.py"), "")
    _write(os.path.join(root, "imageio", "v3.py"), """
        import numpy as np
        def imread(path):
            with open(path, "rb") as fh:
                return np.load(fh, allow_pickle=False)
        def imwrite(path, array, **kwargs):
            with open(path, "wb") as fh:
                np.save(fh, np.asarray(array), allow_pickle=False)
        """)
    # Lung-XRay/lung_punet.py:91-127 (`--predict`): model hyper-parameters, checkpoint key, glob and call as there
    _write(os.path.join(root, "Lung-XRay", "lung_punet.py"), """
        import argparse, os
        import torch
        from prob_utils.my_models import ProbabilisticUnet
        from prob_utils.my_predictions import punet_prediction

        def do_punet_predictions(device, data_path, pred_path):
            model = ProbabilisticUnet(input_channels=1, num_classes=1, num_filters=[64, 128, 256, 512], latent_dim=6,
                                      no_convs_fcomb=3, beta=1.0, rl_swap=False)
            model_save_dir = "checkpoints/punet-source-lung-jsrt1/best.pt"
            model_state = torch.load(model_save_dir, map_location=torch.device("cpu"))["model_state"]
            model.load_state_dict(model_state)
            model.to(device)
            output_path = pred_path + "punet_source/source-jsrt1-target-jsrt2/"
            input_path = data_path + "jsrt2/org_test/*"
            punet_prediction(input_image_path=input_path, output_pred_path=output_path, model=model, device=device)

        if __name__ == "__main__":
            ap = argparse.ArgumentParser()
            ap.add_argument("--predict", action="store_true")
            ap.add_argument("-i", "--data", default="data/")
            ap.add_argument("-o", "--pred", default="pred/")
            args = ap.parse_args()
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
            if args.predict:
                do_punet_predictions(device, args.data, args.pred)
                print("PREDICT-OK")
        """)
    _write(os.path.join(root, "LIVECell", "livecell_pseudo.py"), """
        import argparse
        import torch
        from prob_utils.my_models import ProbabilisticUnet
        from prob_utils.my_predictions import punet_pseudo_prediction

        if __name__ == "__main__":
            ap = argparse.ArgumentParser()
            ap.add_argument("--get_pseudo_labels", action="store_true")
            args = ap.parse_args()
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
            model = ProbabilisticUnet(input_channels=1, num_classes=1, num_filters=[64, 128, 256, 512], latent_dim=6,
                                      no_convs_fcomb=3, beta=1.0, rl_swap=True)
            model.load_state_dict(torch.load("checkpoints/punet-source-livecell-A172/best.pt",
                                             map_location=torch.device("cpu"))["model_state"])
            model.to(device)
            if args.get_pseudo_labels:
                punet_pseudo_prediction(input_image_path="data/images/livecell_train_val_images/", output_pred_path="pseudo/",
                                        model=model, prior_samples=16, device=device, cellname_="A172",
                                        split_name="train")
                print("PSEUDO-OK")
        """)
    return root
