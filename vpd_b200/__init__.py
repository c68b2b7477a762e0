"""vpd_b200 - B200-native (sm_100a) implementation of the VPD student hot path.

Drop-in mirrors of the reference's Python API for this path:
    vpd_b200.RGBF_EmbeddingModel   <- models/rgb.py:46-86
    vpd_b200.ModelTrainer          <- train_vpd_model.py:53-112
    vpd_b200.apply                 <- apply_vpd_model.py:92-179 (corpus extraction)
    vpd_b200.assemble              <- vpd_dataset/{common,single_frame}.py deterministic part
    vpd_b200.targets               <- GenericDataset.load_default's teacher-target construction
    vpd_b200.train                 <- train_vpd_model.main's epoch loop (+ GPU-assembling loader)
    vpd_b200.augment               <- the reference's augment=True random draws (device kernel K1a)
    vpd_b200.keypoint              <- models/module.py FCResNet, models/keypoint.py (teacher apply)
    vpd_b200.keypoint_train        <- models/keypoint.py epoch, FCPoseDecoder, train_vipe_model loop
    vpd_b200.keypoint_apply        <- apply_vipe_model.py (pose files -> teacher .emb.pkl)
All compute runs in libvpd_b200.so (hand-written CUDA behind the C ABI of
include/vpd_b200.h); importing the model classes without that library raises.
"""

__all__ = ['RGBF_EmbeddingModel', 'ModelTrainer', 'FusedAdamW', 'FusedSGD']


def __getattr__(name):
    if name == 'RGBF_EmbeddingModel':
        from .rgb import RGBF_EmbeddingModel
        return RGBF_EmbeddingModel
    if name in ('ModelTrainer', 'FusedAdamW', 'FusedSGD'):
        from . import trainer
        return getattr(trainer, name)
    raise AttributeError(name)
