"""Import the UNMODIFIED reference (jhong93/vpd) from /root/reference.

TEST INFRASTRUCTURE ONLY. Only usable in the build container (the GPU box has
no /root/reference); used by oracle/gen_golden.py to create tests/golden/* and
by tests that are skipped when the reference is absent.

models/rgb.py:3 imports `efficientnet_pytorch` unconditionally; it is not
installed and the resnet path never touches it, so a stub module is inserted
(SURVEY.md §8c).
"""
import os
import sys
import types
import warnings

REFERENCE_DIR = os.environ.get('VPD_REFERENCE_DIR', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, 'models', 'rgb.py'))


def load():
    """Returns a namespace with the reference's hot-path symbols."""
    if not available():
        raise RuntimeError('reference not present at ' + REFERENCE_DIR)
    sys.dont_write_bytecode = True
    if 'efficientnet_pytorch' not in sys.modules:
        stub = types.ModuleType('efficientnet_pytorch')
        stub.EfficientNet = type('EfficientNet', (), {})
        stub.model = None
        sys.modules['efficientnet_pytorch'] = stub
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    warnings.filterwarnings('ignore')
    from models.rgb import RGBF_EmbeddingModel
    from models.module import FCNet
    from train_vpd_model import ModelTrainer
    from vpd_dataset.single_frame import FrameDataset, GenericDataset
    from vpd_dataset.common import RGB_MEAN_STD
    import vpd_dataset.single_frame as single_frame
    ns = types.SimpleNamespace(
        RGBF_EmbeddingModel=RGBF_EmbeddingModel, FCNet=FCNet,
        ModelTrainer=ModelTrainer, FrameDataset=FrameDataset,
        GenericDataset=GenericDataset, RGB_MEAN_STD=RGB_MEAN_STD,
        single_frame=single_frame)
    return ns
