"""Import the UNMODIFIED reference (jhong93/vpd): from /root/reference where it exists
(the build container), else from the verbatim copy `oracle/_ref/` that
`oracle/build_ref.py` makes (git-ignored; it travels to the GPU box with the snapshot).

TEST / MEASUREMENT INFRASTRUCTURE ONLY: used by oracle/gen_golden.py to create
tests/golden/*, by tests that are skipped when no reference is importable, and by
bench.py's CPU arm (`--impl reference`, `cpu_baseline`).

models/rgb.py:3 imports `efficientnet_pytorch` unconditionally; it is not
installed and the resnet path never touches it, so a stub module is inserted
(SURVEY.md §8c).
"""
import os
import sys
import types
import warnings

REFERENCE_DIR = os.environ.get('VPD_REFERENCE_DIR', '/root/reference')
if not os.path.isfile(os.path.join(REFERENCE_DIR, 'models', 'rgb.py')):
    REFERENCE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, 'models', 'rgb.py'))


def source():
    """'reference' (the tree under /root/reference), 'copy' (oracle/_ref) or None"""
    if not available():
        return None
    return 'copy' if REFERENCE_DIR.endswith('_ref') else 'reference'


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
