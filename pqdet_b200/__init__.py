"""pqdet_b200 -- B200-native (sm_100a) implementation of PQDet's detection hot path.

Public surface = the reference's own call signatures:

    parser.Decode / parser.YOLOLayer          model/parser.py:194-249
    loss.loss_per_scale                       model/loss.py:22-115
    interpreter.DetectionHead / _TARGET_MAP   model/interpreter.py:16-20, 72-85
    tools.torch_nms / iou_calc3 / giou / ...  tools.py:335-566
    base_sample.recover_bboxes_prediction_*   dataset/base_sample.py:98-139
    train_dataset.LabelAssigner.create_label  dataset/train_dataset.py:109-150
    fused.decode_nms / decode_nms_host        the whole eval post-process in one kernel (device / pinned-host buffers)
    dist.*                                    shard-by-image multi-GPU plumbing

and the steps either side of it:

    train_dataset.SparseTarget                GT lists + owner maps instead of the dense label tensors
    interpreter.DetectionHead.loss_and_grad   loss + d loss/d head without autograd glue
    interpreter.DetectionHead.forward_from_features   1x1 head conv + decode on the tensor cores (tcgen05, TF32)
    evaluator.DetectionAccumulator            eval/evaluator.py:31-36, 64-183 (add_detections / add_labels / AP)
    augment.letterbox_normalize / Resize      dataset/augment.py:206-259, 390-398 (eval pre-processing)

Everything runs in hand-written CUDA kernels behind the C ABI in include/pqdet_b200.h; there is
no CPU or PyTorch fallback -- a missing extension or a CPU tensor raises.
"""
__version__ = "0.1.0"

from . import _lib, config  # noqa: F401


def load_library():
    """Load the CUDA extension (raises PqdetError with build instructions if it is missing)."""
    return _lib.load()
