"""Process-wide switches of the drop-in layer.

nms_semantics selects which torchvision device the NMS arithmetic reproduces bit-for-bit
(SURVEY.md section 8c): "cuda" = what the reference does when its tensors live on the GPU
(coordinate trick up to 25 000 candidates, FMA IoU order, fp32 threshold compare); "cpu" = what
predict.py does (trick up to 1 000 candidates, non-FMA order, fp64 threshold compare).
"""
nms_semantics = "cuda"

# model/loss.py:110-114 raises RuntimeError('NaN in loss') inside every loss_per_scale call, which
# costs the reference one host sync per FPN level per step.  "sync" keeps that behaviour exactly;
# "lazy" records the flag on the returned tensors (checked by DetectionHead once per step); "off"
# never reads it back.
nan_check = "sync"


def nms_modes():
    if nms_semantics == "cuda":
        return "auto_cuda", "tv_cuda"
    if nms_semantics == "cpu":
        return "auto_cpu", "tv_cpu"
    raise ValueError("nms_semantics must be 'cuda' or 'cpu'")
