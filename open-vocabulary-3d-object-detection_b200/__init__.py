"""ovdet_b200 -- B200-native (sm_100a) detection-geometry and open-vocabulary
matching path of timsu1104/Open-vocabulary-3D-Object-Detection.

Host side mirrors the reference's Python call surface (``utils.box_util``,
``utils.box_intersection``, ``utils.nms``, ``utils.eval_det``,
``utils.ap_calculator``, ``utils.label_formatter``, ``criterion.Matcher``,
``models.model_3detr.BoxProcessor``); every hot function calls the C-ABI CUDA
library ``lib/libovdet_b200.so`` (see include/ovdet_b200.h).  There is no CPU
fallback: a missing library raises at first use.
"""
__version__ = "0.1.0"
