"""The pseudo-label sweep of the reference's ``generate_pseudo_label.py`` without the network: what happens between
``inference`` collecting the box predictions (engine.py:254-302, ``label_formatter.step`` per batch) and
``label_formatter.process(args.topk, args.conf_thresh, args.obj_thresh)`` (generate_pseudo_label.py:209), with the
reference's three knobs (:164-166: ``--topk 50 --conf_thresh 0 --obj_thresh 0``).

The forward pass itself (3DETR + pointnet2) is out of scope; a caller feeds the per-batch prediction
dicts it would have produced."""
from .utils.label_formatter import LabelFormatter


def sweep(batches, scene_list, box_path, output_path, label_path, topk=50, conf_thresh=0.0, obj_thresh=0.0, distributed=False):
    """``batches``: iterable of ``(outputs, batch_data_label)`` as ``engine.inference`` passes them to
    ``LabelFormatter.step`` (outputs: sem_cls_prob [B,Q,C], objectness_prob [B,Q], center_unnormalized / size_unnormalized
    [B,Q,3]; batch_data_label: scan_idx [B]).  Writes ``<scan>_bbox.npy`` (fp64 [N,7] = centre, size, label) per scan of
    ``scene_list`` (this rank's share of them when ``distributed``) under ``output_path`` and returns the number of
    pseudo-label boxes acquired over all ranks.  ``topk`` is accepted for interface parity: the reference's top-k
    selection is commented out (utils/label_formatter.py:128-130), only the two thresholds act."""
    fmt = LabelFormatter(box_path, output_path, label_path, scene_list)
    for outputs, batch_data_label in batches:
        if "outputs" in outputs:
            outputs = outputs["outputs"]
        fmt.step(outputs, batch_data_label)
    return fmt.process(topk, conf_thresh, obj_thresh, distributed=distributed)
