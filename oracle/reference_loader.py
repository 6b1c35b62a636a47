"""Import the UNMODIFIED reference from /root/reference (build container only) -- TEST INFRASTRUCTURE ONLY.

The reference needs `timm` and `ftfy`; oracle/refstubs/ supplies minimal stand-ins.  Nothing on
the GPU box may call this: /root/reference does not exist there.
"""
import logging
import os
import sys

import torch

REFERENCE_ROOT = "/root/reference/segmentation"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "denseclip"))


def import_reference():
    if not reference_available():
        raise RuntimeError("reference not present at " + REFERENCE_ROOT)
    stubs = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refstubs")
    for p in (stubs, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    lvl = logging.root.manager.disable
    logging.disable(logging.CRITICAL)
    try:
        import denseclip as ref  # noqa: F401  (the reference package)
    finally:
        logging.disable(lvl)
    return ref


def load_reference_denseclip(cfg: dict, seed: int):
    """Build the reference DenseCLIP from constructor kwargs and load oracle.seeded_state_dict weights. -> (model, shapes)"""
    from .denseclip_oracle import seeded_state_dict
    ref = import_reference()
    lvl = logging.root.manager.disable
    logging.disable(logging.CRITICAL)
    try:
        import copy
        model = ref.DenseCLIP(**copy.deepcopy(cfg), clip_pretrained_path=None)
    finally:
        logging.disable(lvl)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = seeded_state_dict(shapes, seed)
    model.load_state_dict(sd, strict=True)
    model.eval()
    return model, shapes
