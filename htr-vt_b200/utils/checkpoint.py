"""Checkpoint compatibility with the reference's wire format.

The reference saves `{'model': model.state_dict(), 'state_dict_ema': model_ema.ema.state_dict(), 'optimizer': ...,
'nb_iter', 'best_cer', 'best_wer', 'args', RNG states, ...}` (model_v1/train.py:153-172) and loads for evaluation with
`ckpt['state_dict_ema']`, stripping the `module.` prefix a DataParallel wrapper may have added, then
`load_state_dict(strict=True)` (model_v1/test.py:29-40).  The state_dict schema of this package's modules is the
reference's (keys, shapes, order), so upstream `best_CER.pth` files load as they are; these helpers are the few lines
of test.py:29-40 / utils.load_checkpoint without the logger plumbing.
"""
import re
from collections import OrderedDict

import torch

_MODULE = re.compile(r"^module\.")


def strip_module_prefix(state_dict):
    """`module.` prefixes removed (test.py:33-37 uses re.sub on every key that mentions "module")."""
    return OrderedDict((_MODULE.sub("", k), v) for k, v in state_dict.items())


def extract_state_dict(ckpt, prefer_ema=True):
    """Pick the weights out of whatever torch.load returned: a reference checkpoint dict ('state_dict_ema' / 'model'),
    a bare state_dict, or a {'state_dict': ...} wrapper."""
    if not isinstance(ckpt, dict):
        raise TypeError("expected a checkpoint dict, got %r" % type(ckpt))
    order = ("state_dict_ema", "model", "state_dict") if prefer_ema else ("model", "state_dict", "state_dict_ema")
    for key in order:
        if key in ckpt and isinstance(ckpt[key], dict):
            return strip_module_prefix(ckpt[key])
    if ckpt and all(torch.is_tensor(v) for v in ckpt.values()):
        return strip_module_prefix(ckpt)
    raise KeyError("no state_dict found (looked for %s)" % ", ".join(order))


def load_reference_checkpoint(model, path_or_ckpt, prefer_ema=True, strict=True, map_location="cpu"):
    """model.load_state_dict(<reference checkpoint>) as model_v1/test.py:29-40 does.  Returns the checkpoint dict
    (so callers can read 'nb_iter', 'best_cer', ... as utils.load_checkpoint does)."""
    ckpt = path_or_ckpt
    if not isinstance(ckpt, dict):
        ckpt = torch.load(path_or_ckpt, map_location=map_location, weights_only=False)
    model.load_state_dict(extract_state_dict(ckpt, prefer_ema), strict=strict)
    return ckpt


def save_reference_checkpoint(path, model, model_ema=None, optimizer=None, **extra):
    """Write the reference's checkpoint layout (train.py:153-172): 'model', 'state_dict_ema', 'optimizer' + extras."""
    ckpt = {"model": model.state_dict()}
    if model_ema is not None:
        ckpt["state_dict_ema"] = (model_ema.ema if hasattr(model_ema, "ema") else model_ema).state_dict()
    if optimizer is not None:
        ckpt["optimizer"] = optimizer.state_dict()
    ckpt.update(extra)
    torch.save(ckpt, path)
    return ckpt
