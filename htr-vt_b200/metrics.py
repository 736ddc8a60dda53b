"""Error rates of the validation loop with the Levenshtein distances computed on the device.

Mirrors model_v1/valid.py:49-75: per (prediction, label) pair a character-level edit distance (CER) and a
word-level one over `format_string_for_wer(...).split(" ")` tokens (WER), accumulated as tot_ED / length_of_gt /
norm_ED exactly as the reference names them.  `editdistance.eval` becomes ONE `htrvt_edit_distance` launch per batch
and level (one warp per pair); `cer_from_ids` never leaves the device: it takes the greedy decoder's id rows
(ops.greedy_decode_ids) and the CTC target stream the loss already consumes.
"""
import re

import torch

from . import ops

# model_v1/utils/utils.py:176-179 (in that non-raw literal `\\(` is an escaped parenthesis, not a backslash member)
_WER_PUNCT = re.compile(r"""([\[\]{}/()"'&+*=<>?.;:,!\-—_€#%°])""")


def format_string_for_wer(s):
    s = _WER_PUNCT.sub(r" \1 ", s)
    return re.sub(r"([ \n])+", " ", s).strip()


def _pack(seqs, device):
    """list of int lists -> (concatenated int32, offsets int32, lengths int32) on `device` (one H2D copy each)."""
    lens = [len(s) for s in seqs]
    offs, acc = [], 0
    for n in lens:
        offs.append(acc)
        acc += n
    flat = [v for s in seqs for v in s] or [0]
    return (torch.tensor(flat, dtype=torch.int32).to(device, non_blocking=True),
            torch.tensor(offs, dtype=torch.int32).to(device, non_blocking=True),
            torch.tensor(lens, dtype=torch.int32))


def edit_distances(a_seqs, b_seqs, device="cuda"):
    """Levenshtein distance of each (a, b) pair of id lists -> python ints (one launch, one D2H copy)."""
    if len(a_seqs) != len(b_seqs):
        raise ValueError("edit_distances needs as many predictions as references")
    if not a_seqs:
        return []
    a, ao, al = _pack(a_seqs, device)
    b, bo, bl = _pack(b_seqs, device)
    d = ops.edit_distance(a, al, b, bl, a_off=ao, b_off=bo, max_b_len=int(bl.max()))
    return d.cpu().tolist()


def cer_from_ids(pred_ids, pred_lens, targets, target_lengths):
    """Device-only CER pieces: pred_ids int32 [B, T] + pred_lens [B] (greedy decode output), targets 1-D int32
    concatenated label ids + target_lengths [B] (the CTC loss inputs).  Returns (distances int32 [B] on device,
    sum of label lengths): CER = distances.sum() / that."""
    tl = target_lengths.to(torch.int32)
    tl_dev = tl.to(pred_ids.device)
    off = (torch.cumsum(tl_dev, 0) - tl_dev).to(torch.int32)
    max_b = int(tl.max()) if tl.device.type == "cpu" else pred_ids.shape[1] * 4
    d = ops.edit_distance(pred_ids, pred_lens, targets, tl_dev, b_off=off, max_b_len=max_b)
    return d, tl_dev.sum()


class ErrorRateMeter(object):
    """The running sums of model_v1/valid.py:12-20, fed one batch of strings at a time (valid.py:49-70)."""

    def __init__(self, device="cuda"):
        self.device = device
        self.norm_ED = 0.0
        self.norm_ED_wer = 0.0
        self.tot_ED = 0
        self.tot_ED_wer = 0
        self.length_of_gt = 0
        self.length_of_gt_wer = 0

    def update(self, preds_str, labels):
        vocab = {}

        def ids(tokens):
            return [vocab.setdefault(t, len(vocab)) for t in tokens]

        pc = [ids(list(p)) for p in preds_str]
        gc = [ids(list(g)) for g in labels]
        for d, g in zip(edit_distances(pc, gc, self.device), labels):
            self.norm_ED += 1 if len(g) == 0 else d / float(len(g))
            self.tot_ED += d
            self.length_of_gt += len(g)
        pw = [ids(format_string_for_wer(p).split(" ")) for p in preds_str]
        gw = [ids(format_string_for_wer(g).split(" ")) for g in labels]
        for d, g in zip(edit_distances(pw, gw, self.device), gw):
            self.norm_ED_wer += 1 if len(g) == 0 else d / float(len(g))
            self.tot_ED_wer += d
            self.length_of_gt_wer += len(g)

    @property
    def CER(self):
        return self.tot_ED / float(self.length_of_gt) if self.length_of_gt else 0.0

    @property
    def WER(self):
        return self.tot_ED_wer / float(self.length_of_gt_wer) if self.length_of_gt_wer else 0.0

    def as_dict(self):
        return dict(norm_ED=self.norm_ED, tot_ED=self.tot_ED, length_of_gt=self.length_of_gt,
                    norm_ED_wer=self.norm_ED_wer, tot_ED_wer=self.tot_ED_wer, length_of_gt_wer=self.length_of_gt_wer,
                    CER=self.CER, WER=self.WER)
