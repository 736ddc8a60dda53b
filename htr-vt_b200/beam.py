"""LM-rescored CTC decoding with the reference's contract (model_window/test_with_kenlm.py:25-59).

`simple_ctc_beam_search_with_lm(log_probs[T,C], converter, lm_scorer, beam_size=5) -> str` keeps the reference's
signature and result; the per-frame beam (T * K^2 Python list operations and a D2H copy per frame per line in the
reference) runs as one kernel for the whole batch (csrc/beam.cu), one D2H copy brings back the K collapsed candidate
id rows per line, and only what must stay on the host stays there: id -> char and `lm_scorer.score(text)` (KenLM in
the reference; any object with a `.score(str) -> float` method).
`beam_search_with_lm_batch(preds_log[T,B,C], ...) -> list[str]` is the loop of validation_with_kenlm (:85-88).

`search="prefix"` (all three entry points) swaps the candidate generator for a true CTC prefix beam search
(csrc/prefix_beam.cu): the K candidates are the K most probable LABELLINGS, each scored with the
summed probability of all its alignments, instead of the K best single alignment paths (which usually collapse to two
or three distinct strings).  The default `search="paths"` is the reference's behaviour, bit for bit.
"""
import numpy as np
import torch

from . import ops


def _candidate_strings(ids, lens, converter, merge_repeats=True):
    """converter.decode(np.array(text), [len(text)]) of the reference (utils.py:72-86) on every already collapsed id
    row at once: decode() filters blanks / repeats / out-of-alphabet ids AGAIN (so a doubled letter that survived the
    path collapse through a blank is merged here - reference behaviour, kept).  ids [R, T], lens [R] -> R strings.
    merge_repeats=False (prefix search): the rows are LABELLINGS, a doubled label is a doubled letter and stays."""
    R, T = ids.shape
    keep = (np.arange(T)[None, :] < lens[:, None]) & (ids != 0) & (ids < len(converter.character))
    if merge_repeats:
        keep[:, 1:] &= ids[:, 1:] != ids[:, :-1]
    counts = keep.sum(1)
    ends = np.cumsum(counts)
    kept = ids[keep]                                          # row-major: row 0's survivors, then row 1's ...
    table = converter.character
    if all(len(ch) == 1 for ch in table[1:]):                 # single code points: one UTF-32 decode for the batch
        lut = np.array([32] + [ord(ch) for ch in table[1:]], dtype=np.uint32)
        flat = lut[kept].astype("<u4", copy=False).tobytes().decode("utf-32-le")
        return [flat[e - c:e] for c, e in zip(counts.tolist(), ends.tolist())]
    kept = kept.tolist()
    return ["".join(table[i] for i in kept[e - c:e]) for c, e in zip(counts.tolist(), ends.tolist())]


def kbest_candidates(log_probs, converter, beam_size=5, lengths=None, layout="tbc", search="paths"):
    """-> per line: list of (string, score) for the surviving beams with a non-empty string, in the reference's
    order (test_with_kenlm.py:42-53).  search: "paths" (the reference's per-frame path beam; score = path log-prob)
    or "prefix" (CTC prefix beam search; score = log-probability of the labelling)."""
    if search not in ("paths", "prefix"):
        raise ValueError("search must be 'paths' or 'prefix'")
    fn = ops.ctc_kbest_paths if search == "paths" else ops.ctc_prefix_beam
    ids, lens, scores = fn(log_probs, beam_size, lengths, layout)
    ids, lens, scores = ids.cpu().numpy(), lens.cpu().numpy(), scores.cpu().numpy()
    B, K, T = ids.shape
    strings = _candidate_strings(ids.reshape(B * K, T), np.maximum(lens.reshape(-1), 0), converter,
                                 merge_repeats=search == "paths")
    sc = scores.tolist()
    alive = (lens >= 0).tolist()
    return [[(strings[b * K + r], sc[b][r]) for r in range(K) if alive[b][r] and strings[b * K + r]] for b in range(B)]


def _pick(cands, lm_scorer):
    if not cands:
        return ""
    lm_scores = [lm_scorer.score(c[0]) for c in cands]
    return cands[int(np.argmax(lm_scores))][0]


def simple_ctc_beam_search_with_lm(log_probs, converter, lm_scorer, beam_size=5, search="paths"):
    """Reference signature: log_probs [T, C] of ONE line -> best string by LM score."""
    if not log_probs.is_cuda:
        raise ops.HtrvtError("simple_ctc_beam_search_with_lm needs a CUDA tensor (no CPU fallback)")
    lp = log_probs.float().unsqueeze(1)                       # [T, 1, C]
    return _pick(kbest_candidates(lp, converter, beam_size, search=search)[0], lm_scorer)


def beam_search_with_lm_batch(preds_log, converter, lm_scorer, beam_size=5, lengths=None, search="paths"):
    """preds_log [T, B, C] (validation_with_kenlm's `preds_log`) -> list of B strings: one kernel, one D2H copy."""
    if not preds_log.is_cuda:
        raise ops.HtrvtError("beam_search_with_lm_batch needs a CUDA tensor (no CPU fallback)")
    lp = preds_log.float()
    if lp.stride(-1) != 1:
        lp = lp.contiguous()
    return [_pick(c, lm_scorer) for c in kbest_candidates(lp, converter, beam_size, lengths, search=search)]
