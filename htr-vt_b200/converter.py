"""Label converter with the reference's contract (model_v1/utils/utils.py:55-86).

encode(): identical host logic (ids start at 1, blank 0, the 87-character READ2016 special case).
decode(text_index, length): same inputs/outputs as the reference, but the filtering runs in one
kernel + one D2H copy instead of ~3 device syncs per frame; id -> char stays on the host.
decode_logits(): fused argmax + collapse straight from logits [B,T,C].
"""
import torch

from . import ops


class CTCLabelConverter(object):
    def __init__(self, character, device=None):
        dict_character = list(character)
        self.dict = {ch: i + 1 for i, ch in enumerate(dict_character)}
        if len(self.dict) == 87:
            self.dict['['], self.dict[']'] = 88, 89
        self.character = ['[blank]'] + dict_character
        self.device = torch.device(device) if device is not None else torch.device(
            'cuda' if torch.cuda.is_available() else 'cpu')

    def encode(self, text):
        length = [len(s) for s in text]
        flat = [self.dict[ch] for ch in ''.join(text)]
        return (torch.IntTensor(flat).to(self.device), torch.IntTensor(length).to(self.device))

    def _to_strings(self, ids, lens):
        """ids int32 [B, T] (kept class ids, compacted), lens [B] -> list of str (utils.py:80-84's join)."""
        return self._host_strings(ids.cpu().numpy(), lens.cpu().tolist())

    def _host_strings(self, a, n):
        """a: numpy int32 [B, T] on the host, n: list of B lengths."""
        lut = getattr(self, "_lut", None)
        if lut is None:
            # single-code-point alphabets (all the reference's): id -> UTF-32 code unit, so a whole batch becomes ONE
            # bytes -> str decode + B slices instead of B*T Python-level lookups
            import numpy as np
            ok = all(len(ch) == 1 for ch in self.character[1:])
            self._lut = lut = (np.array([32] + [ord(ch) for ch in self.character[1:]], dtype=np.uint32) if ok else False)
        if lut is not False:
            T = a.shape[1]
            flat = lut[a].astype("<u4", copy=False).tobytes().decode("utf-32-le")
            return [flat[b * T:b * T + n[b]] for b in range(a.shape[0])]
        table = self.character
        return [''.join(table[i] for i in row[:k]) for row, k in zip(a.tolist(), n)]

    def decode(self, text_index, length):
        if not text_index.is_cuda:
            text_index = text_index.to(self.device)
        ids, lens = ops.ctc_collapse(text_index, length, len(self.character))
        return self._to_strings(ids, lens)

    def decode_logits(self, logits, lengths=None):
        ids, lens, _ = ops.greedy_decode_ids(logits.float(), len(self.character), lengths)
        return self._to_strings(ids, lens)

    def decode_logits_async(self, logits, lengths=None):
        """decode_logits without the device synchronisation: argmax + collapse and the copy of ids / lengths into pinned
        host buffers are enqueued on the current stream; the returned handle's `.strings()` waits for that copy only and
        builds the strings.  valid.py:40-47 decodes every batch before it touches the next one, which leaves the GPU idle
        while the host joins characters; with the handle of batch i resolved after batch i + 1 has been enqueued the
        host work runs under the next batch's kernels."""
        ids, lens, _ = ops.greedy_decode_ids(logits.float(), len(self.character), lengths)
        pool = self.__dict__.setdefault("_pinned", {})
        key = (tuple(ids.shape), ids.device.index)
        free = pool.setdefault(key, [])
        if free:
            ids_h, lens_h = free.pop()
        else:
            ids_h = torch.empty(ids.shape, dtype=torch.int32, pin_memory=True)
            lens_h = torch.empty(lens.shape, dtype=torch.int32, pin_memory=True)
        ids_h.copy_(ids, non_blocking=True)
        lens_h.copy_(lens, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(ids.device))
        return PendingStrings(self, ids_h, lens_h, ev, free)


class PendingStrings(object):
    """Result handle of CTCLabelConverter.decode_logits_async."""

    def __init__(self, converter, ids_h, lens_h, event, free):
        self._c, self._ids, self._lens, self._ev, self._free = converter, ids_h, lens_h, event, free
        self._out = None

    def done(self):
        return self._out is not None or self._ev.query()

    def strings(self):
        if self._out is None:
            self._ev.synchronize()
            self._out = self._c._host_strings(self._ids.numpy(), self._lens.tolist())
            self._free.append((self._ids, self._lens))         # the pinned pair goes back to the converter's pool
            self._ids = self._lens = None
        return self._out
