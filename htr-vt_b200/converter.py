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
        ids = ids.cpu().tolist()
        lens = lens.cpu().tolist()
        table = self.character
        return [''.join(table[i] for i in row[:n]) for row, n in zip(ids, lens)]

    def decode(self, text_index, length):
        if not text_index.is_cuda:
            text_index = text_index.to(self.device)
        ids, lens = ops.ctc_collapse(text_index, length, len(self.character))
        return self._to_strings(ids, lens)

    def decode_logits(self, logits, lengths=None):
        ids, lens, _ = ops.greedy_decode_ids(logits.float(), len(self.character), lengths)
        return self._to_strings(ids, lens)
