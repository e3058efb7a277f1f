"""CLIP byte-pair-encoding tokenizer (behaviour of the reference's open_clip/tokenizer.py:29-208): text -> int64 ids
[n, context_length] with <start_of_text> ... <end_of_text>, zero padded, truncated with the end token kept.

Host-side preprocessing, not part of the device hot path; written from the published algorithm:
  1. clean: (ftfy when installed) -> html.unescape twice -> strip -> collapse whitespace -> lower case
  2. split with the CLIP pattern (special tokens | English contractions | letter runs | single digits | other symbol runs)
  3. map each piece's UTF-8 bytes to printable code points (the GPT-2 byte alphabet), mark the last symbol with </w>
  4. repeatedly join the adjacent pair with the lowest merge rank until none is ranked
  5. look the resulting symbols up: ids 0..255 bytes, 256..511 bytes + </w>, then one id per merge, then the specials.

The merge table is DATA the reference ships (bpe_simple_vocab_16e6.txt.gz, 48 894 merges -> vocabulary 49 408); it is not
vendored here.  Point VITLENS_BPE_VOCAB at it (or pass bpe_path)."""
from __future__ import annotations

import gzip
import html
import os
from functools import lru_cache
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import torch

try:  # the reference cleans mojibake with ftfy; without it the text is taken as is
    import ftfy as _ftfy
except Exception:  # pragma: no cover - optional dependency
    _ftfy = None

import regex as _re

SOT, EOT = "<start_of_text>", "<end_of_text>"
N_MERGES = 49152 - 256 - 2  # lines 1 .. 48894 of the merges file (line 0 is a header)


@lru_cache()
def byte_alphabet() -> Dict[int, str]:
    """byte value -> printable code point: printable Latin-1 bytes map to themselves, the other 68 to U+0100 ..."""
    keep = [b for rng in ((0x21, 0x7E), (0xA1, 0xAC), (0xAE, 0xFF)) for b in range(rng[0], rng[1] + 1)]
    table, spare = {}, 0
    for b in keep:
        table[b] = chr(b)
    for b in range(256):
        if b not in table:
            table[b] = chr(256 + spare)
            spare += 1
    return table


def _vocab_order(alphabet: Dict[int, str]) -> List[str]:
    # the reference enumerates the alphabet in dict order: the kept bytes first, then the remapped ones
    keep = [b for rng in ((0x21, 0x7E), (0xA1, 0xAC), (0xAE, 0xFF)) for b in range(rng[0], rng[1] + 1)]
    rest = [b for b in range(256) if b not in set(keep)]
    return [alphabet[b] for b in keep + rest]


def clean_text(text: str) -> str:
    if _ftfy is not None:
        text = _ftfy.fix_text(text)
    text = html.unescape(html.unescape(text)).strip()
    return _re.sub(r"\s+", " ", text).strip().lower()


class ClipBPE:
    def __init__(self, bpe_path: str, special_tokens: Optional[Sequence[str]] = None):
        with gzip.open(bpe_path) as f:
            lines = f.read().decode("utf-8").split("\n")
        merges: List[Tuple[str, ...]] = [tuple(ln.split()) for ln in lines[1:N_MERGES + 1]]
        self.alphabet = byte_alphabet()
        self.rev_alphabet = {c: b for b, c in self.alphabet.items()}
        base = _vocab_order(self.alphabet)
        specials = [SOT, EOT] + list(special_tokens or [])
        symbols = base + [c + "</w>" for c in base] + ["".join(m) for m in merges] + specials
        self.encoder: Dict[str, int] = {s: i for i, s in enumerate(symbols)}
        self.decoder: Dict[int, str] = {i: s for s, i in self.encoder.items()}
        self.rank: Dict[Tuple[str, ...], int] = {m: i for i, m in enumerate(merges)}
        self.specials = specials
        self.pattern = _re.compile("|".join(_re.escape(s) for s in specials) +
                                   r"""|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""", _re.IGNORECASE)
        self.vocab_size = len(self.encoder)
        self.sot_id, self.eot_id = self.encoder[SOT], self.encoder[EOT]
        self.all_special_ids = [self.encoder[s] for s in specials]
        self._word = lru_cache(maxsize=65536)(self._merge_word)

    # ---- step 4: merge one piece
    def _merge_word(self, piece: str) -> Tuple[str, ...]:
        if piece in self.specials:
            return (piece,)
        parts = list(piece[:-1]) + [piece[-1] + "</w>"]
        while len(parts) > 1:
            best, where = None, -1
            for i in range(len(parts) - 1):
                r = self.rank.get((parts[i], parts[i + 1]))
                if r is not None and (best is None or r < best):
                    best, where = r, i
            if best is None:
                break
            a, b = parts[where], parts[where + 1]
            merged, i = [], 0
            while i < len(parts):  # join EVERY occurrence of the chosen pair, left to right
                if i + 1 < len(parts) and parts[i] == a and parts[i + 1] == b:
                    merged.append(a + b)
                    i += 2
                else:
                    merged.append(parts[i])
                    i += 1
            parts = merged
        return tuple(parts)

    def encode(self, text: str) -> List[int]:
        ids: List[int] = []
        for piece in self.pattern.findall(clean_text(text)):
            mapped = "".join(self.alphabet[b] for b in piece.encode("utf-8"))
            ids.extend(self.encoder[s] for s in self._word(mapped))
        return ids

    def decode(self, ids: Iterable[int]) -> str:
        text = "".join(self.decoder[int(i)] for i in ids)
        raw = bytearray()
        for ch in text.replace("</w>", " "):
            raw.append(self.rev_alphabet[ch]) if ch in self.rev_alphabet else raw.extend(ch.encode("utf-8"))
        return raw.decode("utf-8", errors="replace")

    def __call__(self, texts: Union[str, Sequence[str]], context_length: int = 77) -> torch.LongTensor:
        if isinstance(texts, str):
            texts = [texts]
        out = torch.zeros(len(texts), context_length, dtype=torch.long)
        for row, text in enumerate(texts):
            ids = [self.sot_id] + self.encode(text) + [self.eot_id]
            if len(ids) > context_length:
                ids = ids[:context_length]
                ids[-1] = self.eot_id
            out[row, : len(ids)] = torch.tensor(ids, dtype=torch.long)
        return out


SimpleTokenizer = ClipBPE  # the reference's class name (tokenizer.py:79)


def default_bpe() -> Optional[str]:
    path = os.environ.get("VITLENS_BPE_VOCAB")
    if path and os.path.exists(path):
        return path
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bpe_simple_vocab_16e6.txt.gz")
    return here if os.path.exists(here) else None


@lru_cache()
def _default_tokenizer() -> ClipBPE:
    path = default_bpe()
    if path is None:
        raise RuntimeError("tokenize(): the CLIP merge table is not vendored; set VITLENS_BPE_VOCAB to bpe_simple_vocab_16e6.txt.gz "
                           "(open_clip ships it) -- encode_text() also accepts pre-tokenised int64 ids [B, 77]")
    return ClipBPE(path)


def tokenize(texts: Union[str, Sequence[str]], context_length: int = 77) -> torch.LongTensor:
    """texts -> LongTensor [n, context_length] (reference tokenizer.py:177-208)."""
    return _default_tokenizer()(texts, context_length)


def decode(output_ids: torch.Tensor) -> str:
    return _default_tokenizer().decode(output_ids.cpu().tolist())
