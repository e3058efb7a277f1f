"""CLIP BPE tokenizer entry points (reference open_clip/tokenizer.py:177-208).  The BPE merges file is data the
reference ships (bpe_simple_vocab_16e6.txt.gz) and is not vendored here; point VITLENS_BPE_VOCAB at it.  The hot
path consumes int64 token ids [B, 77] directly (SURVEY.md 8(f).4 lists the tokenizer as a later row)."""
import os

import torch


def tokenize(texts, context_length: int = 77) -> torch.LongTensor:
    path = os.environ.get("VITLENS_BPE_VOCAB")
    if not path or not os.path.exists(path):
        raise RuntimeError("tokenize(): set VITLENS_BPE_VOCAB to the CLIP bpe_simple_vocab_16e6.txt.gz file; "
                           "encode_text() accepts pre-tokenised int64 ids [B, 77]")
    raise NotImplementedError("BPE tokenisation is a later coverage row (SURVEY.md 8(f).4)")
