"""open_clip.tokenizer (CLIP BPE, reference tokenizer.py:29-208).  The merge table is data the reference ships and this repo
does not vendor, so: (1) the algorithm is always checked on a synthetic merge table against a direct restatement of
byte-pair encoding; (2) when the real table is reachable (VITLENS_BPE_VOCAB, or the read-only reference checkout in the
build container) the ids are checked against known CLIP ids and against the reference's own tokenizer."""
import gzip
import os
import random

import pytest
import torch

from tests.common import ROOT  # noqa: F401  (sets sys.path)

REF_VOCAB = "/root/reference/vitlens/src/open_clip/bpe_simple_vocab_16e6.txt.gz"


def _naive_bpe(word, rank):
    """Textbook BPE on one piece: join every occurrence of the best-ranked adjacent pair until none is ranked."""
    parts = list(word[:-1]) + [word[-1] + "</w>"]
    while True:
        pairs = {(parts[i], parts[i + 1]) for i in range(len(parts) - 1)}
        ranked = [p for p in pairs if p in rank]
        if not ranked:
            return parts
        a, b = min(ranked, key=rank.get)
        out, i = [], 0
        while i < len(parts):
            if i + 1 < len(parts) and (parts[i], parts[i + 1]) == (a, b):
                out.append(a + b)
                i += 2
            else:
                out.append(parts[i])
                i += 1
        parts = out


def test_bpe_algorithm_on_a_synthetic_merge_table(tmp_path):
    from open_clip.tokenizer import ClipBPE

    merges = [("t", "h"), ("th", "e</w>"), ("a", "n"), ("an", "d</w>"), ("i", "n"), ("in", "g</w>"), ("o", "o"), ("l", "l"),
              ("e", "r</w>"), ("s", "t"), ("e", "s"), ("es", "t</w>"), ("a", "t</w>"), ("c", "at</w>"), ("!", "!"), ("!!", "!</w>")]
    path = tmp_path / "merges.txt.gz"
    with gzip.open(path, "wt", encoding="utf-8") as f:
        f.write("#version: synthetic\n" + "\n".join(" ".join(m) for m in merges) + "\n")
    tok = ClipBPE(str(path))
    rank = {m: i for i, m in enumerate(merges)}
    assert tok.vocab_size == 512 + len(merges) + 1 + 2  # (the trailing empty line is one more "merge", as in the reference)
    rng = random.Random(0)
    words = ["the", "and", "singing", "tooling", "cat", "attest", "better", "a", "!!!", "hello1world", "naïve", "日本語"]
    words += ["".join(rng.choice("theandigolrsc!1 ") for _ in range(rng.randint(1, 12))) for _ in range(200)]
    for text in words:
        ids = tok.encode(text)
        want = []
        for piece in tok.pattern.findall(" ".join(text.split()).lower().strip()):
            mapped = "".join(tok.alphabet[b] for b in piece.encode("utf-8"))
            want += [tok.encoder[s] for s in _naive_bpe(mapped, rank)]
        assert ids == want, text
        assert tok.decode(ids).replace(" ", "") == "".join(text.split()).lower()
    out = tok(["the cat", "and " * 100], context_length=16)
    assert out.shape == (2, 16) and out.dtype == torch.long
    assert out[0, 0] == tok.sot_id and tok.eot_id in out[0].tolist() and out[0, -1] == 0
    assert out[1, 0] == tok.sot_id and out[1, -1] == tok.eot_id  # truncated: the end token is kept


def _real_vocab():
    for p in (os.environ.get("VITLENS_BPE_VOCAB"), REF_VOCAB):
        if p and os.path.exists(p):
            return p
    return None


@pytest.mark.skipif(_real_vocab() is None, reason="CLIP merge table not available (not vendored)")
def test_real_vocabulary_known_ids_and_reference_tokenizer():
    from open_clip.tokenizer import ClipBPE

    tok = ClipBPE(_real_vocab())
    assert tok.vocab_size == 49408 and tok.sot_id == 49406 and tok.eot_id == 49407
    assert tok("a photo of a cat")[0, :7].tolist() == [49406, 320, 1125, 539, 320, 2368, 49407]
    # ids the reference's own tokenizer produced for these texts (oracle/make_tokenizer_golden.py; ftfy absent on both sides)
    import json

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tokenizer_ids.json")) as f:
        gold = json.load(f)
    assert tok(gold["texts"]).tolist() == gold["ids"]
    assert tok(gold["texts"], context_length=8).tolist() == gold["ids_ctx8"]
