"""tests/golden/tokenizer_ids.json from the REFERENCE's tokenizer (open_clip/tokenizer.py:177-208).  TEST INFRASTRUCTURE;
runs only in the build container (needs /root/reference):   python oracle/make_tokenizer_golden.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import cases as C  # noqa: E402
from oracle import ref_import  # noqa: E402

TEXTS = ["a photo of a cat", "A photo of a dog.", "  Hello,   WORLD!! it's 2024 &amp; counting ", "naïve café — 日本語のテキスト", "x" * 400,
         "point cloud of an airplane", "the sound of rain on a tin roof", "", "<start_of_text> nested specials <end_of_text>",
         "they're, we've; I'm: she'll? he'd! don't", "3.14159 and 42 and 1,000,000", "emoji 🙂 and symbols ©®™ ±×÷", "tab\tnewline\nreturn\r end"]

if __name__ == "__main__":
    assert ref_import.available(), "needs /root/reference"
    open_clip, _, _ = ref_import.import_reference()
    ids = open_clip.tokenizer.tokenize(TEXTS)
    short = open_clip.tokenizer.tokenize(TEXTS, context_length=8)
    out = {"texts": TEXTS, "ids": ids.tolist(), "ids_ctx8": short.tolist()}
    path = os.path.join(C.GOLDEN_DIR, "tokenizer_ids.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print(path, os.path.getsize(path), "bytes;", len(TEXTS), "texts")
