#!/usr/bin/env python
"""Golden CER / WER values from the UNMODIFIED reference (models/evaluate.py:94-134).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_cer_golden.py [--ref /root/reference]

Seeded random pairs over a small CJK + Latin alphabet (edits of a common source string, plus unrelated,
empty and identical pairs) -> tests/golden/cer_wer_vectors.json.  Only strings and numbers are stored."""
import argparse
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("FDDM_REF", "/root/reference"))
    args = ap.parse_args()
    sys.dont_write_bytecode = True
    sys.path.insert(0, args.ref)
    from models.evaluate import calculate_cer, calculate_wer
    rng = random.Random(1337)
    alphabet = list("今天氣很好我們去學校吃飯了嗎台灣語音辨識模型abcdefgh ")

    def mutate(s):
        s = list(s)
        for _ in range(rng.randint(0, 6)):
            op = rng.choice("ids")
            pos = rng.randint(0, max(0, len(s) - 1))
            if op == "i":
                s.insert(pos, rng.choice(alphabet))
            elif op == "d" and s:
                del s[pos]
            elif s:
                s[pos] = rng.choice(alphabet)
        return "".join(s)

    pairs = [("", ""), ("", "abc"), ("abc", ""), ("今天天氣好", "今天氣很好"), ("a b c", "a c d e"), ("same same", "same same"),
             ("  leading space", "leading  space "), ("x", "y")]
    for _ in range(56):
        src = "".join(rng.choice(alphabet) for _ in range(rng.randint(1, 60)))
        pairs.append((src, mutate(src)) if rng.random() < 0.8 else
                     (src, "".join(rng.choice(alphabet) for _ in range(rng.randint(0, 60)))))
    out = [{"ref": r, "hyp": h, "cer": calculate_cer(r, h), "wer": calculate_wer(r, h)} for r, h in pairs]
    with open(os.path.join(HERE, "cer_wer_vectors.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False, indent=0)
    print(f"wrote {len(out)} pairs")


if __name__ == "__main__":
    main()
