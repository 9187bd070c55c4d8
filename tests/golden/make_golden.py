"""Writes tests/golden/kat.json: the two known-answer vectors the reference holds for this path,
transcribed from the reference sources (nothing is computed here except KAT2's key packing).

  KAT1  /root/reference/cuda_test.py:19-34   inputs :19-21,:27 ; expected gradient in the string at :34
        ("(0.44,0.08,0.74,0.08,0.2)となるはず" = "should become ..."); forward y follows from :23.
  KAT2  /root/reference/uitility.py:383-393  commented doc example of grouped_cumprod(A, G):
        groups are NOT adjacent there: semantics = stable sort by group key -> scan -> unsort.
Run:  python tests/golden/make_golden.py
"""
import json
import os

kat1 = {
    "source": "cuda_test.py:19-34",
    "x": [0.4, 0.2, 0.1, 0.8, 0.2],
    "grad_out": [0.4, 0.2, 0.1, 0.8, 0.2],          # grad = torch.clone(param), cuda_test.py:20
    "key": [0, 0, 1, 1, 2],                           # cuda_test.py:21
    "seg_end": [2, 4, 5],                             # cuda_test.py:27
    "y": [0.4, 0.08, 0.1, 0.08, 0.2],
    "grad_in": [0.44, 0.08, 0.74, 0.08, 0.2],         # cuda_test.py:34
}
kat2 = {
    "source": "uitility.py:383-393",
    "A": [1, 2, 3, 4, 5, 6, 7],
    "G": [[1, 1], [1, 2], [1, 1], [1, 2], [1, 3], [1, 1], [1, 3]],
    "expected": [1, 2, 3, 8, 5, 18, 35],
}
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.json")
with open(out, "w") as f:
    json.dump({"kat1": kat1, "kat2": kat2}, f, indent=1)
print("wrote", out)
