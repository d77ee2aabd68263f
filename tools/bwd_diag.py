"""Spread of the lambda-gradient error over seeds / layer subsets (development diagnostic)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_checks as G

for layers, B, wseed in [((0, 1, 2, 3), 3, 99), ((2, 3), 2, 1234), ((2, 3), 3, 99), ((3,), 4, 5), ((0, 1, 2, 3), 4, 5), ((1,), 2, 7)]:
    print(layers, B, wseed, G.check_embedding_grad_vs_oracle("tiny", B, layers, 0, wseed), flush=True)
