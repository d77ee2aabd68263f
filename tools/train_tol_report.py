"""Print the measured lambda-gradient errors behind the tolerances of tests/test_gpu_train.py (run on the GPU box)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_checks as G  # noqa: E402

out = {"golden": G.check_training_step_vs_golden("htsat_tiny_b2.npz")}
for layers, B, wseed in [((0, 1, 2, 3), 3, 99), ((2, 3), 2, 1234), ((1,), 2, 7)]:
    out[f"embedding_loss_{layers}_{B}_{wseed}"] = G.check_embedding_grad_vs_oracle("tiny", B, layers, 0, wseed)
out["zero_shot_subset"] = G.check_training_step_vs_oracle("tiny", 2, (1,), cosine=True)
print(json.dumps({k: {kk: (round(vv, 6) if isinstance(vv, float) else vv) for kk, vv in v.items()} for k, v in out.items()}, indent=1))
