"""Writes tests/golden/exact_random40.json: exact posterior expectations (oracle/dense_oracle.py,
160 compound states; ~4 minutes of scipy expm_frechet) for tests/test_gpu_random_model.py."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from test_gpu_random_model import RandomModel, RandomData  # noqa: E402
from test_gpu_toy_exact import _flatten_exact  # noqa: E402
from oracle.dense_oracle import DensePosterior  # noqa: E402
from nxblink_b200.packing import PackedModel  # noqa: E402

M = RandomModel()
D = RandomData(M)
dp = DensePosterior(M, D)
flat = _flatten_exact(dp.expected_summary(), PackedModel(M))
with open(os.path.join(HERE, 'exact_random40.json'), 'w') as f:
    json.dump(dict(likelihood=dp.likelihood,
        values=[[repr(k), float(v)] for k, v in sorted(flat.items(), key=lambda kv: repr(kv[0]))]),
        f, indent=0)
print('wrote', len(flat), 'values')
