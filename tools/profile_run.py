"""Short run for ncu: B replicas of config3, a few MD rounds (launch list / full capture target)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from parity import CONFIGS, test_positions, ue
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = CONFIGS[int(sys.argv[3]) if len(sys.argv) > 3 else 3]
pos = test_positions(cfg, 5)[1:]
be = ue.BatchEngine(cfg, B)
be.set_pos(np.tile(pos, (B // 4 + 1, 1, 1))[:B])
be.md_init(0.8, seed=3)
be.md_run(rounds)
print('ok', be.evaluate(want_deriv=False)[:3])
