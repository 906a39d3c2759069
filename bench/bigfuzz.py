import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import test_gpu_fuzz as t
bad = 0
for seed in range(200, 500):
    try: t.run_trials(seed, 12, 400, 40, 30)
    except AssertionError as e: bad += 1; print("FAIL small", seed, str(e)[:200], flush=True)
for seed in range(500, 560):
    try: t.run_trials(seed, 3, 30_000, 300, 250)
    except AssertionError as e: bad += 1; print("FAIL medium", seed, str(e)[:200], flush=True)
for seed in range(600, 606):
    try: t.run_trials(seed, 2, 300_000, 1200, 900)
    except AssertionError as e: bad += 1; print("FAIL large", seed, str(e)[:200], flush=True)
print("bigfuzz done, failures:", bad)
