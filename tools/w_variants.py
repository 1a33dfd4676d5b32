"""Times LogLikelihoodWeight kernel variants (BE_W_VARIANT) at the hbm_stages size (developer tool).
usage: w_variants.py v0 v1 ...   (each in its own process: the variant is read once per process)"""
import os, subprocess, sys
for v in sys.argv[1:]:
    env = dict(os.environ, BE_W_VARIANT=v)
    out = subprocess.run([sys.executable, "tools/prof_weights.py", "time"], env=env, capture_output=True, text=True)
    print(f"v{v}", out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)
