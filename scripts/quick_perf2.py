import sys, subprocess, os
# run quick_perf for a subset of models under each variant
for g in ("1", "2"):
    for lv in ("1", "0"):
        env = dict(os.environ, KGE_SCAN_G=g, KGE_SCAN_LOOP=lv, QP_MODELS="l2,dm,cx")
        print("== G=%s LOOP=%s" % (g, lv), flush=True)
        subprocess.run([sys.executable, "scripts/quick_perf.py", "1000000", "4096"], env=env)
