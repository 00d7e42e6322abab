import sys, os, subprocess
for g, cfg in (("1","21"),("2","18")):
    for v in (16, 32, 64, 128):
        env = dict(os.environ, MSMB200_VSPAN=str(v))
        out = subprocess.run([sys.executable, "tests/gpu_perf_dev.py", g+":"+cfg], env=env, capture_output=True, text=True).stdout
        for ln in out.splitlines():
            if "method 1" in ln or "method 3" in ln: print("V=%d G%s" % (v, g), ln.strip())
