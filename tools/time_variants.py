"""Time the device-resident step (24 steps = one diurnal cycle) of several builds of the library (tools/build_variant.sh) on one GPU.
usage: python tools/time_variants.py NI NJ name1 name2 ...   (name 'main' = libnoahmp_b200.so)"""
import json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ni, nj = sys.argv[1:3]
res = {}
for name in sys.argv[3:]:
    env = dict(os.environ)
    if name != "main":
        env["NOAHMP_B200_LIB"] = os.path.join(root, "noahmp_b200", f"libnoahmp_b200_{name}.so")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--grid", ni, nj, "--steps", "24", "--warmup",
                          "3", "--no-e2e", "--no-cpu-baseline"], env=env, capture_output=True, text=True)
    try:
        line = json.loads(out.stdout.strip().splitlines()[-1])
        res[name] = (line["ms_per_step"], line["value"])
        print(f"{name:16s} {line['ms_per_step']:9.3f} ms/step  {line['value']/1e6:9.1f} M col-steps/s", flush=True)
    except Exception:
        print(name, "FAILED", out.stdout[-300:], out.stderr[-600:], flush=True)
