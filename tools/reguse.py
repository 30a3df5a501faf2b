"""List registers / stack / shared memory per kernel of the built library (cuobjdump --dump-resource-usage).

    python tools/reguse.py [pattern] [--lib path]
"""
import re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "mr_rl_b200", "_lib", "libmr_rl_b200.so")
args = sys.argv[1:]
if "--lib" in args:
    i = args.index("--lib"); lib = args[i + 1]; del args[i:i + 2]
pat = re.compile(args[0]) if args else None
out = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout.splitlines()
names = [l.split()[1].rstrip(":") for l in out if l.strip().startswith("Function ")]
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
it = iter(dem)
for k, l in enumerate(out):
    if l.strip().startswith("Function "):
        name = next(it)
        short = re.sub(r"\(.*", "", name).replace("void mr::", "")
        res = out[k + 1].strip()
        m = dict(x.split(":") for x in res.split() if ":" in x)
        if pat is None or pat.search(short):
            print(f"{short:70s} REG {m.get('REG'):>4s} STACK {m.get('STACK'):>5s} SHARED {m.get('SHARED'):>6s}")
