"""Build liborgym_b200.so several times with different -D knobs (kernel-tuning experiments) into
or-gym-inventory_b200/csrc/variants/<name>.so; select one at run time with ORGYM_B200_LIB=<path>.
usage: python tools/build_variants.py name1="-DX=1 -DY=2" name2="..."  (the default build is restored at the end)"""
import os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "or-gym-inventory_b200", "csrc")
os.makedirs(os.path.join(CSRC, "variants"), exist_ok=True)
for spec in sys.argv[1:]:
    name, flags = spec.split("=", 1)
    env = dict(os.environ, ORGYM_NVCC_EXTRA=flags)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "__graft_entry__.py"), "--force"], env=env, stdout=subprocess.DEVNULL)
    shutil.copy(os.path.join(CSRC, "liborgym_b200.so"), os.path.join(CSRC, "variants", name + ".so"))
    print("built", name, flags)
subprocess.check_call([sys.executable, os.path.join(ROOT, "__graft_entry__.py"), "--force"], stdout=subprocess.DEVNULL)
