"""Kernel-experiment helper: build libroboy_b200.so with extra nvcc flags into build/variants/<name>/ (git-ignored; it
travels to the GPU box) -- load it with ROBOY_B200_LIB=build/variants/<name>/libroboy_b200.so.
usage: python tools/build_variant.py <name> [-DMACRO=value ...]"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_roboy_b200 import build as B
name, extra = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(os.path.dirname(B._PKG), "build", "variants", name)
os.makedirs(out_dir, exist_ok=True)
flags = [f for f in B.NVCC_FLAGS if f != "-shared"] + extra


def one(src):
    obj = os.path.join(out_dir, os.path.splitext(src)[0] + ".o")
    subprocess.run([B._nvcc()] + flags + ["-c", "-o", obj, os.path.join(B.CSRC, src)], check=True)
    return obj


with ThreadPoolExecutor(max_workers=len(B.SOURCES)) as pool:
    objs = list(pool.map(one, B.SOURCES))
lib = os.path.join(out_dir, "libroboy_b200.so")
subprocess.run([B._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", lib] + objs, check=True)
for o in objs:
    os.remove(o)
print(lib)
