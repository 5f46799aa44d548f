"""Build the library for the CPU emulator (TEST INFRASTRUCTURE, see cuda_emu.h): the product sources csrc/*.cu / *.cuh are
copied to a scratch directory with two purely syntactic rewrites g++ needs,

    kernel<<<grid, block, smem, stream>>>(args...)      ->  launch_pdl(false, kernel, grid, block, smem, stream, args...)
    extern __shared__ T name[];                          ->  T* name = reinterpret_cast<T*>(vch_emu::dynamic_smem());
    cudaFuncSetAttribute(kernel, ...)                    ->  vch_emu::func_set_attribute(kernel, ...)   (a no-op)

compiled as C++ with -DVCH_CPU_EMU -include cuda_emu.h and linked with cudart_shim.cpp instead of libcudart.  The result
exports the same C ABI (include/vch_b200.h); tests load it through the unmodified ctypes binding with VCH_NO_GRAPHS=1.

    python tests/emu/build_emu_lib.py <out_dir>   ->  <out_dir>/libvch_b200_emu.so
"""
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200")
CSRC = os.path.join(PKG, "csrc")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _split_top_level(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite(src: str) -> str:
    # kernel<<<cfg>>>(  ->  launch_pdl(false, kernel, cfg...,
    res, pos = [], 0
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_]*)\s*<<<", src):
        end = src.index(">>>", m.end())
        cfg = _split_top_level(src[m.end():end])
        assert len(cfg) == 4, f"launch configuration with {len(cfg)} entries: {cfg}"
        after = src[end + 3:]
        assert after.lstrip().startswith("("), "kernel launch without an argument list"
        paren = end + 3 + after.index("(")
        res.append(src[pos:m.start()])
        res.append(f"launch_pdl(false, {m.group(1)}, dim3({cfg[0]}), dim3({cfg[1]}), (size_t)({cfg[2]}), {cfg[3]}, ")
        pos = paren + 1
    res.append(src[pos:])
    src = "".join(res)
    # kernel attributes are meaningless on the CPU (and the runtime's template overload exists only under nvcc)
    src = src.replace("cudaFuncSetAttribute(", "vch_emu::func_set_attribute(")
    # dynamic shared memory
    src = re.sub(r"extern\s+__shared__\s+([A-Za-z0-9_]+)\s+([A-Za-z0-9_]+)\[\];",
                 r"\1* \2 = reinterpret_cast<\1*>(vch_emu::dynamic_smem());", src)
    return src


def build(out_dir: str, extra_flags=()) -> str:
    os.makedirs(out_dir, exist_ok=True)
    work = os.path.join(out_dir, "src", "pkg", "csrc")         # keeps the relative include of ../../include/vch_b200.h valid
    os.makedirs(work, exist_ok=True)
    inc = os.path.join(out_dir, "src", "include")
    os.makedirs(inc, exist_ok=True)
    shutil.copy(os.path.join(ROOT, "include", "vch_b200.h"), inc)
    for f in sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))):
        with open(os.path.join(CSRC, f)) as fh:
            text = fh.read()
        with open(os.path.join(work, f), "w") as fh:
            fh.write(rewrite(text))
    lib = os.path.join(out_dir, "libvch_b200_emu.so")
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-DVCH_CPU_EMU", "-I" + os.path.join(CUDA_HOME, "include"),
           "-include", os.path.join(HERE, "cuda_emu.h")] + list(extra_flags) + os.environ.get("VCH_EMU_EXTRA_FLAGS", "").split() + [
           "-x", "c++", os.path.join(work, "vch2d.cu"), os.path.join(work, "vch1d.cu"),
           os.path.join(HERE, "cudart_shim.cpp"), "-o", lib]
    subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    print(build(sys.argv[1] if len(sys.argv) > 1 else "/tmp/vch_emu_build"))
