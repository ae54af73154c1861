"""Build libballenv_b200.so in-tree with nvcc for sm_100a (no torch headers involved).

    python gym_ballenv_b200/build.py [--force]

The kernel instantiations (precision x window) are separate translation units compiled in
parallel; the objects land in gym_ballenv_b200/csrc/_obj/ and the library next to this file
(git-ignored, but it travels to the GPU box with the gpurun snapshot).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj" + os.environ.get("BALLENV_OBJ_SUFFIX", ""))
LIB = os.path.join(HERE, os.environ.get("BALLENV_LIB_NAME", "libballenv_b200.so"))   # A/B experiments: other name
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
if os.environ.get("BALLENV_EXTRA_DEFS"):   # A/B experiments: e.g. "-DBALLENV_EXP_FOO=1"
    FLAGS.extend(os.environ["BALLENV_EXTRA_DEFS"].split())
if os.environ.get("BALLENV_MINBLOCKS"):   # tuning experiments: resident blocks per SM the kernels are compiled for
    FLAGS.append("-DBALLENV_MINBLOCKS=" + os.environ["BALLENV_MINBLOCKS"])

# (positions type, window (0 = any), fast specialisation, launcher name)
INSTANCES = [("float", 5, 0, "launch_f32_w5"), ("float", 10, 0, "launch_f32_w10"), ("float", 0, 0, "launch_f32_wany"),
             ("double", 5, 0, "launch_f64_w5"), ("double", 10, 0, "launch_f64_w10"), ("double", 0, 0, "launch_f64_wany"),
             ("float", 5, 1, "launch_f32_w5_fast"), ("float", 10, 1, "launch_f32_w10_fast"),
             ("float", 0, 1, "launch_f32_wany_fast")]


# thread-per-environment kernels (ballenv_lean.cuh): (window, static obstacles, dynamic obstacles)
LEAN_INSTANCES = [(5, 13, 5), (10, 13, 5), (10, 8, 24), (5, 8, 24)]
LEAN_RT_WINDOWS = [5, 10]


def _sources():
    deps = [os.path.join(CSRC, f) for f in ("ballenv_kernels.cuh", "ballenv_rng.cuh", "ballenv_lean.cuh", "ballenv_lean_rt.cuh",
                                            "ballenv_features.cuh", "ballenv_patches.cuh", "ballenv_reset_fixed.cuh", "ballenv_a2c.cuh")]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "ballenv.h"))
    jobs = [(os.path.join(CSRC, "ballenv_capi.cu"), os.path.join(OBJ, "ballenv_capi.o"), [])]
    for t, w, fast, name in INSTANCES:
        jobs.append((os.path.join(CSRC, "ballenv_inst.cu"), os.path.join(OBJ, name + ".o"),
                     ["-DBALLENV_T=" + t, "-DBALLENV_W=%d" % w, "-DBALLENV_FAST=%d" % fast, "-DBALLENV_NAME=" + name]))
    for w, ks, kd in LEAN_INSTANCES:
        for g in (1, 2):   # lanes per environment
            name = "launch_lean_w%d_s%d_d%d_g%d" % (w, ks, kd, g)
            defs = ["-DBALLENV_W=%d" % w, "-DBALLENV_KS=%d" % ks, "-DBALLENV_KD=%d" % kd, "-DBALLENV_G=%d" % g,
                    "-DBALLENV_NAME=" + name]
            # the policy-in-the-loop rollout (ballenv_rollout_policy) of the same configuration
            defs.append("-DBALLENV_POLICY_NAME=" + name.replace("launch_lean_", "launch_lean_policy_"))
            jobs.append((os.path.join(CSRC, "ballenv_lean_inst.cu"), os.path.join(OBJ, name + ".o"), defs))
    for w in LEAN_RT_WINDOWS:   # the same kernels with the obstacle counts read from the configuration at run time
        for g in (1, 2):
            name = "launch_lean_w%d_rt_g%d" % (w, g)
            jobs.append((os.path.join(CSRC, "ballenv_lean_rt_inst.cu"), os.path.join(OBJ, name + ".o"),
                         ["-DBALLENV_W=%d" % w, "-DBALLENV_KS=-1", "-DBALLENV_KD=-1", "-DBALLENV_G=%d" % g,
                          "-DBALLENV_NAME=" + name, "-DBALLENV_SMEM_NAME=lean_rt_smem_w%d_g%d" % (w, g),
                          "-DBALLENV_POLICY_NAME=" + name.replace("launch_lean_", "launch_lean_policy_")]))
    return deps, jobs


def _stale(target, srcs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def _compile(job, deps, force, log):
    src, obj, defs = job
    if not force and not _stale(obj, [src] + deps):
        return obj
    cmd = [NVCC] + ARCH + FLAGS + defs + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append((os.path.basename(obj), r.stderr))
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (obj, " ".join(cmd), r.stderr))
    with open(obj + ".ptxas", "w") as f:   # `-Xptxas -v` of this object: registers, spills, shared memory per kernel
        f.write(r.stderr)
    return obj


def _kernel_info(objs):
    """{demangled-ish kernel key: {registers, spill_store_bytes, spill_load_bytes, stack_bytes, smem_bytes}} from the
    ptxas logs kept beside the objects (bench.py quotes the step kernels' figures in its roofline object)."""
    import re
    info = {}
    for obj in objs:
        try:
            text = open(obj + ".ptxas").read()
        except OSError:
            continue
        cur = None
        for line in text.splitlines():
            m = re.search(r"Compiling entry function '([^']+)'", line)
            if m:
                cur = m.group(1)
                info[cur] = {"object": os.path.basename(obj)}
                continue
            if cur is None:
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and "stack_bytes" not in info[cur]:   # the entry function's own line (those of its callees follow)
                info[cur].update(stack_bytes=int(m.group(1)), spill_store_bytes=int(m.group(2)), spill_load_bytes=int(m.group(3)))
            m = re.search(r"Used (\d+) registers", line)
            if m:
                info[cur]["registers"] = int(m.group(1))
                ms = re.search(r"(\d+) bytes smem", line)
                info[cur]["smem_bytes"] = int(ms.group(1)) if ms else 0
    return info


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    deps, jobs = _sources()
    log = []
    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        objs = list(ex.map(lambda j: _compile(j, deps, force, log), jobs))
    if force or _stale(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (" ".join(cmd), r.stderr))
    import json
    with open(os.path.join(HERE, "build_info.json"), "w") as f:
        json.dump(_kernel_info(objs), f, indent=0, sort_keys=True)
    if verbose:
        for name, err in log:
            for line in err.splitlines():
                if "Used" in line or "spill" in line or "error" in line or "warning" in line:
                    print("[%s] %s" % (name, line.strip()))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
