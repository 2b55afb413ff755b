"""Builds the ICD shared library, its manifest and the loader-harness (called from cpvulkan_b200/build.py)."""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False):
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    cuda_lib_dir = os.path.join(ROOT, "cpvulkan_b200", "csrc", "build")
    icd = os.path.join(bdir, "libCPVulkan_b200.so")
    harness = os.path.join(bdir, "cpvk_harness")
    headers = [os.path.join(ROOT, "include", "cpvk_vulkan.h"), os.path.join(ROOT, "include", "cpvk_cuda.h")]
    common = ["g++", "-std=c++17", "-O2", "-Wall", "-Wno-unused-function", "-fvisibility=hidden"]
    if force or _newer(icd, [os.path.join(HERE, "cpvk_icd.cpp")] + headers):
        cmd = common + ["-fPIC", "-shared", os.path.join(HERE, "cpvk_icd.cpp"), "-o", icd, "-L" + cuda_lib_dir, "-lcpvk_cuda",
                        "-Wl,-rpath,$ORIGIN/../../csrc/build", "-Wl,--no-undefined"]
        print("+", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    if force or _newer(harness, [os.path.join(HERE, "cpvk_harness.cpp")] + headers):
        cmd = common + [os.path.join(HERE, "cpvk_harness.cpp"), "-o", harness, "-ldl"]
        print("+", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    # the manifest the Vulkan loader (or the harness acting as one) finds through VK_ICD_FILENAMES (CPVulkan/CPVulkan.json:1-6)
    manifest = os.path.join(bdir, "CPVulkan_b200.json")
    with open(manifest, "w") as f:
        json.dump({"file_format_version": "1.0.0", "ICD": {"library_path": "./libCPVulkan_b200.so", "api_version": "1.1.121"}}, f, indent=2)
    return icd
