"""adapters/vulkan_interop.cpp (the Vulkan half of the frame hand-off) cannot be built against the real Vulkan
headers in this image. This test compiles it against tests/cpp/vulkan_stub — declarations of exactly the Vulkan symbols it
uses, written from the specification, nothing defined — and checks that the object binds to the C ABI entry point whose CUDA
half tests/test_external_memory_gpu.py executes. It proves well-formed, type-correct C++; it does not prove Vulkan behaviour."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


@pytest.mark.skipif(shutil.which("g++") is None or not os.path.isdir(CUDA_INC), reason="needs g++ and the CUDA runtime headers")
def test_adapter_compiles_against_declarations_and_binds_to_the_c_abi(tmp_path):
    obj = str(tmp_path / "adapter_use.o")
    cmd = ["g++", "-std=c++20", "-c", "-Wall", "-Wextra", "-Werror", "-Wno-missing-field-initializers",
           "-I", os.path.join(ROOT, "tests", "cpp", "vulkan_stub"), "-I", CUDA_INC,
           os.path.join(ROOT, "tests", "cpp", "adapter_use.cpp"), "-o", obj]
    done = subprocess.run(cmd, capture_output=True, text=True)
    assert done.returncode == 0, done.stderr
    undefined = {line.split()[-1] for line in subprocess.run(["nm", "-u", obj], capture_output=True, text=True).stdout.splitlines()}
    for symbol in ("tpdcu_bind_output_fd", "tpdcu_last_error", "cudaGetDeviceCount", "vkCreateBuffer", "vkAllocateMemory", "vkBindBufferMemory",
                   "vkGetDeviceProcAddr", "vkCmdPipelineBarrier", "vkCmdCopyBufferToImage", "vkGetPhysicalDeviceProperties2"):
        assert symbol in undefined, f"{symbol} is not referenced by the adapter"
