"""CPU check of the staged-epilogue index arithmetic of k_gemm2 / k_gemm2h: the Python scripts under
tests/emulation mirror the CUDA code line by line (staging positions, swizzle, C-order enumeration, both
store paths) and assert that every tile element lands exactly once at its C address."""
import os
import runpy

HERE = os.path.dirname(os.path.abspath(__file__))


def test_gemm2_epilogue_index_math():
    runpy.run_path(os.path.join(HERE, "emulation", "emul_gemm2_epilogue.py"))


def test_gemm2h_epilogue_index_math():
    runpy.run_path(os.path.join(HERE, "emulation", "emul_gemm2h_epilogue.py"))
