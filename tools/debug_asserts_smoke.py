"""Run every kernel path on the debug build (index / invariant asserts compiled in: -DB2M_DEBUG_ASSERTS).

    python -m mlx_mcmc_b200.build --debug          # here, no GPU needed (nvcc cross-compiles); the .so travels with gpurun
    B2M_LIB=mlx_mcmc_b200/csrc/_debug/libb200mcmc.so python tools/debug_asserts_smoke.py       # on the GPU box

compute-sanitizer is closed on the GPU pool (gpurun_out/memcheck.log, round 1); this is the replacement the pool's own
message asks for.  A failed assert traps the kernel: the next CUDA call raises and the script exits non-zero."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.environ.get("B2M_LIB", "")
assert "_debug" in lib and os.path.exists(lib), "set B2M_LIB to the debug build (python -m mlx_mcmc_b200.build --debug)"
sys.argv = [os.path.join(ROOT, "tools", "sanitize_smoke.py")]
runpy.run_path(sys.argv[0], run_name="__main__")
print("debug_asserts_smoke: all kernels ran with B2M_DEBUG_ASSERTS on, no assert fired")
