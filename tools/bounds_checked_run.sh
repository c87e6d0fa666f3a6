#!/bin/bash
# The memory-safety evidence that can be produced on this pool (compute-sanitizer is closed by the operators, see the first
# lines of the log): the library rebuilt with -DIMC_DEBUG_BOUNDS (device-side asserts on every data-dependent index) runs
# every launch shape on small inputs, the randomised cross-check against the CPU oracle, and the two-process fused all-reduce.
#   IMC_LIB_PATH=$PWD/build_dbg/libimc_dbg.so ... ; gpurun --gpus 2 -- bash tools/bounds_checked_run.sh
export IMC_LIB_PATH=$PWD/build_dbg/libimc_dbg.so
if [ ! -f "$IMC_LIB_PATH" ]; then      # ~2 minutes of nvcc
  mkdir -p build_dbg
  IMC_EXTRA_NVCC_FLAGS="-DIMC_DEBUG_BOUNDS" python -c "import importlib.util as u; s=u.spec_from_file_location('b','imcoalhmm_b200/build.py'); m=u.module_from_spec(s); s.loader.exec_module(m); m.build(force=True)"
fi
echo "== compute-sanitizer on this pool:"; compute-sanitizer --tool memcheck python -c "print(1)" 2>&1 | head -3
echo "== sanitize_case.py, one GPU (IMC_DEBUG_BOUNDS build)"; timeout 600 python tools/sanitize_case.py 2>&1 | tail -3
echo "== fuzz_zip.py 150 trials (IMC_DEBUG_BOUNDS build)"; timeout 900 python tools/fuzz_zip.py 150 77 2>&1 | tail -2
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  echo "== sanitize_case.py, two processes / two GPUs, fused peer-to-peer all-reduce"
  rm -f /tmp/imc_sanitize_id
  timeout 600 python tools/sanitize_case.py --ranks 2 --rank 1 2>&1 | tail -2 &
  timeout 600 python tools/sanitize_case.py --ranks 2 --rank 0 2>&1 | tail -2
  wait
fi
