#!/bin/bash
# compute-sanitizer over the forward path on small inputs: memcheck, racecheck and synccheck on one GPU,
# memcheck on the two-process fused all-reduce.  Logs -> gpurun_out/ (copied to profiles/ once read).
#   gpurun --gpus 2 --timeout 1500 -- bash tools/sanitize.sh
out=${1:-gpurun_out}
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py > $out/r02_sanitizer_$tool.log 2>&1
  echo "exit code $?" >> $out/r02_sanitizer_$tool.log
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  rm -f /tmp/imc_sanitize_id
  timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_case.py --ranks 2 --rank 1 > $out/r02_sanitizer_memcheck_2gpu_rank1.log 2>&1 &
  timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_case.py --ranks 2 --rank 0 > $out/r02_sanitizer_memcheck_2gpu_rank0.log 2>&1
  echo "exit code $?" >> $out/r02_sanitizer_memcheck_2gpu_rank0.log
  wait
fi
tail -n 4 $out/r02_sanitizer_*.log
