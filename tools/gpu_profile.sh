#!/bin/bash
# One gpurun call: clocks + library reference point for SGEMM, one `ncu --set full` capture per hot kernel (each only
# after the identical plain command exited 0), and the launch list of bench.py.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
P=python
run_ncu() { # name kernel-regex run_one-args...
  name=$1; k=$2; shift 2
  $P tools/run_one.py "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$name \
      $P tools/run_one.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"; cat gpurun_out/plain_$name.log
}
$P tools/run_one.py sgemm 8192 0 20
$P tools/run_one.py sgemm 8192 1 20
$P tools/run_one.py cublas_sgemm 8192 0 20
$P tools/run_one.py tcgemm 8192 0 50
$P tools/run_one.py tf32gemm 8192 0 20
run_ncu sgemm sgemm_tma sgemm 8192 0 3
run_ncu transpose transpose64_f32 transpose 16384 0 3
B200_TUNE=transpose.impl=1 run_ncu transpose_tma transpose64_tma transpose 16384 0 3
run_ncu hist hist_atomic hist 1073741824 0 3
run_ncu hist_fused hist_allreduce hist_fused 134217728 0 3
run_ncu copy copy_bulk copy 67108864 0 3
run_ncu copy_twin_vec copy_twin_vec copy_twin 67108864 0 3
run_ncu copy_twin_scalar copy_twin_kernel copy_twin 67108864 1 3
run_ncu tcgemm gemm_tcgen05 tcgemm 8192 0 3
run_ncu tf32gemm gemm_tcgen05 tf32gemm 8192 0 3
run_ncu tf32x3gemm gemm_tcgen05 tf32x3gemm 8192 0 3
run_ncu saxpy saxpy saxpy 67108864 0 3
# launch list: few steps per entry so that EVERY kernel of the suite (SGEMM and the tcgen05 GEMMs come last) fits under the cap
$P bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    $P bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
echo "launch list rc=$?"; tail -c 600 gpurun_out/bench_plain.log
