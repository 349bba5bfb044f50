#!/bin/bash
# GPU-side experiment driver: assembly / SpMV kernel variants on the cube (timings only; parity is pytest's job)
mkdir -p gpurun_out
N=${CUBE_N:-64}
run() {  # name, env...
  name=$1; shift
  env "$@" python tools/prof_cube.py --n $N --reps 20 > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err
  python - <<PY
import json
try:
    o = json.load(open("gpurun_out/var_$name.json"))
    print("%-12s asm %.4f ms %6.0f Mtets/s | spmv %.4f ms %5.0f GB/s | cocg_it %.4f ms %5.0f GB/s | bicg_it %.4f ms %5.0f GB/s" % (
        "$name", o["assembly"]["ms"], o["assembly"]["mtets_per_s"], o["spmv"]["ms"], o["spmv"]["gbs"],
        o["cocg_jacobi_it"]["ms"], o["cocg_jacobi_it"]["gbs"], o["bicgstab_jacobi_it"]["ms"], o["bicgstab_jacobi_it"]["gbs"]))
except Exception as e:
    print("$name failed", e, open("gpurun_out/var_$name.err").read()[-800:])
PY
}
run default A=1
run sched_cplx EDGEFEM_B200_ASM_NO_REAL=1
run sched_3cta EDGEFEM_B200_ASM_CTAS=3
