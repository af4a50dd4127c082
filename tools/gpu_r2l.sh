#!/bin/bash
# Round-2 session L: accuracy/time scan of the primitive-quartet cutoff (tools/prim_cut_scan.py), A/B of the L1 prefetch of the
# digestion's density lines (-DCF_PREFETCH_D) and of the one-ahead ket primitive loads (-DCF_PREFETCH_KET) in the warp-group kernels.
TAG=${TAG:-r2l}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -q -x -k "parity or blocks" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
run() {  # name lib workload
  CHINIUM_FOCK_LIB=$PWD/chinium_b200/$2 timeout 600 python bench.py --workload $3 --steps 5 --warmup 3 --per-class --no-cpu-baseline > gpurun_out/${TAG}_$1_$3.json 2> gpurun_out/${TAG}_$1_$3.err
  echo "bench $1 $3 rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_$1_$3.json 3
}
for w in c18 fe4s4 h2o64; do
  run new libchinium_fock.so $w
  run pfd libchinium_fock_pfd.so $w
  run pfk libchinium_fock_pfk.so $w
done
timeout 1500 python tools/prim_cut_scan.py --workloads c18 fe4s4 h2o64 --cuts 1e-22 1e-20 1e-18 1e-16 1e-14 --out gpurun_out/${TAG}_primcut.json 2>&1 | tee gpurun_out/${TAG}_primcut.txt
# one `ncu --set full` capture of the heaviest warp-group kernel (dp|ds, MK 3) and the heaviest thread-per-quartet kernel (dp|ps) of a c18 build
bash tools/ncu_capture.sh ${TAG}_c18_top c18 "eri_jk_wgILi2ELi1ELi2ELi0ELi3E|eri_jk_tpqILi2ELi1ELi1ELi0E" 2
timeout 600 python tools/time_hess.py h2o bo3h3 fe4s4 c18 > gpurun_out/${TAG}_hess_timing.jsonl 2> gpurun_out/${TAG}_hess_timing.err; echo "hess timing rc=$?"; cat gpurun_out/${TAG}_hess_timing.jsonl
