#!/bin/bash
# Round-2 session O: bra-loop vs one-bra-pair-per-item thread-per-quartet kernels after the warp-sum change (developer override
# CF_BRALOOP, no rebuild), and `ncu --set full` of the heaviest (H2O)64 bra-loop kernel and the c18 dp|ps kernel as they are now.
TAG=${TAG:-r2o}
mkdir -p gpurun_out
for w in c18 fe4s4; do
  for b in default 0 1; do
    if [ $b = default ]; then unset CF_BRALOOP; else export CF_BRALOOP=$b; fi
    timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --per-class --no-cpu-baseline > gpurun_out/${TAG}_braloop${b}_$w.json 2> gpurun_out/${TAG}_braloop${b}_$w.err
    echo "bench braloop=$b $w rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_braloop${b}_$w.json 2
  done
done
unset CF_BRALOOP
bash tools/ncu_capture.sh ${TAG}_c18_tpq c18 "eri_jk_tpqILi2ELi1ELi1ELi0E|eri_jk_tpqILi2ELi0ELi1ELi0E" 2
bash tools/ncu_capture.sh ${TAG}_h2o64_tpqa h2o64 "eri_jk_tpqaILi1ELi0ELi0ELi0ELi1ELi1E|eri_jk_tpqaILi1ELi0ELi1ELi0ELi1ELi1E" 2
