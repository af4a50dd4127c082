#!/bin/bash
# Round-2 session U: L1 prefetch hints in the bra-loop kernels (-DCF_TPQA_PF=1 records, 2 ket primitive rows, 4 density lines) vs the default build.
TAG=${TAG:-r2u}
mkdir -p gpurun_out
run() {  # name lib workload
  CHINIUM_FOCK_LIB=$PWD/chinium_b200/$2 timeout 600 python bench.py --workload $3 --steps 3 --warmup 2 --per-class --no-cpu-baseline > gpurun_out/${TAG}_$1_$3.json 2> gpurun_out/${TAG}_$1_$3.err
  echo "bench $1 $3 rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_$1_$3.json 2
}
for w in h2o64 fe4s4; do
  run base libchinium_fock.so $w
  for v in 1 2 4; do run pf$v libchinium_fock_pf$v.so $w; done
done
