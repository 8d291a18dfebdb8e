#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command, full captures of K1 (n = 1e8, p = 3), K1b (p = 48)
# and K3 (p = 48).  Reports are converted to CSV pages on the box (gpurun_out/ is limited to 64 MiB).
mkdir -p gpurun_out
export GSLNLS_WATCHDOG_S=20
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 24 --warmup 3 --no-cpu-baseline --e2e-fits 0 > gpurun_out/r02_launches_bench.json 2> gpurun_out/r02_launches.err
echo "launch list rc=$?"
cap() { # name kernel-regex bench-args...
  local name=$1 rx=$2; shift 2
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s 6 -c 2 -o /tmp/$name -f \
      python bench.py "$@" --steps 12 --warmup 3 --no-cpu-baseline --e2e-fits 0 > /dev/null 2> gpurun_out/$name.err
  echo "$name rc=$?"
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details --csv > gpurun_out/${name}_details.csv 2>/dev/null
}
cap r02_k1_full nls_pass
ncu -i /tmp/r02_k1_full.ncu-rep --page source --csv > gpurun_out/r02_k1_source.csv 2>/dev/null
cap r02_k1b_full nls_pass --config gaussmix48
ncu -i /tmp/r02_k1b_full.ncu-rep --page source --csv > gpurun_out/r02_k1b_source.csv 2>/dev/null
cap r02_k3_p48_full trs_step_warp --config gaussmix48
ls -la gpurun_out/r02_*
