#!/bin/bash
# round-2 GPU session j: where do the QFT-33 fan passes lose time?  Per-pass device times at 31..33 qubits
# (ncu launch list), fans on / off, full captures of a fan pass at 30 and 33 qubits.
mkdir -p gpurun_out
for n in 31 32 33; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r2j_qft${n}_launches.csv python run_configs.py qft --qubits $n > gpurun_out/r2j_qft${n}_launches.log 2>&1
  echo "QFT-$n per-pass ns:"; grep tma_pass gpurun_out/r2j_qft${n}_launches.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo
  grep -v tma_pass gpurun_out/r2j_qft${n}_launches.csv | awk -F'","' 'NR>1{print $5, $NF}' | tr -d '"' | sort | uniq -c | sort -rn | head -5
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r2j_qft33_nofans_launches.csv python run_configs.py qft --qubits 33 --option fans=0 > gpurun_out/r2j_qft33_nofans_launches.log 2>&1
echo "QFT-33 fans=0 per-pass ns:"; grep tma_pass gpurun_out/r2j_qft33_nofans_launches.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tma_pass -s 1 -c 2 -o gpurun_out/r2j_qft30_pass -f python run_configs.py qft --qubits 30 > gpurun_out/r2j_ncu_qft30.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:tma_pass -s 1 -c 1 -o gpurun_out/r2j_qft33_pass -f python run_configs.py qft --qubits 33 > gpurun_out/r2j_ncu_qft33.log 2>&1
tail -2 gpurun_out/r2j_ncu_qft33.log
ls -la gpurun_out/*.ncu-rep
