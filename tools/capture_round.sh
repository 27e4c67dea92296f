#!/bin/bash
# Evidence of one round, taken in ONE gpurun call on one B200 (see profiles/README.md).
#   gpurun --timeout 2400 -- 'bash tools/capture_round.sh r2'
# Every ncu pass runs after the same command has exited 0 without ncu.  Output: gpurun_out/<tag>_*.
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1

python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.txt 2>&1; echo "pytest rc=$?"
tail -2 $out/${tag}_pytest_gpu.txt

python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
echo "reference rc=$?"

python tools/step_breakdown.py > $out/${tag}_step_breakdown.txt 2>&1
python tools/step_breakdown.py --config5 > $out/${tag}_step_breakdown_config5.txt 2>&1
python tools/scalar_latency.py > $out/${tag}_scalar_latency.txt 2>&1
python tools/continuum_bench.py 256 > $out/${tag}_continuum.jsonl 2>&1
python tools/band_cost.py 8 > $out/${tag}_band_cost.jsonl 2>&1
./tools/microbench > $out/${tag}_microbench.jsonl 2>&1
python bench.py --config 4 --steps 5 --warmup 3 > $out/${tag}_config4_n1.json 2> $out/${tag}_config4_n1.err
python bench.py --config 5 --steps 2 --warmup 1 > $out/${tag}_config5_n1.json 2> $out/${tag}_config5_n1.err
tail -1 $out/${tag}_step_breakdown.txt

# launch list of one bench step (serialised, cold cache: shares, not absolutes)
step="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-check"
$step > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/${tag}_launches.csv \
    $step > $out/${tag}_launches_run.log 2>&1
echo "launch list rc=$?"

# --set full of the summation kernel (five launches of a step), then of the near-zone and
# pedestal kernels on the largest gas, then of the coarse-grid kernels.  Reports are
# summarised on the box (details, raw metrics, hot SASS blocks, DRAM bytes); the reports
# themselves come back only while the merge limit of gpurun_out/ allows.
rep=/tmp/ncu_rep
mkdir -p $rep
ncu --set full --clock-control none --import-source on -k regex:sum_cell -s 10 -c 5 -f \
    -o $rep/${tag}_sum_cell $step > $out/${tag}_ncu_sum_cell.log 2>&1
echo "ncu sum_cell rc=$?"
python tools/one_gas.py CO2 2 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on \
    -k regex:"near_block|ped_nodes|ped_chain_runs|apply_kernel" -s 4 -c 4 -f \
    -o $rep/${tag}_near_ped python tools/one_gas.py CO2 2 > $out/${tag}_ncu_near_ped.log 2>&1
echo "ncu near/ped rc=$?"
python tools/one_gas.py CO2 2 --config5 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"sum_kernel|fixup_kernel" -s 2 -c 2 -f \
    -o $rep/${tag}_sum_direct python tools/one_gas.py CO2 2 --config5 > $out/${tag}_ncu_sum_direct.log 2>&1
echo "ncu direct rc=$?"
for r in sum_cell near_ped sum_direct; do
  ncu -i $rep/${tag}_$r.ncu-rep --page details > $out/${tag}_${r}_ncu_details.txt 2>&1
  ncu -i $rep/${tag}_$r.ncu-rep --page raw --csv > $out/${tag}_${r}_ncu_raw.csv 2>&1
done
python tools/ncu_dram.py $rep/${tag}_sum_cell.ncu-rep sum_cell $out/${tag}_sum_cell_dram.json
python tools/sass_hot.py $rep/${tag}_sum_cell.ncu-rep sum_cell 0.8 > $out/${tag}_sass_hot_blocks.txt 2>&1
for k in near_block ped_nodes ped_chain_runs; do
  python tools/sass_hot.py $rep/${tag}_near_ped.ncu-rep $k 1.0 >> $out/${tag}_sass_hot_blocks.txt 2>&1
done
python tools/sass_hot.py $rep/${tag}_sum_direct.ncu-rep sum_kernel 1.0 >> $out/${tag}_sass_hot_blocks.txt 2>&1

# index checks of our own (compute-sanitizer is closed on this pool): the GPU suite and a small
# case through every kernel on a build with device-side asserts at the indexing sites
# (python tools/build_variant.py bounds -DLBL_DEBUG_BOUNDS, before the call)
if [ -f variants/bounds.so ]; then
  cp pylbl_b200/libpylbl_b200.so /tmp/default.so
  cp variants/bounds.so pylbl_b200/libpylbl_b200.so
  python -m pytest tests -m gpu -q > $out/${tag}_bounds_check_pytest.txt 2>&1; echo "bounds pytest rc=$?"
  python tools/memcheck_case.py >> $out/${tag}_bounds_check_pytest.txt 2>&1; echo "bounds case rc=$?"
  cp /tmp/default.so pylbl_b200/libpylbl_b200.so
fi
# the reports, smallest first, while they fit
for r in $(ls -S -r $rep/*.ncu-rep); do
  used=$(du -sm $out | cut -f1); size=$(du -sm $r | cut -f1)
  if [ $((used + size)) -lt 56 ]; then cp $r $out/; fi
done
ls -la $out | tail -40
