#!/bin/bash
# round 2, fifth GPU pass: quantised GEMM with the A operand in tensor memory (tcgen05.st -> tcgen05.mma .ts), batched residual-add epilogue
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T() { ( time timeout "$1" python -m pytest "${@:2}" -q -p no:cacheprovider --durations=6 ) ; }
T 300 tests/test_gpu_batch.py -k "2]" -x > gpurun_out/r2e_t_batch_m2.log 2>&1; tail -14 gpurun_out/r2e_t_batch_m2.log
if ! grep -q " passed" gpurun_out/r2e_t_batch_m2.log || grep -q "failed" gpurun_out/r2e_t_batch_m2.log; then echo "quantised path failed: stopping"; exit 1; fi
T 400 tests/test_gpu_8b_shape.py > gpurun_out/r2e_t_8b.log 2>&1; tail -8 gpurun_out/r2e_t_8b.log
for B in 8 32 64; do timeout 200 python tools/batch_probe.py $B 576 8 2; done > gpurun_out/r2e_probe_q.log 2>&1
grep "^{" gpurun_out/r2e_probe_q.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_launches_q.csv python tools/batch_probe.py 32 576 1 2 > gpurun_out/r2e_ncu_q.log 2>&1
python - <<'PY'
import csv, collections
f = "gpurun_out/r2e_launches_q.csv"
try:
    rows = [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
    body = rows[1:][-261:]
    agg = collections.OrderedDict()
    for r in body:
        k = r[ki].split("(")[0][:60]; v = float(r[vi].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f, "last step: %.1f us over %d launches" % (tot / 1e3, sum(a[0] for a in agg.values())))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]): print("  %-62s n=%4d  %9.1f us  (%.1f us each)" % (k, a[0], a[1] / 1e3, a[1] / 1e3 / a[0]))
    for r in body[2:11] + body[-3:]: print("    %-50s %8.1f us" % (r[ki].split("(")[0][-48:], float(r[vi].replace(",", "")) / 1e3))
except Exception as ex:
    print(f, "failed:", ex)
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qgemm_kernel -s 386 -c 1 -o gpurun_out/r2e_qgemm_lmhead_full python tools/batch_probe.py 32 576 1 2 > gpurun_out/r2e_ncu_full_q.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qgemm_kernel -s 260 -c 1 -o gpurun_out/r2e_qgemm_gateup_full python tools/batch_probe.py 32 576 1 2 > gpurun_out/r2e_ncu_full_q2.log 2>&1
ls -la gpurun_out/r2e*.ncu-rep
( time GL_BENCH_WATCHDOG_S=100 timeout 420 python bench.py --workload config3 --steps 2 --warmup 1 --no-cpu ) > gpurun_out/r2e_bench_c3.json 2> gpurun_out/r2e_bench_c3.err
tail -8 gpurun_out/r2e_bench_c3.err; cut -c1-300 gpurun_out/r2e_bench_c3.json
